"""Host-side TEA (core/random.h:77-90) -- used for ``seed_grad`` derivation
(util.py:505-507) and by the sharding logic's self checks."""


def sample_tea_32(v0: int, v1: int, rounds: int = 4):
    M = 0xffffffff
    v0 &= M; v1 &= M
    s = 0
    for _ in range(rounds):
        s = (s + 0x9e3779b9) & M
        v0 = (v0 + ((((v1 << 4) & M) + 0xa341316c) ^ ((v1 + s) & M) ^ ((v1 >> 5) + 0xc8013ea4))) & M
        v1 = (v1 + ((((v0 << 4) & M) + 0xad90777d) ^ ((v0 + s) & M) ^ ((v0 >> 5) + 0x7e95761e))) & M
    return v0, v1
