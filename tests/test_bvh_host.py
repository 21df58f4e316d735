"""CPU tests of the product's host-side BVH builder (mitsuba3_b200/csrc/bvh.cpp, layout in bvh.h).

The builder replaces the acceleration-structure build the reference delegates to Embree / its kd-tree
(scene.cpp:93); the reference's own guarantee is just "ray_intersect finds the closest triangle", so the
tests pin exactly that: structural invariants of the node array plus tree walk == brute force over all
triangles, on the inputs the reference's mesh tests use as edge cases (single triangle, degenerate and
duplicate triangles, large coordinate ranges) and on the bench geometries.
"""
import ctypes
import os
import subprocess
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "mitsuba3_b200", "csrc")
MAX_LEAF = 2          # bvh.h: BVH_MAX_LEAF


def load_harness():
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libbvh_harness.so")
    srcs = [os.path.join(HERE, "bvh_harness.cpp"), os.path.join(CSRC, "bvh.cpp")]
    deps = srcs + [os.path.join(CSRC, "bvh.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-I", CSRC, "-o", so] + srcs, check=True)
    lib = ctypes.CDLL(so)
    lib.bvh_h_build.restype = ctypes.c_void_p
    lib.bvh_h_build.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
    lib.bvh_h_free.argtypes = [ctypes.c_void_p]
    for f in (lib.bvh_h_n_nodes, lib.bvh_h_depth):
        f.restype = ctypes.c_uint32
        f.argtypes = [ctypes.c_void_p]
    lib.bvh_h_nodes.restype = ctypes.c_void_p
    lib.bvh_h_nodes.argtypes = [ctypes.c_void_p]
    lib.bvh_h_order.restype = ctypes.POINTER(ctypes.c_uint32)
    lib.bvh_h_order.argtypes = [ctypes.c_void_p]
    lib.bvh_h_validate.restype = ctypes.c_int
    lib.bvh_h_validate.argtypes = [ctypes.c_void_p, ctypes.c_uint32]
    lib.bvh_h_trace.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return lib


@pytest.fixture(scope="module")
def harness():
    return load_harness()


class Tree:
    def __init__(self, lib, tri):
        self.lib = lib
        self.tri = np.ascontiguousarray(tri, np.float32).reshape(-1, 9)
        self.h = lib.bvh_h_build(self.tri.ctypes.data, len(self.tri))

    def __del__(self):
        self.lib.bvh_h_free(self.h)

    @property
    def n_nodes(self):
        return self.lib.bvh_h_n_nodes(self.h)

    def nodes(self):
        raw = (ctypes.c_uint8 * (64 * self.n_nodes)).from_address(self.lib.bvh_h_nodes(self.h))
        return np.frombuffer(raw, np.float32).reshape(-1, 16).copy()

    def order(self):
        return np.ctypeslib.as_array(self.lib.bvh_h_order(self.h), (len(self.tri),)).copy()

    def validate(self):
        return self.lib.bvh_h_validate(self.h, MAX_LEAF)

    def trace(self, rays, brute):
        rays = np.ascontiguousarray(rays, np.float32)
        t = np.empty(len(rays), np.float32)
        prim = np.empty(len(rays), np.uint32)
        n = ctypes.c_uint64(0)
        steps = ctypes.c_uint64(0)
        self.lib.bvh_h_trace(self.h, len(rays), rays.ctypes.data, int(brute), t.ctypes.data, prim.ctypes.data, ctypes.byref(n), ctypes.byref(steps))
        self.last_steps = steps.value
        return t, prim, n.value



def random_rays(rng, n, lo, hi, tri):
    """Half of the rays aim at a random point of the bounding box, half at a random point of a random triangle."""
    o = rng.uniform(lo - 0.5 * (hi - lo), hi + 0.5 * (hi - lo), (n, 3))
    target = rng.uniform(lo, hi, (n, 3))
    V = tri.reshape(-1, 3, 3).astype(np.float64)
    w = rng.dirichlet([1, 1, 1], n)
    on_tri = (V[rng.integers(0, len(V), n)] * w[:, :, None]).sum(1)
    target = np.where(rng.random((n, 1)) < 0.5, on_tri, target)
    d = target - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    maxt = np.where(rng.random(n) < 0.25, rng.uniform(0.1, 2.0, n) * np.linalg.norm(hi - lo), np.inf)
    return np.concatenate([o, d, maxt[:, None]], axis=1).astype(np.float32)


def soup(rng, n, extent=1.0, size=0.2):
    c = rng.uniform(-extent, extent, (n, 1, 3))
    return (c + rng.uniform(-size, size, (n, 3, 3))).astype(np.float32).reshape(n, 9)


def grid(n):
    """n x n heightfield quads, two triangles each (the bench's heightfield geometry in miniature)."""
    xs = np.linspace(-1, 1, n + 1, dtype=np.float32)
    X, Z = np.meshgrid(xs, xs)
    Y = (0.08 * np.sin(5 * X) * np.cos(5 * Z)).astype(np.float32)
    P = np.stack([X, Y, Z], -1)
    a, b, c, d = P[:-1, :-1], P[:-1, 1:], P[1:, :-1], P[1:, 1:]
    t1 = np.stack([a, b, c], 2).reshape(-1, 9)
    t2 = np.stack([b, d, c], 2).reshape(-1, 9)
    return np.concatenate([t1, t2]).astype(np.float32)


CASES = {
    "single": lambda rng: soup(rng, 1),
    "two": lambda rng: soup(rng, 2),
    "three": lambda rng: soup(rng, 3),
    "seven": lambda rng: soup(rng, 7),
    "soup_1k": lambda rng: soup(rng, 1000),
    "soup_20k_small": lambda rng: soup(rng, 20000, size=0.02),
    "grid_48": lambda rng: grid(48),
    "duplicates": lambda rng: np.repeat(soup(rng, 5), 40, axis=0),                 # identical centroids: median-split path
    "degenerate": lambda rng: np.concatenate([soup(rng, 50), np.tile(rng.uniform(-1, 1, (20, 1, 3)).astype(np.float32), (1, 3, 1)).reshape(20, 9)]),
    "coplanar": lambda rng: soup(rng, 300) * np.array([1, 0, 1] * 3, np.float32),   # zero-thickness boxes
    "far_from_origin": lambda rng: soup(rng, 500) + np.float32(1000.0),
    "huge_range": lambda rng: soup(rng, 400, extent=1e4, size=50.0),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_invariants_and_walk_equals_brute_force(harness, name):
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    tri = CASES[name](rng)
    tree = Tree(harness, tri)
    assert tree.validate() == 0
    assert sorted(tree.order().tolist()) == list(range(len(tri)))
    lo, hi = tri.reshape(-1, 3).min(0), tri.reshape(-1, 3).max(0)
    hi = np.maximum(hi, lo + 1e-3)
    rays = random_rays(rng, 3000 if len(tri) <= 2000 else 600, lo.astype(np.float64), hi.astype(np.float64), tri)
    t_tree, p_tree, n_tree = tree.trace(rays, brute=False)
    t_ref, p_ref, n_ref = tree.trace(rays, brute=True)
    assert np.array_equal(p_tree, p_ref)
    assert np.array_equal(t_tree, t_ref)
    assert (p_ref != 0xFFFFFFFF).sum() >= 20                      # the rays do hit something
    if len(tri) >= 1000:
        assert n_tree < n_ref // 10                               # and the tree actually culls


def test_empty_scene_has_a_root_that_is_never_entered(harness):
    tree = Tree(harness, np.zeros((0, 9), np.float32))
    assert tree.n_nodes == 1
    nodes = tree.nodes()
    assert nodes.view(np.int32)[0, 12] == 0x7FFFFFFF and nodes.view(np.int32)[0, 13] == 0x7FFFFFFF
    t, p, _ = tree.trace(np.array([[0, 0, 0, 0, 0, 1, np.inf]], np.float32), brute=False)
    assert np.isinf(t[0]) and p[0] == 0xFFFFFFFF


def test_layout_breadth_first_and_leaf_encoding(harness):
    rng = np.random.default_rng(5)
    tri = soup(rng, 513)
    tree = Tree(harness, tri)
    nodes = tree.nodes()
    ints = nodes.view(np.int32)
    children = ints[:, 12:14]
    inner = children[(children >= 0) & (children != 0x7FFFFFFF)]
    # breadth-first numbering: the inner children are exactly 1 .. n-1, in order of appearance
    assert np.array_equal(inner, np.arange(1, tree.n_nodes))
    leaves = ~children[children < 0]
    counts = (leaves & 7) + 1
    assert counts.max() <= MAX_LEAF and counts.sum() == len(tri)
    assert tree.lib.bvh_h_depth(tree.h) < 40
    # padding words are zero (the kernels read whole 16-byte vectors)
    assert not ints[:, 14:16].any()


def test_boxes_are_conservative_for_grazing_hits(harness):
    """Rays aimed exactly at triangle edges and vertices: the inflated boxes must never cull a hit the
    triangle test accepts (bvh.h: boxes are inflated by a few ulps)."""
    rng = np.random.default_rng(11)
    tri = grid(16)
    tree = Tree(harness, tri)
    V = tri.reshape(-1, 3, 3)
    pick = rng.integers(0, len(V), 4000)
    w = rng.random((4000, 1))
    on_edge = V[pick, 0] * w + V[pick, 1] * (1 - w)
    target = np.where(rng.random((4000, 1)) < 0.3, V[pick, 2], on_edge)
    o = np.array([0.3, 2.0, -0.2]) + rng.normal(0, 0.5, (4000, 3)) * [1, 0.1, 1]
    d = target - o
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([o, d, np.full((4000, 1), np.inf)], 1).astype(np.float32)
    t_tree, p_tree, _ = tree.trace(rays, brute=False)
    t_ref, p_ref, _ = tree.trace(rays, brute=True)
    assert np.array_equal(p_tree, p_ref) and np.array_equal(t_tree, t_ref)
    # axis-parallel rays (direction components exactly +0 / -0: the reciprocal is the signed 1e30 of safe_inv)
    n = 1500
    o = rng.uniform(-1.2, 1.2, (n, 3))
    axis = rng.integers(0, 3, n)
    sign = np.where(rng.random(n) < 0.5, 1.0, -1.0)
    d = np.zeros((n, 3))
    d[np.arange(n), axis] = sign
    d[rng.random((n, 3)) < 0.3] *= -1.0          # sprinkle negative zeros
    o[np.arange(n), axis] = -2.0 * sign
    o[axis == 1, 1] = 2.0 * sign[axis == 1]
    d[axis == 1, 1] = -sign[axis == 1]
    rays = np.concatenate([o, d, np.full((n, 1), np.inf)], 1).astype(np.float32)
    t_ref, p_ref, _ = tree.trace(rays, brute=True)
    t_o, p_o, _ = tree.trace(rays, brute=False)
    assert np.array_equal(p_o, p_ref) and np.array_equal(t_o, t_ref)
    assert (p_ref != 0xFFFFFFFF).sum() > 200
