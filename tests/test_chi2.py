"""Chi-square tests of the samplers against their densities, in the spirit of the reference's own
BSDF / emitter tests (src/bsdfs/tests/test_rough_conductor.py:21-96, test_rough_dielectric.py:25-206,
test_principled.py:20-120, src/emitters/tests/test_envmap.py:14-45 -> mitsuba.chi2.ChiSquareTest):
histogram the sampled directions on a (cos theta, phi) grid and compare with the integrated pdf.
CPU: the oracle; GPU (marked): the CUDA tables through the C ABI."""
import numpy as np
import pytest
from scipy import stats

from conftest import ROUGH_SPECS, env_scene

import mitsuba3_b200 as mb

RES = (12, 24)          # bins in cos(theta) in [-1, 1] and phi
SUB = 16                # sub-samples per bin edge for the integrated density (sharp GTR1 / GGX lobes)


def _chi2(wo, expected_fn, n):
    z, phi = np.clip(wo[:, 2], -1, 1), np.mod(np.arctan2(wo[:, 1], wo[:, 0]), 2 * np.pi)
    iz = np.minimum(((z + 1) / 2 * RES[0]).astype(int), RES[0] - 1); ip = np.minimum((phi / (2 * np.pi) * RES[1]).astype(int), RES[1] - 1)
    obs = np.bincount(iz * RES[1] + ip, minlength=RES[0] * RES[1]).astype(np.float64)
    # integrated pdf per bin: midpoint rule on a SUB x SUB lattice, d(omega) = dz dphi
    zz = (np.arange(RES[0] * SUB) + 0.5) / (RES[0] * SUB) * 2 - 1
    pp = (np.arange(RES[1] * SUB) + 0.5) / (RES[1] * SUB) * 2 * np.pi
    Z, P = np.meshgrid(zz, pp, indexing="ij")
    r = np.sqrt(np.maximum(0, 1 - Z * Z))
    dirs = np.stack([r * np.cos(P), r * np.sin(P), Z], -1).reshape(-1, 3).astype(np.float32)
    pdf = expected_fn(dirs).reshape(RES[0], SUB, RES[1], SUB).astype(np.float64)
    exp = pdf.sum(axis=(1, 3)) * (2.0 / (RES[0] * SUB)) * (2 * np.pi / (RES[1] * SUB)) * n
    exp = exp.reshape(-1)
    # pool sparse bins (expected < 5), as mitsuba.chi2 does
    order = np.argsort(exp)
    pooled_o, pooled_e, acc_o, acc_e = [], [], 0.0, 0.0
    for i in order:
        acc_o += obs[i]; acc_e += exp[i]
        if acc_e >= 5:
            pooled_o.append(acc_o); pooled_e.append(acc_e); acc_o = acc_e = 0.0
    if acc_e > 0 and pooled_e:
        pooled_o[-1] += acc_o; pooled_e[-1] += acc_e
    o, e = np.array(pooled_o), np.array(pooled_e)
    stat = ((o - e) ** 2 / e).sum()
    p = stats.chi2.sf(stat, len(e) - 1)
    return p, obs.sum(), exp.sum()


def _bsdf_case(backend, spec, wi, n=120000, seed=0):
    from test_gpu_parity import _bsdf_scene
    sc, idx = _bsdf_scene(spec)
    rng = np.random.default_rng(seed)
    q = np.zeros((n, 11), np.float32); q[:, 0:3] = wi; q[:, 3:6] = [0, 0, 1]; q[:, 8:11] = rng.random((n, 3))
    out = backend(sc).bsdf_eval_pdf_sample(idx, q)
    ok = (out[:, 10:13] > 0).any(axis=1)                         # as mitsuba.chi2.BSDFAdapter: samples with a zero weight are rejected ones
    def expected(dirs):
        qq = np.zeros((dirs.shape[0], 11), np.float32); qq[:, 0:3] = wi; qq[:, 3:6] = dirs; qq[:, 8:11] = 0.5
        return backend(sc).bsdf_eval_pdf_sample(idx, qq)[:, 3]
    p, n_obs, n_exp = _chi2(out[ok, 4:7], expected, n)
    assert abs(n_obs - n_exp) < 0.02 * n, (n_obs, n_exp)         # the density integrates to the sampled fraction
    assert p > 1e-3, p


def _oracle_backend(sc):
    from oracle import oracle
    return oracle.OracleScene(sc)


def _gpu_backend(sc):
    from mitsuba3_b200.integrators import device_scene
    return device_scene(sc)


WI = np.array([1.0, 1.0, 1.0], np.float32) / np.sqrt(3).astype(np.float32)
BSDF_CASES = {
    "roughconductor_beckmann": (ROUGH_SPECS["roughconductor_beckmann_rough"], WI),
    "roughconductor_ggx_aniso": (ROUGH_SPECS["roughconductor_ggx_aniso"], WI),
    "roughdielectric_beckmann": ({"type": "roughdielectric", "alpha": 0.3}, WI),
    "roughdielectric_ggx_inside": ({"type": "roughdielectric", "distribution": "ggx", "alpha": 0.4, "int_ior": 1.5, "ext_ior": 1.0}, WI * np.array([1, 1, -1], np.float32)),
    "principled_rough_metal": ({"type": "principled", "base_color": {"type": "rgb", "value": [0.9, 0.6, 0.2]}, "metallic": 0.8, "roughness": 0.35, "specular": 0.5}, WI),
    "principled_trans_coat": ({"type": "principled", "roughness": 0.5, "metallic": 0.2, "clearcoat": 1.0, "spec_trans": 0.5, "eta": 1.5}, np.array([0.7071, 0, 0.7071], np.float32)),
    "diffuse": ({"type": "diffuse"}, WI),
}


@pytest.mark.parametrize("name", sorted(BSDF_CASES))
def test_bsdf_sampling_matches_pdf_oracle(name):
    _bsdf_case(_oracle_backend, *BSDF_CASES[name])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(BSDF_CASES))
def test_bsdf_sampling_matches_pdf_gpu(name, built):
    _bsdf_case(_gpu_backend, *BSDF_CASES[name], seed=1)


def _env_case(backend, n=120000, seed=0):
    sc = mb.load_dict(env_scene(kind="envmap"))
    b = backend(sc)
    rng = np.random.default_rng(seed)
    q = np.zeros((n, 8), np.float32); q[:, 3:5] = rng.random((n, 2)); q[:, 5:8] = [0, 1, 0]
    out = b.env_query(q)
    def expected(dirs):
        qq = np.zeros((dirs.shape[0], 8), np.float32); qq[:, 5:8] = dirs
        return b.env_query(qq)[:, 17]
    p, n_obs, n_exp = _chi2(out[:, 0:3], expected, n)
    assert abs(n_obs - n_exp) < 0.02 * n, (n_obs, n_exp)
    assert p > 1e-3, p


def test_envmap_sampling_matches_pdf_oracle():
    _env_case(_oracle_backend)


@pytest.mark.gpu
def test_envmap_sampling_matches_pdf_gpu(built):
    _env_case(_gpu_backend, seed=1)
