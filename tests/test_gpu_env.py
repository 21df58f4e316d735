"""Environment emitters (envmap / constant) on the CUDA path, through the C ABI, against the CPU
oracle (which tests/test_oracle_golden.py pins to the reference's tables and renders).
Tolerances as in test_gpu_parity.py; the emitter tables are computed with the same fp32 operation
order on both sides and must agree to a few ulp."""
import numpy as np
import pytest

from conftest import compare_images, env_scene, golden

import mitsuba3_b200 as mb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle_mod(built):
    from oracle import oracle
    return oracle


@pytest.mark.parametrize("kind", ["envmap", "constant"])
def test_environment_emitter_tables(kind, oracle_mod):
    from mitsuba3_b200.integrators import device_scene
    g = golden("env.npz")
    sc = mb.load_dict(env_scene(kind=kind))
    q = np.concatenate([g[f"{kind}_p"], g[f"{kind}_u"], g[f"{kind}_dirs"]], 1)
    rng = np.random.default_rng(5)
    extra = np.concatenate([(rng.random((4096, 3)) - 0.5) * 6, rng.random((4096, 2)), rng.normal(size=(4096, 3))], 1).astype(np.float32)
    extra[:, 5:8] /= np.linalg.norm(extra[:, 5:8], axis=1, keepdims=True)
    q = np.concatenate([q, extra], 0)
    out = device_scene(sc).env_query(q)
    ref = oracle_mod.OracleScene(sc).env_query(q)
    scale = np.maximum(np.abs(ref), 1e-3)
    err = np.abs(out - ref) / scale
    assert np.isfinite(out).all()
    assert err.max() < 2e-6, (kind, err.max(), np.unravel_index(np.argmax(err), err.shape))
    # and against the reference's own table (same tolerances as the oracle's golden test)
    gref = g[f"{kind}_sample"]
    n = gref.shape[0]
    assert np.abs(out[:n, 0:3] - gref[:, 0:3]).max() < 1e-4
    e = np.abs(out[:n, 7:14] - gref[:, 7:14]) / np.maximum(np.abs(gref[:, 7:14]), 1e-3)
    assert e.max() < 1e-4


ENV_CASES = [dict(kind="envmap"), dict(kind="envmap", hide=True, max_depth=3), dict(kind="envmap", area_light=True),
             dict(kind="constant"), dict(kind="constant", area_light=True, max_depth=4),
             dict(kind="envmap", area_light=True, integrator="prb")]


@pytest.mark.parametrize("case", ENV_CASES, ids=lambda c: "-".join(f"{k}{v}" for k, v in c.items()))
def test_environment_render_matches_oracle(case, oracle_mod):
    """Escaped rays -> environment queue -> k_shade_env (MIS with the previous BSDF sample), environment
    NEE inside the material kernels, hide_emitters, path and prb estimators."""
    sc = mb.load_dict(env_scene(res=48, spp=16, **case))
    img = mb.render(sc, spp=16, seed=3)
    ref = oracle_mod.OracleScene(sc).render(spp=16, seed=3, mode=0)
    assert ref.mean() > 0.1
    compare_images(img, ref, max_bad_frac=0.01)


def test_environment_gaussian_filter_and_sharding(oracle_mod):
    """Gaussian splat + two pixel-tile shards accumulated into one film = the single-pass image."""
    d = env_scene(res=64, spp=8, area_light=True)
    d["sensor"]["film"]["rfilter"] = {"type": "gaussian"}
    sc = mb.load_dict(d)
    img = mb.render(sc, spp=8, seed=1)
    ref = oracle_mod.OracleScene(sc).render(spp=8, seed=1, mode=0)
    compare_images(img, ref, max_bad_frac=0.01)


def test_constant_environment_radiance_gradient(oracle_mod):
    """PRB adjoint: d/d(radiance) of the constant environment = Le term (k_shade_env) + NEE term."""
    from mitsuba3_b200.integrators import PRBIntegrator
    sc = mb.load_dict(env_scene(kind="constant", res=32, spp=8, area_light=True, integrator="prb"))
    gi = np.random.default_rng(2).random(sc.film_shape).astype(np.float32) * 1e-2
    g = PRBIntegrator(max_depth=4).render_backward(sc, gi, seed=5, spp=8)
    o = oracle_mod.OracleScene(sc); o.grad_zero(); o.render_backward(gi, spp=8, seed=5, max_depth=4)
    for k in ("sky.radiance.value", "lamp.emitter.radiance.value", "grey.reflectance.value"):
        ref_g = o.grad(sc.parameters()[k])
        assert np.abs(ref_g).max() > 0
        assert np.abs(g[k] - ref_g).max() / np.abs(ref_g).max() < 5e-3, (k, g[k], ref_g)


def test_scenes_with_different_shared_memory_needs_coexist(oracle_mod):
    """The dynamic shared-memory attribute of the kernels is process-wide: creating a second scene
    that stages less into shared memory must not break launches of the first one."""
    from conftest import materials_cbox
    a = mb.load_dict(materials_cbox(res=32, spp=8, max_depth=6))
    ia = mb.render(a, spp=8, seed=0)
    b = mb.load_dict(env_scene(res=32, spp=8))
    mb.render(b, spp=8, seed=0)
    assert np.array_equal(mb.render(a, spp=8, seed=0), ia)


def test_envmap_data_gradient_and_update(oracle_mod):
    """PRB w.r.t. the envmap `data` parameter: the escaped-ray term (k_shade_env) and the NEE term scatter
    into the four real texels of EnvironmentMapEmitter::eval_spectrum (halo columns routed back,
    envmap.cpp:228-246); forward mode is the transpose; updating `data` rebuilds halo + warp."""
    from mitsuba3_b200.integrators import PRBIntegrator, update_params
    sc = mb.load_dict(env_scene(res=32, spp=8, area_light=True, integrator="prb", max_depth=4))
    P = sc.parameters(); i = P["sky.data"]
    integ = PRBIntegrator(max_depth=4)
    rng = np.random.default_rng(3)
    gi = rng.random(sc.film_shape).astype(np.float32) * 1e-2
    g = integ.render_backward(sc, gi, seed=5, spp=8)
    o = oracle_mod.OracleScene(sc); o.grad_zero(); o.render_backward(gi, spp=8, seed=5, max_depth=4)
    ref_g = o.grad(i)
    assert np.abs(ref_g).max() > 0
    assert np.abs(g["sky.data"] - ref_g).max() / np.abs(ref_g).max() < 5e-3
    # forward mode = transpose of backward
    v = rng.random(ref_g.shape).astype(np.float32)
    fwd = integ.render_forward(sc, {"sky.data": v}, seed=5, spp=8)
    lhs, rhs = float((gi.astype(np.float64) * fwd).sum()), float((g["sky.data"].astype(np.float64) * v).sum())
    assert abs(lhs - rhs) <= 2e-3 * abs(lhs), (lhs, rhs)
    # parameter update: new data -> same image as a scene created with that data
    new = (sc.textures[i].data * (0.5 + rng.random(sc.textures[i].data.shape))).astype(np.float32)
    update_params(sc, {"sky.data": new})
    img = mb.render(sc, spp=8, seed=2)
    sc2 = mb.load_dict(env_scene(res=32, spp=8, area_light=True, integrator="prb", max_depth=4, img=new))
    assert np.array_equal(img, mb.render(sc2, spp=8, seed=2))
    compare_images(img, oracle_mod.OracleScene(sc2).render(spp=8, seed=2, mode=0), max_bad_frac=0.01)


def test_smooth_mesh_scene_matches_oracle(oracle_mod):
    from conftest import smooth_mesh_scene
    sc = mb.load_dict(smooth_mesh_scene(res=48, spp=16, max_depth=5))
    img = mb.render(sc, spp=16, seed=2)
    compare_images(img, oracle_mod.OracleScene(sc).render(spp=16, seed=2, mode=0), max_bad_frac=0.01)


@pytest.mark.parametrize("mat,tag", [("aniso_principled", ""), ("aniso_principled", "_m"), ("aniso_roughconductor", "_m"), ("aniso_roughconductor", "")])
def test_anisotropic_mesh_with_packed_tangent_frames_matches_oracle(oracle_mod, mat, tag):
    """Packed tangent frames (mesh.cpp:2339-2351,2417-2429, interaction.h:571-597): anisotropic principled / rough conductor on
    the UV sphere of tests/golden/tangent_mesh.npz (frames and FaceUVFlipped bits as the reference's loader packs them; the
    oracle reproduces the reference's frames to 2e-6 and its renders per pixel, tests/test_oracle_golden.py)."""
    from conftest import tangent_mesh_scene
    sc = mb.load_dict(tangent_mesh_scene(mat, tag, res=48, spp=16, max_depth=5))
    img = mb.render(sc, spp=16, seed=3)
    compare_images(img, oracle_mod.OracleScene(sc).render(spp=16, seed=3, mode=0), max_bad_frac=0.01)


def test_reference_envmap_lookup_known_answers_through_the_abi():
    """src/emitters/tests/test_envmap.py:201-263 through b200pt_env_query: exact texel-centre / align-corners read-back of a ramp,
    continuity across the phi seam, clamping poles (the oracle passes the same checks in tests/test_oracle_golden.py)."""
    from mitsuba3_b200.integrators import device_scene

    def env_eval(img, u, v):
        d = env_scene(img=np.asarray(img, np.float32))
        d["sky"] = {"type": "envmap", "bitmap": np.asarray(img, np.float32)}
        sc = mb.load_dict(d)
        u, v = np.atleast_1d(np.asarray(u, np.float64)), np.atleast_1d(np.asarray(v, np.float64))
        phi, theta = u * 2 * np.pi, v * np.pi
        dirs = np.stack([np.sin(phi) * np.sin(theta), np.cos(theta), -np.cos(phi) * np.sin(theta)], 1).astype(np.float32)
        q = np.concatenate([np.zeros((len(u), 3), np.float32), np.full((len(u), 2), 0.5, np.float32), dirs], 1)
        return device_scene(sc).env_query(q)[:, 14:17]

    W, H = 16, 8
    img = np.broadcast_to(np.arange(W, dtype=np.float32)[None, :, None], (H, W, 3)).copy()
    t = np.linspace(1.0 / W, 1.0 - 1.0 / W, 50)
    assert np.abs(env_eval(img, t, np.full_like(t, 0.5))[:, 0] - (t * W - 0.5)).max() < 1e-4
    W, H = 8, 16
    img = np.broadcast_to(np.arange(H, dtype=np.float32)[:, None, None], (H, W, 3)).copy()
    t = np.linspace(1.0 / H, 1.0 - 1.0 / H, 50)
    assert np.abs(env_eval(img, np.zeros_like(t), t)[:, 0] - t * (H - 1)).max() < 1e-4
    W, H = 64, 8
    img = np.broadcast_to(np.cos(2 * np.pi * np.arange(W) / W).astype(np.float32)[None, :, None], (H, W, 3)).copy()
    u = np.linspace(-0.1, 0.1, 201)
    assert np.abs(env_eval(img, u, np.full_like(u, 0.5))[:, 0] - np.cos(2 * np.pi * ((u - np.floor(u)) * W - 0.5) / W)).max() < 5e-3
    W, H = 8, 16
    img = np.zeros((H, W, 3), np.float32); img[0] = 10.0; img[H - 1] = 20.0
    assert abs(env_eval(img, 0.0, 0.0)[0, 0] - 10.0) < 1e-3 and abs(env_eval(img, 0.0, 1.0)[0, 0] - 20.0) < 1e-3
