"""PRB adjoint (b200pt_render_backward): against the oracle's adjoint at equal seeds,
and against central finite differences of the primal renderer (the criterion of
src/integrators/tests/test_ad_integrators.py:102-150,1360-1396)."""
import numpy as np
import pytest

from conftest import cbox

import mitsuba3_b200 as mb

pytestmark = pytest.mark.gpu


def _prb_scene(res=32, rfilter="box", tex=None):
    d = cbox(res=res, rfilter=rfilter, spp=16, max_depth=5)
    d["integrator"] = {"type": "prb", "max_depth": 5}
    if tex is not None:
        d["tex"] = {"type": "diffuse", "reflectance": {"type": "bitmap", "data": tex, "raw": True, "filter_type": "bilinear", "wrap_mode": "clamp"}}
        d["back"]["bsdf"] = {"type": "ref", "id": "tex"}
    return mb.load_dict(d)


@pytest.mark.parametrize("rfilter", ["box", "gaussian"])
def test_backward_matches_oracle(built, rfilter):
    from oracle import oracle
    from mitsuba3_b200.integrators import PRBIntegrator
    tex = (0.3 + 0.4 * np.random.default_rng(1).random((8, 8, 3))).astype(np.float32)
    sc = _prb_scene(rfilter=rfilter, tex=tex)
    grad_in = np.random.default_rng(2).normal(size=sc.film_shape).astype(np.float32) * 1e-2
    g = PRBIntegrator(max_depth=5).render_backward(sc, grad_in, seed=4, spp=16)
    o = oracle.OracleScene(sc); o.grad_zero(); o.render_backward(grad_in, spp=16, seed=4, max_depth=5)
    names = sc.parameters()
    for k, i in names.items():
        ref = o.grad(i)
        scale = np.abs(ref).max() + 1e-12
        assert np.abs(g[k] - ref).max() / scale < 5e-3, (k, g[k], ref)
    assert np.abs(g["tex.reflectance.data"]).max() > 0
    assert np.abs(g["light.emitter.radiance.value"]).max() > 0


def test_backward_checkerboard_colors(built):
    """Gradients w.r.t. the two colours of a checkerboard reflectance (checkerboard.cpp) vs the oracle."""
    from oracle import oracle
    from mitsuba3_b200.integrators import PRBIntegrator
    d = cbox(res=32, rfilter="box", spp=16, max_depth=4); d["integrator"] = {"type": "prb", "max_depth": 4}
    d["chk"] = {"type": "diffuse", "reflectance": {"type": "checkerboard", "color0": {"type": "rgb", "value": [0.8, 0.3, 0.2]},
                                                    "color1": {"type": "rgb", "value": [0.2, 0.4, 0.9]}, "to_uv": [[5, 0, 0], [0, 5, 0], [0, 0, 1]]}}
    d["floor"]["bsdf"] = {"type": "ref", "id": "chk"}
    sc = mb.load_dict(d)
    grad_in = np.random.default_rng(3).random(sc.film_shape).astype(np.float32) * 1e-2
    g = PRBIntegrator(max_depth=4).render_backward(sc, grad_in, seed=7, spp=16)
    o = oracle.OracleScene(sc); o.grad_zero(); o.render_backward(grad_in, spp=16, seed=7, max_depth=4)
    ref = o.grad(sc.parameters()["chk.reflectance.colors"])
    got = g["chk.reflectance.colors"]
    assert got.shape == (2, 3) and np.abs(ref).min() > 0
    assert np.abs(got - ref).max() / np.abs(ref).max() < 5e-3, (got, ref)


def test_backward_matches_finite_differences(built):
    """d(mean image)/d(albedo, radiance) vs central differences of the primal renderer."""
    from mitsuba3_b200.integrators import PRBIntegrator, update_params
    sc = _prb_scene(res=32)
    integ = PRBIntegrator(max_depth=5)
    H, W, _ = sc.film_shape
    grad_in = np.full(sc.film_shape, 1.0 / (H * W * 3), np.float32)        # loss = mean(image)
    g = integ.render_backward(sc, grad_in, seed=1, spp=256)
    eps = 1e-2
    for key, ch in [("white.reflectance.value", 0), ("red.reflectance.value", 0), ("green.reflectance.value", 1), ("light.emitter.radiance.value", 2)]:
        i = sc.parameters()[key]
        base = sc.textures[i].value.copy()
        vals = []
        for sgn in (+1, -1):
            v = base.copy(); v[ch] += sgn * eps * max(1.0, abs(base[ch]))
            update_params(sc, {key: v})
            vals.append(float(integ.render(sc, seed=3, spp=1024).mean()))       # same seed for +-eps
        update_params(sc, {key: base})
        fd = (vals[0] - vals[1]) / (2 * eps * max(1.0, abs(base[ch])))
        assert abs(g[key][ch] - fd) / max(abs(fd), 1e-3) < 0.05, (key, g[key][ch], fd)


def test_backward_is_linear_and_accumulates(built):
    from mitsuba3_b200.integrators import PRBIntegrator
    sc = _prb_scene(res=16)
    integ = PRBIntegrator(max_depth=4)
    gi = np.random.default_rng(0).random(sc.film_shape).astype(np.float32)
    g1 = integ.render_backward(sc, gi, seed=2, spp=8)
    g2 = integ.render_backward(sc, 2 * gi, seed=2, spp=8)
    g3 = integ.render_backward(sc, gi, seed=2, spp=8, zero=False)          # accumulates on top of g2
    for k in g1:
        assert np.allclose(g2[k], 2 * g1[k], rtol=2e-4, atol=1e-7)
        assert np.allclose(g3[k], 3 * g1[k], rtol=2e-4, atol=1e-7)


def test_inverse_rendering_loop_converges(built):
    """The optimisation loop of tutorials/inverse_rendering/gradient_based_opt.ipynb with torch as the
    AD currency: recover the red wall's albedo from a reference image (render_torch = mi.render with
    _RenderOp semantics: primal with `seed`, adjoint with sample_tea_32(seed, 1)[0])."""
    import torch
    from mitsuba3_b200.integrators import PRBIntegrator, render_torch, update_params
    d = cbox(res=64, spp=32, max_depth=4); d["integrator"] = {"type": "prb", "max_depth": 4}
    sc = mb.load_dict(d)
    key = "red.reflectance.value"
    target = sc.textures[sc.parameters()[key]].value.copy()
    integ = PRBIntegrator(max_depth=4)
    ref = torch.from_numpy(integ.render(sc, seed=1234, spp=256))
    p = torch.tensor([0.3, 0.3, 0.3], requires_grad=True)
    opt = torch.optim.Adam([p], lr=0.05)
    losses = []
    for it in range(60):
        if it == 35:
            for g_ in opt.param_groups:
                g_["lr"] = 0.01
        opt.zero_grad()
        img = render_torch(sc, {key: p}, integrator=integ, seed=it, spp=32)
        loss = ((img - ref) ** 2).mean()
        loss.backward()
        opt.step()
        with torch.no_grad():
            p.clamp_(0.0, 1.0)
        losses.append(float(loss.detach()))
    # the loss sits on a Monte Carlo noise floor (32 spp vs the 256 spp reference): the parameters are the criterion
    assert np.abs(p.detach().numpy() - target).max() < 0.08, (p, target, losses[::8])
    update_params(sc, {key: target})


def test_forward_mode_matches_oracle_fd_and_is_the_transpose_of_backward(built):
    """b200pt_render_forward (RBIntegrator.render_forward, common.py:560-623): (a) equals the oracle's
    forward replay at equal seeds, (b) equals central finite differences of the primal at equal seeds,
    (c) <grad_in, forward(v)> == <backward(grad_in), v> (same samples: exact transpose up to fp32)."""
    from oracle import oracle
    from mitsuba3_b200.integrators import PRBIntegrator, update_params
    tex = (0.3 + 0.4 * np.random.default_rng(1).random((8, 8, 3))).astype(np.float32)
    sc = _prb_scene(rfilter="gaussian", tex=tex)
    integ = PRBIntegrator(max_depth=5)
    P = sc.parameters()
    rng = np.random.default_rng(7)
    tangents = {"tex.reflectance.data": rng.random(tex.shape).astype(np.float32),
                "white.reflectance.value": rng.random(3).astype(np.float32),
                "light.emitter.radiance.value": rng.random(3).astype(np.float32)}
    fwd = integ.render_forward(sc, tangents, seed=4, spp=32)
    o = oracle.OracleScene(sc)
    ref = o.render_forward({P[k]: v for k, v in tangents.items()}, spp=32, seed=4, max_depth=5)
    assert np.isfinite(fwd).all() and np.abs(ref).max() > 0
    assert np.linalg.norm(fwd - ref) / np.linalg.norm(ref) < 2e-3
    # (c) transpose test
    gi = rng.normal(size=sc.film_shape).astype(np.float32)
    g = integ.render_backward(sc, gi, seed=4, spp=32)
    lhs = float((gi.astype(np.float64) * fwd).sum())
    rhs = float(sum((g[k].astype(np.float64) * v).sum() for k, v in tangents.items()))
    assert abs(lhs - rhs) <= 2e-3 * max(abs(lhs), abs(rhs)), (lhs, rhs)
    # (b) finite differences of the primal (prb estimator) along the same direction
    base = {k: sc.textures[P[k]].array().copy() for k in tangents}
    h = 2e-3
    update_params(sc, {k: base[k] + h * tangents[k] for k in tangents}); a = integ.render(sc, seed=4, spp=32)
    update_params(sc, {k: base[k] - h * tangents[k] for k in tangents}); b = integ.render(sc, seed=4, spp=32)
    update_params(sc, base)
    fd = (a.astype(np.float64) - b) / (2 * h)
    assert np.linalg.norm(fwd - fd) / np.linalg.norm(fd) < 2e-2
