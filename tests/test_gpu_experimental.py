"""Opt-in GPU tests of the traversal variants staged behind environment switches (all off by default;
profiles/r01_simt_model.md section 3): B200PT_WAVE_ORDER, B200PT_CELL_ORDER, B200PT_TRACE_PHASES,
B200PT_BVH_WIDE, and of the folding reduction of the gaussian splat (B200PT_SPLAT_FOLD, section 5).

They are skipped unless B200PT_TEST_EXPERIMENTAL=1:

    B200PT_TEST_EXPERIMENTAL=1 python -m pytest tests/test_gpu_experimental.py -m gpu -x -q

The switches only change the order in which rays are traced and shaded and which tree is walked; every lane
carries its own RNG stream and film position and the triangle test and its tie-break are shared, so a render
with a switch on must reproduce the default render sample for sample. The film is accumulated with fp32
atomics whose order may differ, hence "equal within a few ulps of the accumulated sum" instead of bit-equal.
"""
import itertools
import os

import numpy as np
import pytest

from conftest import cbox, env_scene, materials_cbox

import mitsuba3_b200 as mb

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("B200PT_TEST_EXPERIMENTAL") != "1",
                                 reason="staged traversal variants: set B200PT_TEST_EXPERIMENTAL=1 to run")]

SWITCHES = ["B200PT_WAVE_ORDER", "B200PT_CELL_ORDER", "B200PT_TRACE_PHASES", "B200PT_BVH_WIDE", "B200PT_SPLAT_FOLD"]


def _heightfield(res=48, spp=8):
    d = mb.cornell_box_heightfield(96)       # 18k triangles: the tree does not fit the shared-memory window
    d["sensor"]["film"].update(width=res, height=res, rfilter={"type": "box"})
    d["sensor"]["sampler"]["sample_count"] = spp
    return d


SCENES = {
    "cbox_box": lambda: cbox(48, "box", 32, 8),
    "cbox_gauss": lambda: cbox(48, "gaussian", 16, 8),
    "materials": lambda: materials_cbox(48, "box", 32, 8),
    "envmap": lambda: env_scene("envmap", res=48, spp=16, max_depth=6, area_light=True),
    "heightfield": _heightfield,
}


def _render(desc, env):
    """The switches are read when the device scene is created (B200PT_SPLAT_FOLD is process-wide: it is set or
    cleared explicitly by every call here)."""
    old = {k: os.environ.get(k) for k in SWITCHES}
    try:
        for k in SWITCHES:
            os.environ.pop(k, None)
        os.environ.update(env)
        os.environ.setdefault("B200PT_SPLAT_FOLD", "0")
        sc = mb.load_dict(desc)
        return mb.render(sc, seed=3), None
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _assert_same(img, ref):
    assert np.isfinite(img).all()
    err = np.abs(img.astype(np.float64) - ref.astype(np.float64))
    tol = 4e-6 * np.maximum(np.abs(ref), 1e-3)      # fp32 atomic accumulation order
    assert (err <= tol).all(), f"max abs err {err.max():.3e}, {(err > tol).mean():.4f} of the values off"


@pytest.mark.parametrize("scene", sorted(SCENES))
@pytest.mark.parametrize("switch", SWITCHES)
def test_single_switch_reproduces_the_default_render(built, scene, switch):
    desc = SCENES[scene]()
    ref, _ = _render(desc, {})
    img, _ = _render(desc, {switch: "1"})
    _assert_same(img, ref)


@pytest.mark.parametrize("combo", [c for n in (2, 3, 4) for c in itertools.combinations(SWITCHES, n)], ids="+".join)
def test_switch_combinations(built, combo):
    desc = cbox(48, "box", 32, 8)
    ref, _ = _render(desc, {})
    img, _ = _render(desc, {k: "1" for k in combo})
    _assert_same(img, ref)


@pytest.mark.parametrize("switch", SWITCHES)
def test_prb_gradient_with_switch(built, switch):
    """The adjoint replay goes through the same traversal launches (no shadow rays there: phase 1 is all empty jobs)."""
    desc = cbox(24, "box", 16, 6)
    desc["integrator"]["type"] = "prb"

    def grad(env):
        old = {k: os.environ.get(k) for k in SWITCHES}
        try:
            for k in SWITCHES:
                os.environ.pop(k, None)
            os.environ.update(env)
            os.environ.setdefault("B200PT_SPLAT_FOLD", "0")
            sc = mb.load_dict(desc)
            from mitsuba3_b200.integrators import make_integrator
            integ = make_integrator(sc)
            img = integ.render(sc, seed=1)
            g = integ.render_backward(sc, np.ones_like(img), seed=1)
            return img, g
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v

    img0, g0 = grad({})
    img1, g1 = grad({switch: "1"})
    _assert_same(img1, img0)
    for name in g0:
        a, b = np.asarray(g0[name], np.float64), np.asarray(g1[name], np.float64)
        assert np.allclose(a, b, rtol=2e-4, atol=1e-5 * max(1.0, np.abs(a).max())), name
