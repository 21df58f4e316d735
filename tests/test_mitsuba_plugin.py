"""The integrator plugin against a live, unmodified Mitsuba 3 (optional: needs the runtime
snapshot made by oracle/ref_snapshot.sh; skipped when it is absent)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle.ref_env import reference_env


def _run(mode, script="_mitsuba_live.py"):
    env = reference_env(ROOT)
    if env is None:
        pytest.skip("no snapshot of the reference runtime (oracle/build_ref.sh)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", script), mode], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_extraction_from_live_mitsuba(built):
    assert "LIVE_CPU_OK" in _run("cpu")


@pytest.mark.gpu
def test_mi_render_through_registered_plugin(built):
    assert "LIVE_GPU_OK" in _run("gpu")


@pytest.mark.gpu
def test_dr_backward_through_the_registered_prb_plugin(built):
    """mi.render(scene, params) + dr.backward + mi.ad.Adam with `b200_prb` under the reference's llvm_ad_rgb variant
    (util.py:344-395 _RenderOp -> render_backward -> dr.accum_grad): gradients equal the reference's own `prb` at equal
    seeds. Needs Dr.Jit's LLVM backend (oracle/llvm_shim in an image without libLLVM)."""
    if "DRJIT_LIBLLVM_PATH" not in (reference_env(ROOT) or {}):
        pytest.skip("no LLVM for the reference's AD variant")
    assert "LIVE_AD_OK" in _run("ad", "_mitsuba_live_ad.py")
