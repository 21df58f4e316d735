"""The integrator plugin against a live, unmodified Mitsuba 3 (optional: needs the runtime
snapshot made by oracle/ref_snapshot.sh; skipped when it is absent)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT
from oracle.ref_env import reference_env


def _run(mode):
    env = reference_env(ROOT)
    if env is None:
        pytest.skip("no snapshot of the reference runtime (oracle/ref_snapshot.sh)")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "_mitsuba_live.py"), mode], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


def test_extraction_from_live_mitsuba(built):
    assert "LIVE_CPU_OK" in _run("cpu")


@pytest.mark.gpu
def test_mi_render_through_registered_plugin(built):
    assert "LIVE_GPU_OK" in _run("gpu")
