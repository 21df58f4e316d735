"""Host-side logic: scene parsing, property validation, pixel-tile sharding,
2-process gloo film reduction (CPU)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, cbox, materials_cbox

import mitsuba3_b200 as mb
from mitsuba3_b200 import abi


def test_parser_rejects_out_of_scope_plugins():
    d = mb.cornell_box(); d["sphere"] = {"type": "sphere"}
    with pytest.raises(NotImplementedError):
        mb.load_dict(d)
    d = mb.cornell_box(); d["integrator"] = {"type": "volpath"}
    with pytest.raises(NotImplementedError):
        mb.load_dict(d)


def test_integrator_property_validation():
    from mitsuba3_b200.integrators import PathIntegrator, PRBIntegrator
    with pytest.raises(RuntimeError, match="max_depth"):
        PathIntegrator(max_depth=-2)
    with pytest.raises(RuntimeError, match="rr_depth"):
        PathIntegrator(rr_depth=0)
    with pytest.raises(RuntimeError, match="unreferenced"):
        PathIntegrator(bogus=1)
    assert PRBIntegrator().max_depth == 6 and PathIntegrator().max_depth == -1


def test_materials_scene_parses():
    sc = mb.load_dict(materials_cbox())
    types = sorted(b.type for b in sc.bsdfs)
    assert abi.BSDF_CONDUCTOR in types and abi.BSDF_DIELECTRIC in types and abi.BSDF_PRINCIPLED in types
    assert any(b.twosided for b in sc.bsdfs)
    desc, keep = sc.build_desc()
    assert desc.n_shapes == 8 and desc.n_emitters == 1
    glass = [b for b in sc.bsdfs if b.type == abi.BSDF_DIELECTRIC][0]
    assert abs(glass.eta - 1.5046 / 1.000277) < 1e-6


def test_mesh_shape_from_arrays():
    d = cbox()
    d["tri"] = {"type": "mesh", "positions": [[0, 0, 0], [1, 0, 0], [0, 1, 0]], "faces": [[0, 1, 2]], "bsdf": {"type": "ref", "id": "white"}}
    sc = mb.load_dict(d)
    assert sc.n_triangles == 37


def test_gloo_two_rank_film_reduce():
    """World-size-2 film all-reduce + develop on CPU (gloo): the sharding / reduction
    plumbing of mitsuba3_b200.dist with a stand-in per-rank renderer."""
    script = os.path.join(ROOT, "tests", "_gloo_worker.py")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29541", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", script],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "GLOO_OK" in r.stdout
