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


@pytest.mark.parametrize("world", [1, 2, 3, 4, 5, 8])
@pytest.mark.parametrize("size", [(512, 512), (1024, 1024), (200, 136)])
def test_tile_deal_partitions_the_film_evenly(world, size):
    """Pixel-tile sharding: every pixel has one owner, the ranks get (nearly) the same number of
    pixels, and -- the reason for the diagonal deal -- the tiles of a rank are spread over the image: it owns
    tiles in at least half of the tile columns and of the tile rows. (The plain `t % N` deal gave a rank only
    tiles_x / N columns whenever N divided the tiles per row: 2 of 16 at 8 GPUs on the 512 x 512 Cornell box,
    16 % load imbalance.)"""
    from mitsuba3_b200.dist import tile_owner
    W, H = size
    ts = 32
    ys, xs = np.mgrid[0:H, 0:W]
    owner = tile_owner(xs, ys, W, ts, world)
    assert owner.min() >= 0 and owner.max() < world
    counts = np.bincount(owner.ravel(), minlength=world)
    assert counts.sum() == W * H
    tiles_x, tiles_y = -(-W // ts), -(-H // ts)
    if W >= 512:
        assert counts.max() <= counts.mean() * 1.02
    t_owner = owner[::ts, ::ts]                                    # one entry per tile
    if tiles_x * tiles_y >= 16 * world:
        for r in range(world):
            ty, tx = np.nonzero(t_owner == r)
            assert len(set(tx.tolist())) >= tiles_x // 2 and len(set(ty.tolist())) >= tiles_y // 2
