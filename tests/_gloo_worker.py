"""World-size-2 CPU (gloo) check of the sharding + film-reduction plumbing.
Each rank renders ITS pixel tiles with the CPU oracle standing in for the per-rank
renderer (test infrastructure), reduces with mitsuba3_b200.dist and compares with a
single-process full-frame film."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mitsuba3_b200 as mb
from mitsuba3_b200 import dist as mbd
from oracle import oracle

dist.init_process_group("gloo")
rank, ws = mbd.world()
assert ws == 2
d = mb.cornell_box(); d["sensor"]["film"].update(width=40, height=24, rfilter={"type": "gaussian"})
sc = mb.load_dict(d)
o = oracle.OracleScene(sc)
_, film = o.render(spp=4, seed=1, return_film=True, shard_rank=rank, shard_count=ws, tile_size=8)
# every pixel of this rank's film that received a box-centre sample belongs to it
t = torch.from_numpy(film.copy())
mbd.all_reduce_film(t)
img = mbd.develop(t).numpy()
ref_img, ref_film = o.render(spp=4, seed=1, return_film=True)
assert np.allclose(t.numpy(), ref_film, rtol=1e-5, atol=1e-6), np.abs(t.numpy() - ref_film).max()
assert np.allclose(img, ref_img, rtol=1e-5, atol=1e-6)
own = np.array([[mbd.tile_owner(x, y, 40, 8, ws) for x in range(40)] for y in range(24)])
assert (own == rank).sum() + (own == 1 - rank).sum() == 40 * 24 and 0 < (own == rank).sum() < 40 * 24
dist.barrier()
if rank == 0:
    print("GLOO_OK")
dist.destroy_process_group()
