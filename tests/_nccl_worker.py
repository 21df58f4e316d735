"""N-rank NCCL check (one process per GPU): the sharded render + film all-reduce equals the
single-GPU full-frame render; same for the PRB gradients."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mitsuba3_b200 as mb
from mitsuba3_b200 import dist as mbd
from mitsuba3_b200.integrators import PathIntegrator, PRBIntegrator

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
rank, ws = mbd.world()
for rf in ("box", "gaussian"):
    d = mb.cornell_box(); d["sensor"]["film"].update(width=96, height=64, rfilter={"type": rf})
    sc = mb.load_dict(d)
    integ = PathIntegrator(max_depth=6)
    img = mbd.render_distributed(sc, integ, seed=3, spp=16, tile_size=16, device=local).cpu().numpy()
    ref = integ.render(sc, seed=3, spp=16, device=local)          # whole frame on this GPU
    if rf == "box":
        assert np.array_equal(img, ref), np.abs(img - ref).max()   # disjoint pixel ownership + deterministic sums
    else:
        assert np.allclose(img, ref, rtol=2e-5, atol=1e-6), np.abs(img - ref).max()
    pint = PRBIntegrator(max_depth=4)
    gi = np.random.default_rng(0).random(sc.film_shape).astype(np.float32)
    g = mbd.render_backward_distributed(sc, gi, pint, seed=5, spp=8, tile_size=16, device=local)
    g1 = pint.render_backward(sc, gi, seed=5, spp=8, device=local)
    for k in g1:
        assert np.allclose(g[k], g1[k], rtol=1e-3, atol=1e-6), (k, g[k], g1[k])
dist.barrier()
if rank == 0:
    print("NCCL_OK", ws)
dist.destroy_process_group()
