"""Where does a PRB gradient step spend its time? Host wall clock with a device synchronisation around each phase."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mitsuba3_b200 import dist as mbd
from mitsuba3_b200.integrators import PRBIntegrator, device_scene
scene, (w, h, spp1, md, rf) = bench.build_scene(bench.DEFAULT_WORKLOAD, textured_wall=True)
pint = PRBIntegrator(max_depth=md)
gi = torch.full((h, w, 3), 1.0 / (h * w * 3), device="cuda:0")
def T(f):
    torch.cuda.synchronize(); t = time.perf_counter(); r = f(); torch.cuda.synchronize(); return (time.perf_counter() - t) * 1e3, r
for it in range(4):
    a, _ = T(lambda: mbd.render_distributed(scene, pint, seed=it, spp=64, device=0))
    st1 = device_scene(scene, 0).stats()
    b, _ = T(lambda: mbd.render_backward_distributed(scene, gi, pint, seed=100 + it, spp=64, device=0))
    st2 = device_scene(scene, 0).stats()
    print(f"iter {it}: primal render {a:.2f} ms (device {st1['device_ms']:.2f}, launches {st1['kernel_launches']}), backward {b:.2f} ms (device {st2['device_ms']:.2f}, launches {st2['kernel_launches']})")
os.environ["B200PT_INLINE_VISIBILITY"] = "1"
for it in range(2):
    b, _ = T(lambda: mbd.render_backward_distributed(scene, gi, pint, seed=100 + it, spp=64, device=0))
    print(f"inline visibility: backward {b:.2f} ms (device {device_scene(scene, 0).stats()['device_ms']:.2f})")
