#!/usr/bin/env python3
"""End-to-end leg of bench.py through a LIVE, unmodified Mitsuba: `mi.render(scene, spp=...)` on a scene whose integrator
is the registered `b200_path` plugin (mitsuba3_b200/mitsuba_plugin.py -> libb200pt.so). Runs in the environment of the
reference runtime (oracle/_ref, see oracle/run_ref.py); prints one JSON line.

    bench_mi_render.py <w> <h> <spp> <max_depth> <rfilter> <steps> <device>

Per step: mi.traverse-cached parameter check, b200pt_render (upload of changed parameters, render, image back to the
host), mi.TensorXf of the image -- what a user of mi.render gets. Wall clock around the calls.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
w, h, spp, md, rf, steps, device = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5], int(sys.argv[6]), int(sys.argv[7])
os.environ.setdefault("CUDA_VISIBLE_DEVICES", str(device))

import mitsuba as mi

variant = "scalar_rgb"
mi.set_variant(variant)
from mitsuba3_b200 import mitsuba_plugin as plug

plug.register(mi)
d = mi.cornell_box()
d["sensor"]["film"].update(width=w, height=h, rfilter={"type": rf})
d["integrator"] = {"type": "b200_path", "max_depth": md}
scene = mi.load_dict(d)
for i in range(2):
    img = mi.render(scene, spp=spp, seed=100 + i)
t0 = time.perf_counter()
for i in range(steps):
    img = mi.render(scene, spp=spp, seed=i)
    a = np.array(img)
dt = (time.perf_counter() - t0) / steps
print("\n" + json.dumps({"value": w * h * spp / dt / 1e6, "unit": "Msamples/s", "ms_per_step": dt * 1e3, "steps": steps, "checksum": float(a.mean()),
                  "api": f"mi.render(scene, spp) of mitsuba {mi.__version__} ({variant}) with the registered b200_path integrator", "d2h_bytes_per_step": w * h * 12}))
