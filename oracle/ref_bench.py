"""Times the UNMODIFIED reference (mitsuba scalar_rgb, all host threads) on a bounded sample of
a bench workload. Run in a subprocess with the environment of oracle.ref_env; prints one
JSON line. Used by bench.py (cpu_baseline kind "reference", --impl reference)."""
import json
import sys
import time

import mitsuba as mi

mi.set_variant("scalar_rgb")
import drjit as dr

w, h, spp, max_depth, rfilter = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 1
d = mi.cornell_box()
d["sensor"]["film"].update(width=w, height=h, rfilter={"type": rfilter})
d["integrator"] = {"type": "path", "max_depth": max_depth}
scene = mi.load_dict(d)
mi.render(scene, spp=1)                    # warm-up: thread pool, page-in
times = []
for i in range(reps):
    t0 = time.perf_counter()
    img = mi.render(scene, spp=spp, seed=i)
    times.append(time.perf_counter() - t0)
t = sum(times) / len(times)
print(json.dumps({"msamples_per_s": w * h * spp / t / 1e6, "seconds": t, "threads": mi.Thread.thread_count() if hasattr(mi.Thread, "thread_count") else dr.thread_count(),
                  "mean": float(dr.mean(img.array)) if hasattr(img, "array") else 0.0, "version": mi.__version__, "accel": "kd-tree (MI_ENABLE_EMBREE=OFF build)"}))
