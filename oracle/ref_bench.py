"""Times the UNMODIFIED reference on a bounded sample of a bench workload. Run in a subprocess with the environment of
oracle.ref_env (python oracle/run_ref.py oracle/ref_bench.py ...); prints one JSON line. Used by bench.py
(cpu_baseline kind "reference", --impl reference). Test / measurement infrastructure, never imported by the product.

    ref_bench.py <variant> <workload> <w> <h> <spp> <max_depth> <rfilter> <reps> [prb_spp]

variant: llvm_ad_rgb (the variant BASELINE.json names; needs oracle/llvm_shim in an image without libLLVM) or
scalar_rgb. With prb_spp > 0 and an AD variant the PRB gradient step (primal render + dr.backward through
mi.render, prb integrator, wall-albedo bitmap texture as the parameter) is timed as well.
"""
import json
import os
import sys
import time

variant, workload = sys.argv[1], sys.argv[2]
w, h, spp, max_depth, rfilter, reps = int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), sys.argv[7], int(sys.argv[8])
prb_spp = int(sys.argv[9]) if len(sys.argv) > 9 else 0

import mitsuba as mi

try:
    mi.set_variant(variant)
    import drjit as dr
    if variant.startswith("llvm"):
        dr.eval(dr.arange(mi.Float, 4) + 1)        # forces the LLVM backend to initialise (raises without libLLVM)
except Exception as e:  # noqa: BLE001
    print(json.dumps({"error": f"{variant}: {type(e).__name__}: {str(e)[:200]}"}))
    sys.exit(0)
import numpy as np

jit = variant.startswith("llvm")


def scene_dict(integrator, textured=False):
    if True:
        d = mi.cornell_box()
        if textured:     # BASELINE.json configs[2]: back wall albedo = 64x64x3 bilinear bitmap
            d["wall-tex"] = {"type": "diffuse", "reflectance": {"type": "bitmap", "bitmap": mi.Bitmap(np.full((64, 64, 3), 0.5, np.float32)),
                                                                "raw": True, "filter_type": "bilinear", "wrap_mode": "clamp"}}
            d["back"]["bsdf"] = {"type": "ref", "id": "wall-tex"}
    d["sensor"]["film"].update(width=w, height=h, rfilter={"type": rfilter})
    d["integrator"] = {"type": integrator, "max_depth": max_depth}
    return d


def sync(x):
    if jit:
        dr.eval(x); dr.sync_thread()


if workload.startswith("matpreview"):      # the reference's own asset, rebuilt from the committed fixture arrays
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from ref_matpreview import load_matpreview
    scene = load_matpreview(mi, w, h, spp, max_depth)
else:
    scene = mi.load_dict(scene_dict("path"))
sync(mi.render(scene, spp=1))                    # warm-up: thread pool, page-in, (JIT) kernel compilation at this launch size
if jit:
    sync(mi.render(scene, spp=spp, seed=99))     # the timed launch size compiles / caches its kernel here
times = []
mean = 0.0
for i in range(reps):
    t0 = time.perf_counter()
    img = mi.render(scene, spp=spp, seed=i)
    sync(img)
    times.append(time.perf_counter() - t0)
    mean = float(np.array(img).mean())
t = sum(times) / len(times)
out = {"msamples_per_s": w * h * spp / t / 1e6, "seconds": t, "threads": dr.thread_count() if hasattr(dr, "thread_count") else None,
       "mean": mean, "version": mi.__version__, "variant": variant,
       "accel": "Embree" if getattr(mi, "MI_ENABLE_EMBREE", False) else "kd-tree (MI_ENABLE_EMBREE=OFF build)"}

if prb_spp > 0 and jit and not workload.startswith("matpreview"):
    try:
        scene = mi.load_dict(scene_dict("prb", textured=True))
        params = mi.traverse(scene)
        key = next(k for k in params.keys() if k.endswith("reflectance.data"))
        dr.enable_grad(params[key]); params.update()

        def grad_step(seed):
            img = mi.render(scene, params, spp=prb_spp, seed=seed)
            dr.backward(dr.mean(img))
            g = dr.grad(params[key]); sync(g)
            dr.set_grad(params[key], 0)
            return g
        grad_step(50); grad_step(51)
        ts = []
        for i in range(max(2, reps)):
            t0 = time.perf_counter(); grad_step(i); ts.append(time.perf_counter() - t0)
        out["prb_ms_per_grad_step"] = 1e3 * sum(ts) / len(ts)
        out["prb_spp"] = prb_spp
    except Exception as e:  # noqa: BLE001
        out["prb_error"] = f"{type(e).__name__}: {str(e)[:200]}"
print(json.dumps(out))
