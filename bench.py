#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200 path tracer (contract: see the task
statement; metric and config: BASELINE.json).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, C ABI)
  python bench.py --impl reference --gpus N --steps K ...  reference arm (CPU)

A *step* is one full render of the workload frame: Cornell box 512x512, 256 spp,
path integrator, max_depth 8 (BASELINE.json configs[1]); `value` is Msamples/s
with the scene resident in HBM and the developed image left on the device, `e2e`
is the same metric through the host API (`mitsuba3_b200.render`: parameters
uploaded from host memory, image copied back to the host every step).
N > 1: one process per GPU (torchrun), frame sharded by pixel tiles, ONE NCCL
all-reduce of the raw film per step; weak scaling: spp = 256 * N so that the
per-GPU work is fixed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (width, height, spp per GPU, max_depth, rfilter)
    "cornell_box_512x512_256spp_8bounce": (512, 512, 256, 8, "gaussian"),
    "cornell_box_256x256_64spp_8bounce": (256, 256, 64, 8, "gaussian"),
    # synthetic stand-in for the 200k-triangle config (BASELINE.json configs[4], asset not in the reference tree)
    "heightfield205k_1024x1024_64spp_8bounce": (1024, 1024, 64, 8, "gaussian"),
    # synthetic stand-in for configs[3] (matpreview: principled BSDF + envmap; asset not in the reference tree):
    # 261k-triangle principled sphere + metallic sphere on a checkerboard plane, lit only by a 1024x512 HDR envmap
    "matpreview_like_1024x1024_128spp_8bounce": (1024, 1024, 128, 8, "gaussian"),
}
DEFAULT_WORKLOAD = "cornell_box_512x512_256spp_8bounce"
METRIC = "Msamples/sec (fwd path, Cornell box)"


def build_scene(workload, textured_wall=False):
    import mitsuba3_b200 as mb
    w, h, spp, md, rf = WORKLOADS[workload]
    if workload.startswith("heightfield"):
        d = mb.cornell_box_heightfield(320)
    elif workload.startswith("matpreview"):
        d = mb.matpreview_like()
    else:
        d = mb.cornell_box()
    if textured_wall and "back" in d:
        # BASELINE.json configs[2] / SURVEY 8(d): the back wall's albedo is a 64x64x3 bitmap texture (initial 0.5,
        # bilinear, clamp, raw) -- the parameter the PRB gradient step differentiates
        d["wall-tex"] = {"type": "diffuse", "reflectance": {"type": "bitmap", "data": np.full((64, 64, 3), 0.5, np.float32), "raw": True,
                                                            "filter_type": "bilinear", "wrap_mode": "clamp"}}
        d["back"]["bsdf"] = {"type": "ref", "id": "wall-tex"}
    d["sensor"]["film"].update(width=w, height=h, rfilter={"type": rf})
    d["sensor"]["sampler"]["sample_count"] = spp
    d["integrator"] = {"type": "path", "max_depth": md}
    return mb.load_dict(d), (w, h, spp, md, rf)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_baseline_port(scene, spp_sample):
    """CPU oracle (port of the reference algorithm, OpenMP over pixels) on a bounded sample."""
    from oracle import oracle
    o = oracle.OracleScene(scene)
    H, W, _ = scene.film_shape
    o.render(spp=1, seed=0, mode=0)                      # warm-up (page-in, thread pool)
    t0 = time.perf_counter()
    o.render(spp=spp_sample, seed=0, mode=0)
    dt = time.perf_counter() - t0
    return W * H * spp_sample / dt / 1e6, dt


def cpu_baseline_reference(workload, spp_sample, reps=1):
    """The UNMODIFIED reference (mitsuba scalar_rgb, all host threads) when a snapshot of its
    runtime travels with the repo (oracle/ref_snapshot.sh); None otherwise."""
    from oracle.ref_env import reference_env
    env = reference_env(ROOT)
    if env is None or not workload.startswith("cornell_box"):
        return None
    w, h, spp, md, rf = WORKLOADS[workload]
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "ref_bench.py"), str(w), str(h), str(spp_sample), str(md), rf, str(reps)],
                           env=env, capture_output=True, text=True, timeout=1500)
        return json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        return None


def cpu_baseline(scene, workload, spp1):
    """Reported CPU baseline on a bounded sample (same frame, fewer spp): the reference itself when
    available, else the oracle port."""
    cores = os.cpu_count() or 1
    spp_sample = max(1, min(spp1, int(round(32 * cores / 8))))
    w, h, _, md, _ = WORKLOADS[workload]
    ref = cpu_baseline_reference(workload, spp_sample)
    if ref is not None:
        return {"value": ref["msamples_per_s"], "unit": "Msamples/s", "cores": cores, "kind": "reference", "seconds": ref["seconds"],
                "sample": f"mitsuba {ref['version']} scalar_rgb, {ref['accel']}, same frame {w}x{h}, max_depth {md}, {spp_sample} of {spp1} spp ({ref['seconds']:.1f} s)"}
    spp_sample = max(1, spp_sample // 2)
    v, dt = cpu_baseline_port(scene, spp_sample)
    return {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "seconds": dt,
            "sample": f"CPU oracle (OpenMP), same frame {w}x{h}, max_depth {md}, {spp_sample} of {spp1} spp ({dt:.1f} s)"}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path on the host cores --
    mitsuba scalar_rgb from the runtime snapshot when it travelled with the repo, else the CPU
    oracle that restates it (pinned per pixel to scalar_rgb renders)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    scene, (w, h, spp, md, rf) = build_scene(args.workload)
    vals = []
    base = None
    for _ in range(max(1, args.steps)):
        base = cpu_baseline(scene, args.workload, spp)
        vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": base["seconds"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "parallelism": "host cores", "sample": base["sample"]},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--prb", action="store_true", help="(default on) also time the PRB gradient step (ms/grad-step)")
    ap.add_argument("--no-prb", action="store_true", help="skip the PRB gradient-step timing")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    g.build()
    import mitsuba3_b200 as mb
    from mitsuba3_b200 import dist as mbd
    from mitsuba3_b200.integrators import PathIntegrator, PRBIntegrator, device_scene, update_params

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    if args.gpus != world and rank == 0 and world > 1:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE {world}", file=sys.stderr)
    n_gpus = world

    os.environ.setdefault("B200PT_PROFILE", "1")          # per-launch CUDA events around the traversal kernel
    scene, (w, h, spp1, md, rf) = build_scene(args.workload)
    spp = spp1 * n_gpus if args.scaling == "weak" else spp1
    integ = PathIntegrator(max_depth=md)
    ds = device_scene(scene, local)
    samples_per_step = w * h * spp

    def step_device(seed):
        return mbd.render_distributed(scene, integ, seed=seed, spp=spp, device=local)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for i in range(args.warmup):
        step_device(1000 + i)
    sync_all()
    clocks = ClockSampler(local); clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = bounces = shadow = trace_launches = trace_rays = 0
    trace_ms = 0.0
    t0 = time.perf_counter(); ev0.record()
    for i in range(args.steps):
        step_device(i)              # working set per step (wavefront state) >> L2, see DESIGN.md
        st = ds.stats()
        launches += st["kernel_launches"]; bounces += st["bounces"]; shadow += st["shadow_rays"]
        trace_ms += st["trace_ms"]; trace_launches += st["trace_launches"]; trace_rays += st["trace_rays"]
    ev1.record()
    sync_all()
    wall = time.perf_counter() - t0
    dev_ms = ev0.elapsed_time(ev1)
    clk = clocks.stop()
    t = torch.tensor([dev_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = samples_per_step / (ms_per_step * 1e-3) / 1e6

    # ---- e2e: host API, params uploaded + image copied back every step -------------------
    e2e = None
    if rank == 0 or world > 1:
        names = scene.parameters()
        h2d = sum(scene.textures[i].size * 4 for i in names.values())
        img_bytes = w * h * 3 * 4
        def step_host(seed):
            update_params(scene, {k: scene.textures[i].array() for k, i in names.items()}, local)
            if world == 1:
                return integ.render(scene, seed=seed, spp=spp, device=local)
            return mbd.render_distributed(scene, integ, seed=seed, spp=spp, device=local).cpu().numpy()
        step_host(77)
        sync_all()
        t1 = time.perf_counter()
        for i in range(args.steps):
            img = step_host(i)
        sync_all()
        e2e_s = (time.perf_counter() - t1) / args.steps
        tt = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": samples_per_step / float(tt.item()) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": img_bytes, "checksum": float(np.asarray(img).mean())}

    # ---- PRB gradient step (BASELINE.json: "ms/grad-step (PRB)"): primal + adjoint, device-timed -------
    prb = None
    if not args.no_prb:
        pint = PRBIntegrator(max_depth=md)
        scene_main = scene
        scene, _ = build_scene(args.workload, textured_wall=True)      # configs[2]: wall albedo texture is the parameter
        gi = torch.full((h, w, 3), 1.0 / (h * w * 3), device=f"cuda:{local}")
        spp_g = 64 * n_gpus if args.scaling == "weak" else 64
        n_grad = max(3, args.steps // 2)
        for i in range(2):
            mbd.render_distributed(scene, pint, seed=50 + i, spp=spp_g, device=local)
            mbd.render_backward_distributed(scene, gi, pint, seed=150 + i, spp=spp_g, device=local)
        sync_all()
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for i in range(n_grad):
            mbd.render_distributed(scene, pint, seed=i, spp=spp_g, device=local)
            mbd.render_backward_distributed(scene, gi, pint, seed=100 + i, spp=spp_g, device=local)
        pe1.record()
        sync_all()
        tp = torch.tensor([pe0.elapsed_time(pe1)], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        n_params = int(sum(t.size for t in scene.textures if t.differentiable))
        scene = scene_main
        prb = {"ms_per_grad_step": float(tp.item()) / n_grad, "spp": spp_g, "grad_steps": n_grad, "differentiated_floats": n_params,
               "what": "wall albedo = 64x64x3 bitmap texture; primal render + render_backward (PRB adjoint, atomicAdd gradient scatter%s), CUDA events, "
                       "max over ranks, max_depth %d" % (" + gradient all-reduce" if world > 1 else "", md),
               "reference": "unmeasurable here: prb needs an AD variant (llvm_ad_rgb) and the image has no libLLVM"}

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
        b_bar = bounces / max(1, (samples_per_step // n_gpus) * args.steps)   # rank 0's lanes
        # dominant kernel: k_trace. Algorithmic bytes per ray (DESIGN.md "Roofline"): closest-hit ray 28 B read +
        # 20 B hit record written; shadow ray 40 B record read + 24 B result read-modify-write.
        closest = trace_rays - shadow
        trace_bytes = closest * 48.0 + shadow * 64.0
        trace_avg_ms = trace_ms / max(1, trace_launches)
        trace_gbs = trace_bytes / max(trace_ms * 1e-3, 1e-12) / 1e9 if trace_ms > 0 else None
        # DRAM traffic of the dominant kernel per launch: dram__bytes_{read,write}.sum per lane from the committed
        # `ncu --set full` capture (profiles/r01_trace_traffic.json) x the lanes one launch of this run processed
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_trace_traffic.json")))
            if "dram_bytes_per_ray" in tj:      # ncu dram__bytes_{read,write}.sum summed over ALL traversal launches of a frame / rays traced
                traffic = tj["dram_bytes_per_ray"] * trace_rays / max(1, trace_launches)
            else:
                traffic = tj["dram_bytes_per_lane"] * (trace_rays / 2.0) / max(1, trace_launches)   # ~2 rays (shadow + closest) per lane
        except Exception:
            pass
        step_bytes = (144.0 + 304.0 * b_bar) * (samples_per_step // n_gpus)
        step_gbs = step_bytes / (ms_per_step * 1e-3) / 1e9
        out = {
            "metric": METRIC if args.workload.startswith("cornell") else "Msamples/sec (fwd path, %s)" % args.workload, "value": value, "unit": "Msamples/s", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "spp_total": spp, "global_samples_per_step": samples_per_step,
                       "parallelism": f"pixel-tile x{n_gpus}, 1 film all-reduce" if n_gpus > 1 else "single GPU",
                       "l2": "wavefront state per chunk (~400 MB) > L2 (126 MB); film and scene are L2-resident by design",
                       "mean_bounces_per_sample": b_bar},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "wall_ms_per_step": wall / args.steps * 1e3,
            "roofline": {"kernel": "k_trace (BVH traversal: NEE shadow ray + closest hit)", "bound": "hbm",
                         "achieved": trace_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": (trace_gbs / hbm_peak) if trace_gbs else None, "traffic": traffic,
                         "algorithmic_bytes_per_launch": trace_bytes / max(1, trace_launches), "peak_source": peak_src,
                         "avg_launch_ms": trace_avg_ms, "launches": int(trace_launches), "rays_per_launch": trace_rays / max(1, trace_launches),
                         "share_of_step": trace_ms / max(dev_ms, 1e-9),
                         "step": {"bytes_per_sample": 144.0 + 304.0 * b_bar, "achieved": step_gbs, "frac": step_gbs / hbm_peak}},
        }
        switches = sorted(k for k in ("B200PT_WAVE_ORDER", "B200PT_CELL_ORDER", "B200PT_TRACE_PHASES", "B200PT_BVH_WIDE", "B200PT_SPLAT_FOLD") if os.environ.get(k, "0") not in ("", "0"))
        if switches:     # staged traversal variants (off by default): a line measured with one of them says so
            out["config"]["switches"] = switches
        if prb:
            out["prb"] = prb
        if not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(scene, args.workload, spp1)
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
