#!/usr/bin/env python3
"""Turn the captures of tools/ncu_capture.sh (gpurun_out/<TAG>_*) into the tracked summaries under profiles/:
  profiles/<TAG>_launches.csv, <TAG>_trace_traffic.csv   copies of the launch list and the per-launch DRAM traffic
  profiles/<TAG>_ncu.md                                  per-kernel share of the frame, key metrics of the traversal and the
                                                         diffuse shading kernel, instruction / thread accounting of the traversal
                                                         kernel by code region (source page)
  profiles/r02_trace_traffic.json                        DRAM bytes per ray of the traversal kernel (bench.py: roofline.traffic)
Usage: python tools/ncu_summarise.py r02k"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02k"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
out = [f"# ncu summary `{tag}` (tools/ncu_capture.sh, 1 x B200, `--clock-control none`; per-launch times under ncu are cold-cache and serialised)\n"]


def num(x):
    return float(x.replace(",", ""))


# ---- launch list ---------------------------------------------------------------------------------------------------------------
f = os.path.join(G, f"{tag}_launches.csv")
if os.path.exists(f):
    shutil.copy(f, os.path.join(P, f"{tag}_launches.csv"))
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        k = r[ki].split("(")[0].replace("void ", "")
        if not k.startswith("pt::"):
            continue
        agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += num(r[vi]) / 1e6
    tot = sum(v[1] for v in agg.values())
    out.append("## Launch list (`gpu__time_duration.sum`), all frames of the capture\n\n| kernel | launches | ms | share |\n|---|---|---|---|")
    for k, v in agg.items():
        out.append(f"| `{k}` | {v[0]} | {v[1]:.3f} | {v[1] / tot * 100:.1f} % |")
    out.append("")

# ---- DRAM traffic of the traversal launches --------------------------------------------------------------------------------------
f = os.path.join(G, f"{tag}_trace_traffic.csv")
if os.path.exists(f):
    shutil.copy(f, os.path.join(P, f"{tag}_trace_traffic.csv"))
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    h = rows[0]; mi, vi, ui, ii = h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        v = num(r[vi]); u = r[ui].lower()
        if "byte" in u:
            v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
        per.setdefault(r[ii], {})[r[mi]] = v
    tot_b = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in per.values())
    # rays of the same frame: bench stats of the capture run (last line of the log is the bench JSON)
    rays = None
    try:
        j = json.loads(open(os.path.join(G, f"{tag}_ncu_d.log")).read().strip().splitlines()[-1])
        rays = j["roofline"]["rays_per_launch"] * j["roofline"]["launches"]
    except Exception:
        pass
    out.append(f"## DRAM traffic of the {len(per)} traversal launches of one frame\n")
    out.append(f"`dram__bytes_read.sum + dram__bytes_write.sum` = {tot_b / 1e9:.2f} GB" + (f" for {rays / 1e6:.0f} M rays = **{tot_b / rays:.1f} B/ray** (algorithmic: 48 B per closest-hit ray, 64 B per shadow ray)" if rays else ""))
    out.append("")
    if rays:
        json.dump({"dram_bytes_per_ray": tot_b / rays, "dram_bytes_frame": tot_b, "rays_frame": rays, "source": f"profiles/{tag}_trace_traffic.csv",
                   "kernel": "k_trace_flat"}, open(os.path.join(P, "r02_trace_traffic.json"), "w"))

# ---- full captures -----------------------------------------------------------------------------------------------------------------
WANT = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for name in ("trace", "shade"):
    rep = os.path.join(G, f"{tag}_{name}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    kn = v[h.index("Kernel Name")] if "Kernel Name" in h else name
    out.append(f"## `ncu --set full` of one launch: `{kn[:90]}`\n\n| metric | value |\n|---|---|")
    for i, n in enumerate(h):
        if n in WANT:
            out.append(f"| `{n}` | {v[i]} {u[i]} |")
    out.append("")
    if name == "trace":
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(src.splitlines()))
        h = rows[1]; ai, ii, ti = h.index("Address"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
        data = []
        for r in rows[2:]:
            try:
                data.append((int(r[ai], 16) if r[ai].startswith("0x") else int(r[ai]), int(r[ii]), int(r[ti])))
            except Exception:
                pass
        tot = sum(d[1] for d in data)
        seg, cur = [], None
        for d in data:
            if cur is None or abs(d[1] - cur[0]) > 0.15 * max(d[1], cur[0], 1):
                cur = [d[1], d[0], d[0], 0, 0, 0]; seg.append(cur)
            cur[2] = d[0]; cur[3] += d[1]; cur[4] += d[2]; cur[5] += 1
        out.append("Code regions of the kernel (consecutive SASS instructions with the same execution count; source page):\n\n| SASS range | instructions | executions each | share of warp instructions | threads / instruction |\n|---|---|---|---|---|")
        for s in seg:
            if s[3] > 0.006 * tot:
                out.append(f"| {s[1] & 0xffff:04x}-{s[2] & 0xffff:04x} | {s[5]} | {s[0]} | {s[3] / tot * 100:.1f} % | {s[4] / max(s[3], 1):.1f} |")
        out.append(f"\ntotal warp instructions {tot}\n")
f = os.path.join(G, f"{tag}_shade.csv")
if os.path.exists(f):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    h = rows[0]; mi, vi, ui, ki = h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit"), h.index("Kernel Name")
    out.append(f"## Key metrics of one launch of the diffuse shading kernel: `{rows[1][ki][:80]}`\n\n| metric | value |\n|---|---|")
    for r in rows[1:]:
        out.append(f"| `{r[mi]}` | {r[vi]} {r[ui]} |")
    out.append("")
open(os.path.join(P, f"{tag}_ncu.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out)[:3000])
