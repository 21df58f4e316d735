#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${TAG:-r02g}
timeout 500 python -m pytest tests -m gpu -q --timeout 200 --tb=short 2>&1 | tail -40 > gpurun_out/gpu_tests_${TAG}.log; tail -5 gpurun_out/gpu_tests_${TAG}.log
B="python bench.py --steps 5 --warmup 3 --no-prb --no-cpu-baseline --no-mi-render"
run() {  # name workload env...
    local name=$1 wl=$2; shift 2
    env "$@" timeout 200 $B --workload $wl > gpurun_out/${TAG}_${name}.json 2> gpurun_out/${TAG}_${name}.err
    python - "$name" "$TAG" <<'PY'
import json, sys
try:
    j = json.loads(open(f"gpurun_out/{sys.argv[2]}_{sys.argv[1]}.json").read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:24s} {j['value']:9.1f} Msamples/s {j['ms_per_step']:8.2f} ms/step  trace share {j['roofline']['share_of_step']:.2f} avg launch {j['roofline']['avg_launch_ms']:.3f} ms checksum {j['e2e']['checksum']:.6f}")
except Exception as e:
    print(sys.argv[1], "failed:", e)
PY
}
C=cornell_box_512x512_256spp_8bounce
run c_flat4 $C B200PT_FLAT_BLOCKS_PER_SM=4


