#!/bin/bash
# First GPU call of the next round (profiles/r01_simt_model.md section 4): parity and timing of the staged
# traversal switches, one by one and together. Run on the GPU box from the repository root:
#
#   gpurun --timeout 1500 -- 'bash tools/run_switch_matrix.sh'
#
# Results land in gpurun_out/switch_*.json (bench lines) and gpurun_out/switch_tests.log.
set -u
mkdir -p gpurun_out
LOG=gpurun_out/switch_tests.log
: > "$LOG"

echo "== switch on == switch off, sample for sample" | tee -a "$LOG"
B200PT_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_experimental.py -m gpu -q 2>&1 | tail -15 | tee -a "$LOG"

bench() {   # name, env assignments...
    local name=$1; shift
    env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-prb --no-cpu-baseline > "gpurun_out/switch_${name}.json" 2>> "$LOG"
    env "$@" timeout 300 python bench.py --workload heightfield205k_1024x1024_64spp_8bounce --steps 3 --warmup 3 --no-prb --no-cpu-baseline \
        > "gpurun_out/switch_hf_${name}.json" 2>> "$LOG"
    python - "$name" <<'PY' | tee -a "$LOG"
import json, sys
name = sys.argv[1]
for tag in ("", "hf_"):
    try:
        j = json.loads(open(f"gpurun_out/switch_{tag}{name}.json").read().strip().splitlines()[-1])
        r = j["roofline"]
        print(f"{tag}{name:28s} {j['value']:8.1f} Msamples/s  {j['ms_per_step']:7.2f} ms/step  trace share {r['share_of_step']:.2f}  launches {j['gpu_launches']}")
    except Exception as e:
        print(f"{tag}{name}: no bench line ({e})")
PY
}

bench default B200PT_NONE=0
bench wave_order B200PT_WAVE_ORDER=1
bench cell_order B200PT_CELL_ORDER=1
bench phases B200PT_TRACE_PHASES=1
bench wide B200PT_BVH_WIDE=1
bench splat_fold B200PT_SPLAT_FOLD=1
bench wave_order+phases B200PT_WAVE_ORDER=1 B200PT_TRACE_PHASES=1
bench wave_order+phases+wide B200PT_WAVE_ORDER=1 B200PT_TRACE_PHASES=1 B200PT_BVH_WIDE=1
bench all B200PT_WAVE_ORDER=1 B200PT_CELL_ORDER=1 B200PT_TRACE_PHASES=1 B200PT_BVH_WIDE=1 B200PT_SPLAT_FOLD=1
