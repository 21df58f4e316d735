#!/bin/bash
# One GPU call: the -m gpu suite, bench lines of the traversal variants, the default bench line and the reference arm.
# Results under gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 2>&1 | tail -60 > gpurun_out/gpu_tests.log
tail -15 gpurun_out/gpu_tests.log
b() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 5 --warmup 3 --no-prb --no-cpu-baseline --no-mi-render > gpurun_out/v_$name.json 2>gpurun_out/v_$name.err
      env "$@" timeout 300 python bench.py --workload heightfield205k_1024x1024_64spp_8bounce --steps 3 --warmup 3 --no-prb --no-cpu-baseline --no-mi-render > gpurun_out/v_hf_$name.json 2>>gpurun_out/v_$name.err
      python - $name <<'PY'
import json, sys
for tag in ("", "hf_"):
    try:
        j = json.loads(open(f"gpurun_out/v_{tag}{sys.argv[1]}.json").read().strip().splitlines()[-1]); r = j["roofline"]
        print(f"{tag}{sys.argv[1]:24s} {j['value']:8.1f} Msamples/s {j['ms_per_step']:7.2f} ms  trace {r['avg_launch_ms']*r['launches']/3:7.2f} ms/frame share {r['share_of_step']:.2f}")
    except Exception as e:
        print(tag, sys.argv[1], "failed", e)
PY
}
b old B200PT_TRACE_QUEUE=0
b queue5 B200PT_TRACE_QUEUE=1
b queue4 B200PT_TRACE_QUEUE=1 B200PT_TRACEQ_MINB=4
b queue5_r16 B200PT_TRACE_QUEUE=1 B200PT_REFILL_IDLE=16
b queue4_r4 B200PT_TRACE_QUEUE=1 B200PT_TRACEQ_MINB=4 B200PT_REFILL_IDLE=4
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json | cut -c1-600
