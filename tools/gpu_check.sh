#!/bin/bash
# One GPU call: the -m gpu suite, the PRB breakdown, the default bench line (+ launch list under ncu) and the reference arm.
set -u
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --maxfail=40 ${PYTEST_ARGS:-} 2>&1 | tail -60 > gpurun_out/gpu_tests.log
tail -8 gpurun_out/gpu_tests.log
timeout 300 python tools/prb_breakdown.py 2>&1 | tail -8 | tee gpurun_out/prb_breakdown.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -2 gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
if [ -n "${NCU:-}" ]; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-prb --no-cpu-baseline --no-mi-render > gpurun_out/ncu_b.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace_dyn -s 3 -c 1 -o gpurun_out/r02_trace_dyn python bench.py --steps 1 --warmup 1 --no-prb --no-cpu-baseline --no-mi-render > gpurun_out/ncu_c.log 2>&1
  timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_trace_dyn -c 9 --csv --log-file gpurun_out/r02_trace_traffic.csv python bench.py --steps 1 --warmup 0 --no-prb --no-cpu-baseline --no-mi-render > gpurun_out/ncu_d.log 2>&1
fi
