#!/bin/bash
# One GPU call that validates and measures the current tree: the -m gpu suite, the ncu captures of the default bench command
# (tools/ncu_capture.sh), the PRB breakdown and the default bench line.   TAG=r02k bash tools/gpu_check.sh
set -u
TAG=${TAG:-r02k}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout 200 --tb=short ${PYTEST_ARGS:-} 2>&1 | tail -40 > gpurun_out/gpu_tests_${TAG}.log
tail -4 gpurun_out/gpu_tests_${TAG}.log
TAG=$TAG bash tools/ncu_capture.sh > /dev/null 2>&1
timeout 200 python tools/prb_breakdown.py 2>&1 | tail -4 | tee gpurun_out/${TAG}_prb_breakdown.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -2 gpurun_out/${TAG}_bench_n1.err; cut -c1-300 gpurun_out/${TAG}_bench_n1.json
if [ -n "${REF:-}" ]; then
  timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; tail -2 gpurun_out/${TAG}_bench_ref.err; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
fi
