#!/bin/bash
# ncu evidence for the default bench line (one GPU): launch list, full captures of the traversal and the diffuse shading kernel, DRAM traffic
# of the nine traversal launches of a frame. Usage: TAG=r02f bash tools/ncu_capture.sh   -> gpurun_out/${TAG}_*
set -u
TAG=${TAG:-r02f}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-prb --no-cpu-baseline --no-mi-render ${BENCH_ARGS:-}"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 3 -c 1 -f -o gpurun_out/${TAG}_trace $B > gpurun_out/${TAG}_ncu_b.log 2>&1
# (one full report is ~33 MB and gpurun brings back 64 MB: the shading kernel's key metrics go to a CSV instead of a second report)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread --clock-control none -k regex:k_shade -s 2 -c 1 --csv --log-file gpurun_out/${TAG}_shade.csv $B > gpurun_out/${TAG}_ncu_c.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_trace -c 9 --csv --log-file gpurun_out/${TAG}_trace_traffic.csv python bench.py --steps 1 --warmup 0 --no-prb --no-cpu-baseline --no-mi-render ${BENCH_ARGS:-} > gpurun_out/${TAG}_ncu_d.log 2>&1
ls -la gpurun_out | grep ${TAG}
