set -u
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 200 --tb=short 2>&1 | tail -15 > gpurun_out/gpu_tests_r02h.log; grep -n 'AssertionError\|passed\|failed' gpurun_out/gpu_tests_r02h.log | head -5
B="python bench.py --steps 1 --warmup 1 --no-prb --no-cpu-baseline --no-mi-render"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_flat -s 3 -c 1 -f -o gpurun_out/r02h_trace_flat $B > gpurun_out/r02h_ncu_b.log 2>&1
timeout 200 python bench.py --steps 5 --warmup 3 --no-prb --no-cpu-baseline --no-mi-render --workload heightfield205k_1024x1024_64spp_8bounce > gpurun_out/r02h_h.json 2>/dev/null; python -c "
import json; j=json.loads(open('gpurun_out/r02h_h.json').read().strip().splitlines()[-1]); print('heightfield', j['value'], j['ms_per_step'], j['roofline']['avg_launch_ms'])"
