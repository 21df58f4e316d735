set -u
mkdir -p gpurun_out
TAG=r02k bash tools/ncu_capture.sh > /dev/null 2>&1
ls -la gpurun_out | grep r02k | awk '{print $5, $9}'
timeout 300 python -m pytest tests/test_gpu_prb.py tests/test_prb_reference.py -m gpu -q --timeout 200 --tb=short 2>&1 | tail -3
timeout 200 python tools/prb_breakdown.py 2>&1 | tail -4 | tee gpurun_out/r02k_prb_breakdown.log
NS="1" WL="matpreview heightfield_full" bash tools/gpu_scale.sh
