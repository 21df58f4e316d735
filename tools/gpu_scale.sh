#!/bin/bash
# Strong scaling of the multi-GPU workloads (BASELINE.json configs[1], [3], [4]) on one box:
#   gpurun --gpus 8 -- 'NS="8" WL="matpreview heightfield_full" bash tools/gpu_scale.sh'      results: gpurun_out/scale_*.json
# NS: rank counts to run (each <= the GPUs of the box), WL: which of cbox / matpreview / heightfield_full.
set -u
mkdir -p gpurun_out
run() {   # name N workload extra-args...
    local name=$1 n=$2 wl=$3; shift 3
    if [ "$n" = 1 ]; then
        timeout 500 python bench.py --gpus 1 --workload "$wl" "$@" > gpurun_out/scale_${name}_n$n.json 2> gpurun_out/scale_${name}_n$n.err
    else
        timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
            bench.py --gpus $n --workload "$wl" "$@" > gpurun_out/scale_${name}_n$n.json 2> gpurun_out/scale_${name}_n$n.err
    fi
    python - "$name" "$n" <<'PY'
import json, sys
try:
    j = json.loads(open(f"gpurun_out/scale_{sys.argv[1]}_n{sys.argv[2]}.json").read().strip().splitlines()[-1]); p = j["per_rank"]
    print(f"{sys.argv[1]:16s} N={sys.argv[2]} {j['value']:9.1f} Msamples/s {j['ms_per_step']:9.2f} ms/step  e2e {j['e2e']['value']:9.1f}  kernels/rank {['%.2f' % x for x in p['render_kernels_ms']]} allreduce {p['film_allreduce_us']} us"
          + (f"  prb {j['prb']['ms_per_grad_step']:.2f} ms" if 'prb' in j else ""))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed:", e)
PY
}
NS=${NS:-"8 4 2"}
WL=${WL:-"cbox matpreview heightfield_full"}
for w in $WL; do
  for n in $NS; do
    case $w in
      cbox) run cbox $n cornell_box_512x512_256spp_8bounce --steps 10 --warmup 3 --no-cpu-baseline --no-mi-render;;
      matpreview) run matpreview $n matpreview_1024x1024_128spp_8bounce --steps 4 --warmup 3 --no-cpu-baseline --no-mi-render --no-prb;;
      heightfield_full) run heightfield_full $n heightfield205k_1920x1080_512spp_8bounce --steps 2 --warmup 3 --no-cpu-baseline --no-mi-render --no-prb;;
    esac
  done
done
