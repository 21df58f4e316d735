#!/bin/bash
# One call on an 8-GPU box (gpurun --gpus 8): strong scaling of the three multi-GPU workloads. Results: gpurun_out/scale_*.json
set -u
mkdir -p gpurun_out
run() {   # name N workload extra-args...
    local name=$1 n=$2 wl=$3; shift 3
    if [ "$n" = 1 ]; then
        timeout 400 python bench.py --gpus 1 --workload "$wl" "$@" > gpurun_out/scale_${name}_n$n.json 2> gpurun_out/scale_${name}_n$n.err
    else
        timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
            bench.py --gpus $n --workload "$wl" "$@" > gpurun_out/scale_${name}_n$n.json 2> gpurun_out/scale_${name}_n$n.err
    fi
    python - "$name" "$n" <<'PY'
import json, sys
try:
    j = json.loads(open(f"gpurun_out/scale_{sys.argv[1]}_n{sys.argv[2]}.json").read().strip().splitlines()[-1]); p = j["per_rank"]
    print(f"{sys.argv[1]:12s} N={sys.argv[2]} {j['value']:9.1f} Msamples/s {j['ms_per_step']:8.2f} ms/step  e2e {j['e2e']['value']:9.1f}  kernels/rank {['%.2f' % x for x in p['render_kernels_ms']]} allreduce {p['film_allreduce_us']} us"
          + (f"  prb {j['prb']['ms_per_grad_step']:.2f} ms" if 'prb' in j else ""))
except Exception as e:
    print(sys.argv[1], sys.argv[2], "failed:", e)
PY
}
NS=${NS:-"8 4 2"}
for n in $NS; do run cbox $n cornell_box_512x512_256spp_8bounce --steps 10 --warmup 3 --no-cpu-baseline --no-mi-render; done
for n in $NS; do run matpreview $n matpreview_1024x1024_128spp_8bounce --steps 5 --warmup 3 --no-cpu-baseline --no-mi-render --no-prb; done
for n in ${NS_BIG:-8}; do run heightfield_full $n heightfield205k_1920x1080_512spp_8bounce --steps 3 --warmup 3 --no-cpu-baseline --no-mi-render --no-prb; done
