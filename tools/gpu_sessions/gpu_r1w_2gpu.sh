#!/bin/bash
# 2-GPU pass (round-1 session w): sweep of the chunked forward phase (chunks x CTAs given to the x pass)
mkdir -p gpurun_out
for cfg in "2 64" "2 80" "2 96" "4 80" "4 96" "4 112" "8 96"; do
set -- $cfg
MRL_SLAB_CHUNKS=$1 MRL_SLAB_XCTAS=$2 MRL_SLAB_MODE=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_c$1_x$2.json 2> gpurun_out/bench_2gpu_c$1_x$2.err
echo "bench chunks=$1 xctas=$2 rc=$?"; python - <<PY
import json
s=open('gpurun_out/bench_2gpu_c$1_x$2.json').read()
s=s[s.index('{'):]
d=json.loads(s); print(round(d['value'],1), round(d['ms_per_step'],4), list(d['phases_ms'].values()))
PY
done
