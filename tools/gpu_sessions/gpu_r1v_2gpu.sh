#!/bin/bash
# 2-GPU pass (round-1 session v): forward phase in y-chunks (z pass of chunk i+1 overlaps the x pass + peer stores of chunk i)
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_slab_gpu.py -m gpu -q --timeout 300 > gpurun_out/pt_slab2.log 2>&1
echo "slab tests rc=$?"; tail -5 gpurun_out/pt_slab2.log
for cfg in "1 48" "2 48" "4 48" "4 32" "8 48" "4 64"; do
set -- $cfg
MRL_SLAB_CHUNKS=$1 MRL_SLAB_XCTAS=$2 MRL_SLAB_MODE=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_c$1_x$2.json 2> gpurun_out/bench_2gpu_c$1_x$2.err
echo "bench chunks=$1 xctas=$2 rc=$?"; python - <<PY
import json
s=open('gpurun_out/bench_2gpu_c$1_x$2.json').read()
s=s[s.index('{'):]
d=json.loads(s); print(round(d['value'],1), round(d['ms_per_step'],4), list(d['phases_ms'].values()))
PY
grep -v "OMP_NUM\|\*\*\*\*\|NCCL version" gpurun_out/bench_2gpu_c$1_x$2.err | tail -3
done
