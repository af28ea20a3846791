#!/bin/bash
# 8-GPU pass (round-1 session x): slab bench 512^3 at N=4 and N=8 with / without the chunked forward phase; 1024^3 at N=8
mkdir -p gpurun_out
for cfg in "8 512 1" "8 512 4" "8 512 2" "4 512 1" "4 512 4" "8 1024 4"; do
set -- $cfg
MRL_BENCH_N=$2 MRL_SLAB_CHUNKS=$3 MRL_SLAB_MODE=peer timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus $1 --steps 20 --warmup 5 > gpurun_out/bench_$1gpu_n$2_c$3.json 2> gpurun_out/bench_$1gpu_n$2_c$3.err
echo "bench $1 gpus n=$2 chunks=$3 rc=$?"; python - <<PY
import json
try:
    s=open('gpurun_out/bench_$1gpu_n$2_c$3.json').read()
    s=s[s.index('{'):]
    d=json.loads(s); print(round(d['value'],1), round(d['ms_per_step'],4), list(d['phases_ms'].values()), d['e2e']['value'])
except Exception as ex:
    print('no json', ex)
PY
grep -v "OMP_NUM\|\*\*\*\*\|NCCL version" gpurun_out/bench_$1gpu_n$2_c$3.err | tail -3
done
