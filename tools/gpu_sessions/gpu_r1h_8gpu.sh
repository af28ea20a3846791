#!/bin/bash
# 8-GPU pass: slab bench at N=4 and N=8 (peer mode), N=8 nccl mode
mkdir -p gpurun_out
for cfg in "4 peer" "8 peer" "8 nccl"; do
set -- $cfg
MRL_SLAB_MODE=$2 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus $1 --steps 20 --warmup 5 > gpurun_out/bench_$1gpu_$2.json 2> gpurun_out/bench_$1gpu_$2.err
echo "bench $1 $2 rc=$?"; cut -c1-400 gpurun_out/bench_$1gpu_$2.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_$1gpu_$2.err | tail -3
done
