#!/bin/bash
# 8-GPU pass (round-1 session t): slab bench 512^3 at N=4 and N=8 (peer), 1024^3 at N=8 (peer)
mkdir -p gpurun_out
for cfg in "4 512" "8 512" "8 1024"; do
set -- $cfg
MRL_BENCH_N=$2 MRL_SLAB_MODE=peer timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus $1 --steps 20 --warmup 5 > gpurun_out/bench_$1gpu_n$2.json 2> gpurun_out/bench_$1gpu_n$2.err
echo "bench $1 gpus n=$2 rc=$?"; cut -c1-330 gpurun_out/bench_$1gpu_n$2.json; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_$1gpu_n$2.err | tail -3
done
