#!/bin/bash
# 2-GPU pass: slab parity tests, slab bench in peer and nccl modes
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus2.txt 2>&1
nvidia-smi topo -m >> gpurun_out/gpus2.txt 2>&1
timeout 600 python -X faulthandler -m pytest tests/test_slab_gpu.py -m gpu -q --timeout 300 > gpurun_out/pt_slab2.log 2>&1
echo "slab tests rc=$?"; tail -5 gpurun_out/pt_slab2.log
for mode in peer nccl; do
MRL_SLAB_MODE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_$mode.json 2> gpurun_out/bench_2gpu_$mode.err
echo "bench $mode rc=$?"; cut -c1-1500 gpurun_out/bench_2gpu_$mode.json; tail -3 gpurun_out/bench_2gpu_$mode.err
done
