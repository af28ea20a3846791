#!/bin/bash
# 2-GPU pass (round-1 session ag): final check of the multi-GPU bench leg (pipelined e2e) and the slab tests
mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_slab_gpu.py -m gpu -q --timeout 300 > gpurun_out/pt_slab2.log 2>&1
echo "slab tests rc=$?"; tail -3 gpurun_out/pt_slab2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
echo "bench rc=$?"; cut -c1-2500 gpurun_out/bench_2gpu.json; grep -v "OMP_NUM\|\*\*\*\*\|NCCL version" gpurun_out/bench_2gpu.err | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29913 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference > gpurun_out/bench_2gpu_ref.json 2> gpurun_out/bench_2gpu_ref.err
echo "reference arm rc=$?"; cut -c1-400 gpurun_out/bench_2gpu_ref.json
