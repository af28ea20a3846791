#!/bin/bash
# 2-GPU pass (round-1 session u): device-side slab barrier vs NCCL all-reduce barrier; per-phase times
mkdir -p gpurun_out
timeout 600 python -X faulthandler -m pytest tests/test_slab_gpu.py -m gpu -q --timeout 300 > gpurun_out/pt_slab2.log 2>&1
echo "slab tests rc=$?"; tail -5 gpurun_out/pt_slab2.log
for bar in dev nccl; do
MRL_SLAB_BARRIER=$bar MRL_SLAB_MODE=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu_$bar.json 2> gpurun_out/bench_2gpu_$bar.err
echo "bench $bar rc=$?"; python -c "
import json;d=json.load(open('gpurun_out/bench_2gpu_$bar.json'));print(d['value'],d['ms_per_step'],d['phases_ms'])"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_2gpu_$bar.err | tail -3
done
