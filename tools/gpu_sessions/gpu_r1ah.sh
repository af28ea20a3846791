#!/bin/bash
# round-1 session ah: CUDA-graph replay of the steady-state substep sequence (mrl_split_substeps); BM1a object of the bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "graph_replay" 2>&1 | tail -30 > gpurun_out/pytest_ah.log
tail -30 gpurun_out/pytest_ah.log | cut -c1-300
timeout 300 python -c "
import bench, json
print(json.dumps(bench.bm1_bench()))
" 2>&1 | tail -3
