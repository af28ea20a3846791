#!/bin/bash
# round-1 session ai: generic (non power-of-two) kernels with fewer pencils per CTA on small grids; BM1a object
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_expr.py -m gpu -q --timeout 600 2>&1 | tail -8 > gpurun_out/pytest_ai.log
tail -8 gpurun_out/pytest_ai.log | cut -c1-300
timeout 300 python -c "
import bench, json
print(json.dumps(bench.bm1_bench()))
" 2>&1 | tail -3
