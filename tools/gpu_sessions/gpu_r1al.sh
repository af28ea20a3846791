#!/bin/bash
# round-1 session al: 2-D power-of-two sizes (BASELINE configs[0]: 256^2) on the TMA-pipelined passes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "100_substeps" 2>&1 | tail -12 > gpurun_out/pytest_al.log
tail -12 gpurun_out/pytest_al.log | cut -c1-300
PT_DIMS=256,256 timeout 120 python tools/pass_times.py 2>&1 | tail -2 | cut -c1-300
