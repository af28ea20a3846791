#!/bin/bash
# round-1 session p: 2-D mechanics (2x2 tensors) on the CUDA path
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mech.py tests/test_gpu_host.py -m gpu -q --timeout 600 -k "mech" 2>&1 | tail -40 > gpurun_out/pytest_p.log
tail -40 gpurun_out/pytest_p.log
