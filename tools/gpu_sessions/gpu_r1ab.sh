#!/bin/bash
# round-1 session ab: Broyden kernels vs torch, Broyden through the host driver
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host.py tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "broyden" 2>&1 | tail -30 > gpurun_out/pytest_ab.log
tail -30 gpurun_out/pytest_ab.log
