#!/bin/bash
# round-1 session an: float32 mechanics on the fused Green-projection pass
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mech.py -m gpu -q --timeout 300 -k "float32" 2>&1 | tail -25 > gpurun_out/pytest_an.log
tail -25 gpurun_out/pytest_an.log | cut -c1-250
