#!/bin/bash
# round-1 session am: TMA-size coverage - mixed 3-D sizes, buffer-based M/L, AB3, float32
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -k "tma_sizes" 2>&1 | tail -25 > gpurun_out/pytest_am.log
tail -25 gpurun_out/pytest_am.log | cut -c1-250
