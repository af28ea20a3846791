#!/bin/bash
# round-1 session ad: the host solver's fusion pattern (parenthesised product) - fused path through the driver
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_host.py -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/pytest_ad.log
tail -60 gpurun_out/pytest_ad.log | cut -c1-300
(time marlin_b200/marlin_b200-opt -i tests/inputs/bm2_ostwald.i TensorSolver/substeps=2000 Executioner/num_steps=2 Problem/print_debug_output=true --output-dir /tmp) 2>&1 | grep -v "^ *->" | tail -8
