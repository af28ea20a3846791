#!/bin/bash
# round-1 session ac: PFHub BM2a (five coupled fields) through the host driver vs the oracle
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host.py -m gpu -q --timeout 600 -k "bm2" 2>&1 | tail -30 > gpurun_out/pytest_ac.log
tail -30 gpurun_out/pytest_ac.log
(time marlin_b200/marlin_b200-opt -i tests/inputs/bm2_ostwald.i TensorSolver/substeps=2000 Executioner/num_steps=2 Problem/print_debug_output=true --output-dir /tmp) 2>&1 | tail -12
