#!/bin/bash
# round-1 session ae: PFHub BM1a through the host driver vs the oracle; time per substep of the 2-D benchmarks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host.py -m gpu -q --timeout 600 -k "bm1 or bm2" 2>&1 | tail -30 > gpurun_out/pytest_ae.log
tail -30 gpurun_out/pytest_ae.log | cut -c1-300
for f in bm1_spinodal bm2_ostwald; do
( time marlin_b200/marlin_b200-opt -i tests/inputs/$f.i Executioner/num_steps=10 --output-dir /tmp ) 2>&1 | grep "real\|launches"
( time marlin_b200/marlin_b200-opt -i tests/inputs/$f.i Executioner/num_steps=10 TensorSolver/fuse=false --output-dir /tmp ) 2>&1 | grep "real\|launches"
done
