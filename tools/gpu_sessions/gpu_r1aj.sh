#!/bin/bash
# round-1 session aj: host solver batches the steady-state substeps (CUDA graph) on the context's own stream
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_host.py -m gpu -q --timeout 600 2>&1 | tail -30 > gpurun_out/pytest_aj.log
tail -30 gpurun_out/pytest_aj.log | cut -c1-300
for f in bm1_spinodal bm2_ostwald; do
( time marlin_b200/marlin_b200-opt -i tests/inputs/$f.i Executioner/num_steps=10 --output-dir /tmp ) 2>&1 | grep "real\|launches"
( time marlin_b200/marlin_b200-opt -i tests/inputs/$f.i Executioner/num_steps=10 TensorSolver/batch_substeps=false --output-dir /tmp ) 2>&1 | grep "real\|launches"
done
