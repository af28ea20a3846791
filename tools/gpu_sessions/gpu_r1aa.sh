#!/bin/bash
# round-1 session aa: BroydenSolver and TensorInterfaceVelocityPostprocessor through the host driver
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host.py -m gpu -q --timeout 600 -k "broyden or interface" 2>&1 | tail -40 > gpurun_out/pytest_aa.log
tail -40 gpurun_out/pytest_aa.log
marlin_b200/marlin_b200-opt -i tests/inputs/broyden_coupled.i TensorSolver/verbose=true 2>&1 | grep -i "converged\|Time Step" | head
