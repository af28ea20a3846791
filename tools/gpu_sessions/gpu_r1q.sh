#!/bin/bash
# round-1 session q: Swift-Hohenberg secant chain through the host driver
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host.py -m gpu -q --timeout 600 -k "swift or mech2d" 2>&1 | tail -40 > gpurun_out/pytest_q.log
tail -40 gpurun_out/pytest_q.log
marlin_b200/marlin_b200-opt -i tests/inputs/swift_hohenberg_secant.i Executioner/num_steps=3 2>&1 | tail -8
