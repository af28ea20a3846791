#!/bin/bash
# round-1 session r: KKS no-flux model through the host driver
mkdir -p gpurun_out
(time marlin_b200/marlin_b200-opt -i tests/inputs/kks_no_flux.i Executioner/num_steps=1 --output-dir /tmp) 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_host.py -m gpu -q --timeout 600 -k "kks or swift" 2>&1 | tail -40 > gpurun_out/pytest_r.log
tail -40 gpurun_out/pytest_r.log
