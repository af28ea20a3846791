#!/bin/bash
# round-1 session y: ComputeVonMisesStress / ComputeDisplacements through the host driver; staged transfers; full suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -25 > gpurun_out/pytest_y.log
tail -25 gpurun_out/pytest_y.log
