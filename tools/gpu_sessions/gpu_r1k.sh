#!/bin/bash
# round-1 session k: vectorised CG vector kernels + direction update fused into the tangent kernel; staged transfers (pipelined e2e)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mech.py tests/test_gpu_parity.py -m gpu -q --timeout 600 2>&1 | tail -30 > gpurun_out/pytest_k.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_mech.csv python tools/mech_bench.py 256 > /dev/null 2>&1
tail -12 gpurun_out/pytest_k.log; cut -c1-6000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
