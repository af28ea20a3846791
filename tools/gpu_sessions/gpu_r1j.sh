#!/bin/bash
# round-1 session j: mechanics with the Green projection fused into the x pass
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mech.py tests/test_gpu_host.py -m gpu -q --timeout 600 2>&1 | tail -30 > gpurun_out/pytest_mech.log
timeout 600 python tools/mech_bench.py 256 > gpurun_out/mech256.json 2> gpurun_out/mech256.err
MRL_MECH_FUSED=0 timeout 600 python tools/mech_bench.py 256 > gpurun_out/mech256_unfused.json 2>> gpurun_out/mech256.err
timeout 600 python tools/mech_bench.py 128 > gpurun_out/mech128.json 2>> gpurun_out/mech256.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_mech.csv python tools/mech_bench.py 256 > /dev/null 2>&1
tail -12 gpurun_out/pytest_mech.log; cat gpurun_out/mech256.json gpurun_out/mech256_unfused.json gpurun_out/mech128.json; tail -3 gpurun_out/mech256.err
