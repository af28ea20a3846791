#!/bin/bash
# round-1 session l: coupled solver (mrl_coupled_solve + host class), vectorised tangent kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host.py tests/test_gpu_mech.py -m gpu -q --timeout 600 2>&1 | tail -40 > gpurun_out/pytest_l.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 600 -k "coupled" 2>&1 | tail -30 >> gpurun_out/pytest_l.log
timeout 600 python tools/mech_bench.py 256 > gpurun_out/mech256.json 2> gpurun_out/mech256.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_mech.csv python tools/mech_bench.py 256 > /dev/null 2>&1
tail -70 gpurun_out/pytest_l.log; cat gpurun_out/mech256.json; tail -3 gpurun_out/mech256.err
