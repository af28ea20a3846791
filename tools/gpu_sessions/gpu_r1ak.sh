#!/bin/bash
# round-1 session ak: final validation - full GPU suite, smoke, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12 > gpurun_out/pytest_ak.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -5 gpurun_out/pytest_ak.log; tail -2 gpurun_out/smoke.log; cut -c1-5500 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
