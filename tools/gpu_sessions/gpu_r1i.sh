#!/bin/bash
# round-1 session i: host driver on the GPU (tests/test_gpu_host.py), full GPU suite, smoke, bench, mechanics
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_host.py -m gpu -q --timeout 600 2>&1 | tail -60 > gpurun_out/pytest_host.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_gpu_host.py 2>&1 | tail -30 > gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python tools/mech_bench.py 256 > gpurun_out/mech256.json 2> gpurun_out/mech256.err
tail -60 gpurun_out/pytest_host.log; tail -8 gpurun_out/pytest.log; tail -2 gpurun_out/smoke.log; cut -c1-3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/mech256.json; tail -3 gpurun_out/mech256.err
