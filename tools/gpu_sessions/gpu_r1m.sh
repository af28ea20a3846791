#!/bin/bash
# round-1 session m: full GPU suite after the mechanics / coupled-solver / staged-transfer work; mech fused-pass variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -15 > gpurun_out/pytest_m.log
for v in 0 1 2; do MRL_MECH_V=$v timeout 300 python tools/mech_bench.py 256 > gpurun_out/mech256_v$v.json 2>> gpurun_out/mech_v.err; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
tail -8 gpurun_out/pytest_m.log; cat gpurun_out/mech256_v*.json | cut -c1-330; tail -2 gpurun_out/smoke.log
