#!/bin/bash
# round-1 session o: per-pass memory-side ceilings of the CH substep (butterflies skipped) next to the real passes after hoisting the debug-flag load
mkdir -p gpurun_out
timeout 600 python tools/pass_times.py > gpurun_out/pass_times.txt 2>&1
MRL_DEBUG_NOFFT=1 timeout 600 python tools/pass_times.py > gpurun_out/pass_times_nofft.txt 2>&1
timeout 300 python tools/mech_bench.py 256 > gpurun_out/mech256.json 2> gpurun_out/mech256.err
cat gpurun_out/pass_times.txt gpurun_out/pass_times_nofft.txt gpurun_out/mech256.json
