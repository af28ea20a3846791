#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/exp3.jsonl
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -5
run() { env "$@" timeout 300 python tools/pass_times.py 512 >> gpurun_out/exp3.jsonl 2>> gpurun_out/exp3.err; }
run MRL_X=padded
run MRL_NOPAD=1
run MRL_DEBUG_NOFFT=1
run PT_DIMS=256,256,256
run PT_DIMS=128,128,128
run PT_DIMS=1024,256,512
cut -c1-300 gpurun_out/exp3.jsonl; tail -3 gpurun_out/exp3.err
timeout 300 python tools/pass_times.py 512 f32 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err
cut -c1-1600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
