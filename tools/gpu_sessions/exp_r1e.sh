#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/exp2.jsonl
run() { env "$@" timeout 300 python tools/pass_times.py 512 >> gpurun_out/exp2.jsonl 2>> gpurun_out/exp2.err; }
# H1: aligned rows (nzc = 256) vs misaligned (257), copy mode and real
run MRL_DEBUG_NOFFT=1 PT_DIMS=512,512,510
run PT_DIMS=512,512,510
# H4: 128-byte vs 256-byte row segments on a 256-point strided axis of the same array size
run MRL_DEBUG_NOFFT=1 PT_DIMS=2048,256,512
run MRL_DEBUG_NOFFT=1 PT_DIMS=2048,256,512 MRL_STRIDED_V=1
run PT_DIMS=2048,256,512
run PT_DIMS=2048,256,512 MRL_STRIDED_V=1
# H2: LDG kernels in the same run
run MRL_TMA=0
cut -c1-300 gpurun_out/exp2.jsonl; tail -3 gpurun_out/exp2.err
