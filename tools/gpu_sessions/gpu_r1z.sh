#!/bin/bash
# round-1 session z: re-check the kernel variant switches after hoisting the debug-flag load
mkdir -p gpurun_out
: > gpurun_out/variants.txt
for v in "" "MRL_FUSED_V=1" "MRL_ZFWD_V=1" "MRL_ZFWD_V=2" "MRL_ZFWD_V=3" "MRL_STRIDED_V=1" "MRL_ZINV_V=1" "MRL_ZINV_V=3" "MRL_L2PROMO=256" "MRL_L2PROMO=64"; do
env $v timeout 300 python tools/pass_times.py >> gpurun_out/variants.txt 2>&1
done
cut -c1-260 gpurun_out/variants.txt
