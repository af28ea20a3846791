#!/bin/bash
# round-1c GPU pass: parity tests, variant sweep (LDS + register twiddles + shuffle split), ncu
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 > gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
: > gpurun_out/sweep.jsonl
for v in 0 1 2 3; do
  MRL_STRIDED_V=$v MRL_FUSED_V=$v MRL_ZFWD_V=$v MRL_ZINV_V=$v timeout 300 python tools/pass_times.py 512 >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
done
timeout 300 python tools/pass_times.py 256 >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
timeout 300 python tools/pass_times.py 128 >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
timeout 300 python tools/pass_times.py 512 f32 >> gpurun_out/sweep.jsonl 2>> gpurun_out/sweep.err
cat gpurun_out/sweep.jsonl
tail -5 gpurun_out/sweep.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tma -s 25 -c 5 -o gpurun_out/prof_tma2 python tools/pass_times.py 512 > gpurun_out/ncu_tma.log 2>&1
tail -2 gpurun_out/ncu_tma.log
