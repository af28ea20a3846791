#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/exp.jsonl
run() { env "$@" timeout 300 python tools/pass_times.py 512 >> gpurun_out/exp.jsonl 2>> gpurun_out/exp.err; }
run MRL_X=base
run MRL_L2PROMO=256
run MRL_L2PROMO=0
run MRL_CHUNK_X=8
run MRL_CHUNK_X=16
run MRL_CHUNK_X=32
run MRL_CHUNK_X=16 MRL_L2PROMO=256
cut -c1-330 gpurun_out/exp.jsonl; tail -3 gpurun_out/exp.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:tma -s 60 -c 40 --csv --log-file gpurun_out/chunk_ncu.csv env MRL_CHUNK_X=16 python tools/pass_times.py 512 > /dev/null 2>&1
tail -42 gpurun_out/chunk_ncu.csv | cut -d, -f5,12- | cut -c1-200 | head -60
