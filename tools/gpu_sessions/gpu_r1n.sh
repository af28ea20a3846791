#!/bin/bash
# round-1 session n: where does the fused x pass lose time?  (a) memory-side ceiling with the butterflies skipped, (b) ncu full set
mkdir -p gpurun_out
MRL_DEBUG_NOFFT=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_mech_nofft.csv python tools/mech_bench.py 256 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_mech_fused_tma -s 3 -c 1 -o gpurun_out/mech_fused_full -f python tools/mech_bench.py 256 > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/mech_fused_full.ncu-rep --page raw --csv > gpurun_out/mech_fused_raw.csv 2>/dev/null
ncu -i gpurun_out/mech_fused_full.ncu-rep --page source --csv > gpurun_out/mech_fused_source.csv 2>/dev/null
ls -la gpurun_out | tail -8; tail -3 gpurun_out/ncu_full.log
