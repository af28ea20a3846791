#!/bin/bash
# round-1 re-entry GPU pass: tests, smoke, bench, mech bench, ncu launch list + full captures of the TMA passes
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -30 > gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python tools/mech_bench.py 256 > gpurun_out/mech256.json 2> gpurun_out/mech256.err
timeout 300 python tools/mech_bench.py 128 > gpurun_out/mech128.json 2>> gpurun_out/mech256.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_ncu.log 2>&1
for k in k_fused_tma k_strided_tma k_zfwd_tma k_zinv_tma; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 1 -o gpurun_out/prof_$k -f python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_$k.log 2>&1
done
tail -8 gpurun_out/pytest.log; tail -2 gpurun_out/smoke.log; cut -c1-3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/mech256.json gpurun_out/mech128.json; tail -3 gpurun_out/mech256.err
