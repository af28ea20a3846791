#!/bin/bash
# first GPU pass: tests, smoke, bench, ncu launch list + one full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests -m gpu -q -x --timeout 1200 2>&1 | tail -40 > gpurun_out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o gpurun_out/prof_fused python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_fused.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_strided -s 6 -c 1 -o gpurun_out/prof_strided python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_strided.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zfwd -s 3 -c 1 -o gpurun_out/prof_zfwd python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_zfwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_zinv -s 3 -c 1 -o gpurun_out/prof_zinv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_zinv.log 2>&1
tail -5 gpurun_out/pytest.log; cat gpurun_out/smoke.log | tail -2; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
