#!/bin/bash
# One parameterised GPU session (replaces the per-session scripts of round 1).  Run under gpurun:
#   gpurun --timeout 1500 -- 'bash tools/gpu_session.sh TAG step [step ...]'
# steps: tests[:pytest-args]  testsel:"paths"  smoke  bench  bench:N(torchrun N ranks)  launches  ncu:<kernel-regex>  traffic  mech  passes
# Everything lands in gpurun_out/<TAG>_*.
TAG=$1; shift
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > $O/${TAG}_gpu.txt 2>&1
for step in "$@"; do
  kind=${step%%:*}; arg=""; [[ "$step" == *:* ]] && arg=${step#*:}
  case $kind in
    tests)    timeout 1700 python -m pytest tests -m gpu -q --timeout 900 -x $arg 2>&1 | tail -40 > $O/${TAG}_pytest.log; tail -15 $O/${TAG}_pytest.log ;;
    testsel)  timeout 1700 python -m pytest $arg -m gpu -q --timeout 900 --durations=15 2>&1 | tail -60 > $O/${TAG}_pytest.log; tail -25 $O/${TAG}_pytest.log ;;
    testsall) timeout 1700 python -m pytest tests -m gpu -q --timeout 900 --durations=25 $arg 2>&1 | tail -80 > $O/${TAG}_pytest.log; tail -25 $O/${TAG}_pytest.log ;;
    smoke)    timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; tail -3 $O/${TAG}_smoke.log ;;
    bench)    if [ -z "$arg" ] || [ "$arg" == 1 ]; then
                timeout 900 python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err
              else
                timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $arg --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $arg --steps 20 --warmup 5 > $O/${TAG}_bench_${arg}gpu.json 2> $O/${TAG}_bench_${arg}gpu.err
              fi
              cut -c1-6000 $O/${TAG}_bench*.json | tail -2; tail -3 $O/${TAG}_bench*.err ;;
    benchfast) timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --fast > $O/${TAG}_benchfast.json 2> $O/${TAG}_benchfast.err; cut -c1-3000 $O/${TAG}_benchfast.json; tail -3 $O/${TAG}_benchfast.err ;;
    launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 80 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --fast > $O/${TAG}_bench_ncu.log 2>&1 ;;
    ncu)      timeout 600 ncu --set full --clock-control none --import-source on -k regex:$arg -s 4 -c 1 -o $O/${TAG}_prof_$arg -f python bench.py --steps 2 --warmup 3 --no-cpu --fast > $O/${TAG}_ncu_$arg.log 2>&1 ;;
    traffic)  timeout 600 python tools/measure_traffic.py > $O/${TAG}_traffic.log 2>&1; tail -3 $O/${TAG}_traffic.log; cp profiles/traffic.json $O/${TAG}_traffic.json ;;
    mech)     timeout 600 python tools/mech_bench.py ${arg:-256} > $O/${TAG}_mech.json 2> $O/${TAG}_mech.err; cat $O/${TAG}_mech.json; tail -2 $O/${TAG}_mech.err ;;
    passes)   timeout 300 env $arg python tools/pass_times.py 512 >> $O/${TAG}_passes.jsonl 2>> $O/${TAG}_passes.err; tail -1 $O/${TAG}_passes.jsonl ;;
    sh)       timeout 900 bash -c "$arg" > $O/${TAG}_sh.log 2>&1; tail -20 $O/${TAG}_sh.log ;;
  esac
done
