cd /root/repo
O=gpurun_out
timeout 900 python -m pytest tests/test_slab_gpu.py -q -x --timeout 600 2>&1 | tail -15 > $O/r2b_pytest_slab.log; tail -8 $O/r2b_pytest_slab.log
for cfg in "barrier 0 4" "flags 0 4" "flags 48 4" "flags 48 1"; do
  set -- $cfg
  MRL_SLAB_SYNC=$1 MRL_SLAB_INV_CTAS=$2 MRL_SLAB_CHUNKS=$3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2b_bench_2gpu_$1_$2_c$3.json 2> $O/r2b_bench_2gpu_$1_$2_c$3.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2b_bench_2gpu_$1_$2_c$3.json").read().strip().splitlines()[-1])
    print("$cfg", round(d["value"],1), d["ms_per_step"], d["phases_ms"], d["parity"]["status"], d["parity"].get("rel_l2_c_vs_single_gpu_plan"), d["e2e"]["value"])
except Exception as e:
    print("$cfg FAILED", e); print(open("$O/r2b_bench_2gpu_$1_$2_c$3.err").read()[-1500:])
PY
done
