cd /root/repo
O=gpurun_out
NG=${NG:-8}
timeout 600 python -m pytest tests/test_slab_gpu.py -q -x --timeout 500 2>&1 | tail -15 > $O/r2c_pytest_slab_${NG}gpu.log; tail -4 $O/r2c_pytest_slab_${NG}gpu.log
run() { # sync inv chunks n
  MRL_BENCH_N=$4 MRL_SLAB_SYNC=$1 MRL_SLAB_INV_CTAS=$2 MRL_SLAB_CHUNKS=$3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 20 --warmup 5 > $O/r2c_bench_${NG}gpu_n$4_$1_$2_c$3.json 2> $O/r2c_bench_${NG}gpu_n$4_$1_$2_c$3.err
  python - <<PY
import json
try:
    d=json.loads(open("$O/r2c_bench_${NG}gpu_n$4_$1_$2_c$3.json").read().strip().splitlines()[-1])
    print("$*", round(d["value"],1), round(d["ms_per_step"],4), list(d["phases_ms"].values()), d["parity"]["status"], d["parity"].get("rel_l2_c_vs_single_gpu_plan"), round(d["e2e"]["value"],1))
except Exception as e:
    print("$* FAILED", e); print(open("$O/r2c_bench_${NG}gpu_n$4_$1_$2_c$3.err").read()[-1500:])
PY
}
run barrier 0 4 512
run barrier 0 1 512
run barrier 0 2 512
run flags 48 4 512
run barrier 0 4 1024
