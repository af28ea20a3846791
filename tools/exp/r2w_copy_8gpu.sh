# copy-engine exchange at 8 GPUs: forward y-chunks x y-pass x-chunks
for cfg in "4 2" "2 4" "1 1"; do
set -- $cfg
MRL_SLAB_EXCHANGE=copy MRL_SLAB_CHUNKS=$1 MRL_SLAB_YCHUNKS=$2 MRL_BENCH_HOST_DRIVER=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$1 bench.py --gpus 8 --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('chunks=$1 ychunks=$2', round(d['ms_per_step'],3), list(d['phases_ms'].values()), d['parity']['status'])"
done
