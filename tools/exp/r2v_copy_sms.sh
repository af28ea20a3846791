for sms in 0 16 32; do
MRL_SLAB_EXCHANGE=copy MRL_SLAB_COPY_SMS=$sms MRL_BENCH_HOST_DRIVER=0 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$((sms/16)) bench.py --gpus 2 --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('sms=$sms', round(d['ms_per_step'],3), list(d['phases_ms'].values()))"
done
