cd /root/repo
O=gpurun_out
NG=${NG:-8}
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/a2a_probe.py > $O/r2d_a2a_${NG}gpu.json 2> $O/r2d_a2a.err; cat $O/r2d_a2a_${NG}gpu.json; tail -3 $O/r2d_a2a.err
