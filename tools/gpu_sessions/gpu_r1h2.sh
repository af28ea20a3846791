#!/bin/bash
# per-file GPU test runs with full logs (crash isolation)
mkdir -p gpurun_out
for f in test_gpu_expr test_gpu_mech test_gpu_parity test_slab_gpu; do
  timeout 900 python -X faulthandler -m pytest tests/$f.py -m gpu -q --timeout 600 -x > gpurun_out/pt_$f.log 2>&1
  echo "$f rc=$?"; grep -n "Fatal\|passed\|failed\|error" gpurun_out/pt_$f.log | head -5
done
timeout 900 python -X faulthandler -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pt_all.log 2>&1
echo "all rc=$?"; grep -n "Fatal\|passed\|failed" gpurun_out/pt_all.log | head; head -c 3000 gpurun_out/pt_all.log
