#!/bin/bash
# round-1 session s: explicit CH (+ de-aliasing), 3-D map_to_aux gold, postprocessor golds through the host driver
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_host.py -m gpu -q --timeout 600 -k "explicit or map_to_aux or postprocessors" 2>&1 | tail -60 > gpurun_out/pytest_s.log
tail -60 gpurun_out/pytest_s.log
