#!/bin/bash
# round-1 session aq: TensorHistogram ([VectorPostprocessors]) through the host driver
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_host.py -m gpu -q --timeout 150 -k "histogram or postprocessors" 2>&1 | tail -25 > gpurun_out/pytest_aq.log
tail -25 gpurun_out/pytest_aq.log | cut -c1-300
