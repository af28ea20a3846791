#!/bin/bash
# r1ar: SmoothRectangleCompute through the driver vs gold smooth_rectangle.h5
mkdir -p gpurun_out/r1ar
timeout 25 python -m pytest tests/test_gpu_host.py -q -x -k "smooth_rectangle" > gpurun_out/r1ar/pytest.log 2>&1
tail -5 gpurun_out/r1ar/pytest.log
