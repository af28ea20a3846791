#!/bin/bash
# round-1 session ao: XDMFTensorOutput through the host driver (gold cahnhilliard.xmf), mechanics inputs with outputs active
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_host.py -m gpu -q --timeout 300 -k "xdmf or mech or ch2d_input" 2>&1 | tail -25 > gpurun_out/pytest_ao.log
tail -25 gpurun_out/pytest_ao.log | cut -c1-300
