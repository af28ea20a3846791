#!/bin/bash
# round-1 session ap: FFTQuasistaticElasticity / FFTElasticChemicalPotential through the host driver vs the oracle
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_host.py -m gpu -q --timeout 200 -k "quasistatic" 2>&1 | tail -25 > gpurun_out/pytest_ap.log
tail -25 gpurun_out/pytest_ap.log | cut -c1-300
