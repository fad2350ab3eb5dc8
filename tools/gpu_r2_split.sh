#!/bin/bash
# round 2 (1 GPU): xmode 4 (z halves) on emulated ranks
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2x_pytest.log
tail -5 gpurun_out/r2x_pytest.log
