#!/bin/bash
# round 2, third GPU call (1 GPU): full GPU suite incl. the two-point statistics, default bench, launch list
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -15 gpurun_out/r2c_pytest.log
timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
FB_COLS_TMA=1 timeout 900 python bench.py --no-cpu --no-one-gpu > gpurun_out/r2c_bench_tma.json 2> gpurun_out/r2c_bench_tma.err; echo "bench tma rc=$?"
