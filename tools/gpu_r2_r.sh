#!/bin/bash
# round 2, call r (1 GPU): one-load twiddles: full GPU suite + headline bench + every stage
set -x
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
tail -4 gpurun_out/r2r_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu --no-one-gpu --no-e2e > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err
timeout 600 python tools/bench_all.py 1024 > gpurun_out/r2r_all_1024.txt 2> gpurun_out/r2r_all_1024.err
