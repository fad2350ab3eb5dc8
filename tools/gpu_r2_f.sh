#!/bin/bash
# round 2, call f (1 GPU): TMA x passes: parity + bench with / without
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_passes.py -m gpu -q -x -k "tma or x_real" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -5 gpurun_out/r2f_pytest.log
FB_X_TMA=1 FB_COLS_TMA=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-one-gpu > gpurun_out/r2f_bench_xtma.json 2> gpurun_out/r2f_bench_xtma.err; echo "rc=$?"
FB_X_TMA=0 FB_COLS_TMA=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-one-gpu > gpurun_out/r2f_bench_x0.json 2> gpurun_out/r2f_bench_x0.err; echo "rc=$?"
FB_X_TMA=1 FB_COLS_TMA=1 timeout 600 python tools/bench_all.py 1024 > gpurun_out/r2f_all_1024.json 2> gpurun_out/r2f_all_1024.err; echo "rc=$?"
