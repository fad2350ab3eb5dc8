#!/bin/bash
# round 2, call d (1 GPU): parity of the Stockham-order noise prologue + A/B bench against the quad-order build
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_shim.py tests/test_gpu_stats.py tests/test_gpu_bigsize.py tests/test_gpu_scale.py -m gpu -q -x > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -5 gpurun_out/r2d_pytest.log
FB_COLS_TMA=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-one-gpu > gpurun_out/r2d_bench_stockham.json 2> gpurun_out/r2d_bench_stockham.err; echo "rc=$?"
FB_LIB=$PWD/fastbox_b200/libfastbox_b200_q.so FB_COLS_TMA=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-one-gpu > gpurun_out/r2d_bench_quad.json 2> gpurun_out/r2d_bench_quad.err; echo "rc=$?"
