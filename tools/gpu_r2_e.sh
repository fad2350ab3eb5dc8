#!/bin/bash
# round 2, call e (1 GPU): packed-f32 FFT arithmetic: parity + A/B/C bench (packed / scalar / packed + Stockham-order prologue)
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_passes.py tests/test_gpu_pipeline.py tests/test_gpu_stats.py tests/test_gpu_bigsize.py -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
tail -5 gpurun_out/r2e_pytest.log
for v in "" _q _s; do
FB_LIB=$PWD/fastbox_b200/libfastbox_b200$v.so FB_COLS_TMA=1 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-one-gpu > gpurun_out/r2e_bench$v.json 2> gpurun_out/r2e_bench$v.err; echo "rc=$?"
done
for v in "" _q; do
FB_LIB=$PWD/fastbox_b200/libfastbox_b200$v.so timeout 600 python bench.py --config filter_beam_poles_1024 --steps 5 --warmup 2 --no-cpu > gpurun_out/r2e_cfg3$v.json 2> gpurun_out/r2e_cfg3$v.err; echo "rc=$?"
done
