#!/bin/bash
# round 2 (1 GPU): quick check of a host-side change: smoke, golden-vector tests, a short bench with the host-buffer leg
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_shim.py -m gpu -q -x -k "golden or box" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -2 gpurun_out/r2q_pytest.log
timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu --no-one-gpu > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"
