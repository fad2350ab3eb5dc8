#!/bin/bash
# round 2 (1 GPU): Poisson certain-zero shortcut: parity tests of everything that samples counts, stage table, config 4
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_shim.py tests/test_gpu_catalogue.py tests/test_gpu_cube.py -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -3 gpurun_out/r2h_pytest.log
timeout 400 python tools/bench_all.py 1024 > gpurun_out/r2h_all_1024.txt 2> gpurun_out/r2h_all_1024.err; echo "all rc=$?"
grep -E "halo|PCA" gpurun_out/r2h_all_1024.txt
timeout 400 python bench.py --config halos_cross_1024 --steps 5 --warmup 3 > gpurun_out/r2h_cfg_halos_cross_1024.json 2> gpurun_out/r2h_cfg.err; echo "cfg rc=$?"
