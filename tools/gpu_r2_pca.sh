#!/bin/bash
# round 2 (1 GPU): PCA covariance on the FP64 tensor path: parity tests, stage table
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cube.py -m gpu -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -3 gpurun_out/r2p_pytest.log
timeout 300 python tools/pca_variants.py 1024 64:0:0 128:16:9 128:8:9 > gpurun_out/r2p_variants_1024.txt 2>&1; echo "variants rc=$?"
grep "pca cov" gpurun_out/r2p_variants_1024.txt
PCA_REPS=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_pca_cov_mma -c 1 -o gpurun_out/prof_r2_pca_mma python tools/pca_variants.py 1024 128:16:9 > gpurun_out/r2p_ncu.log 2>&1; echo "ncu rc=$?"
