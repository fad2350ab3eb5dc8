#!/bin/bash
# round 2 (1 GPU): 128 x 128-tile PCA covariance kernel: parity tests, variants at 1024^3 / 512^3, ncu of both kernels
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cube.py -m gpu -q -x > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest.log
tail -3 gpurun_out/r2p_pytest.log
timeout 600 python tools/pca_variants.py 1024 64:0:0 128:8:9 128:16:9 128:16:18 > gpurun_out/r2p_variants_1024.txt 2>&1; echo "variants rc=$?"
cat gpurun_out/r2p_variants_1024.txt | grep "pca cov"
timeout 300 python tools/pca_variants.py 512 64:0:0 128:8:9 128:16:9 > gpurun_out/r2p_variants_512.txt 2>&1
grep "pca cov" gpurun_out/r2p_variants_512.txt
PCA_REPS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_pca_cov -c 3 -o gpurun_out/prof_r2_pca python tools/pca_variants.py 1024 64:0:0 128:16:9 > gpurun_out/r2p_ncu.log 2>&1; echo "ncu rc=$?"
