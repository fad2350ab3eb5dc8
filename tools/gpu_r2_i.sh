#!/bin/bash
# round 2, call i (1 GPU): RSD remap in cell units (parity + timing), L2 prefetch distance sweep for the first pass,
# ncu of the TMA x pass
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_shim.py tests/test_gpu_bigsize.py -m gpu -q -x -k "rsd or redshift" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -5 gpurun_out/r2i_pytest.log
timeout 600 python bench.py --config lognormal_rsd_512 --steps 5 --warmup 2 --no-cpu > gpurun_out/r2i_cfg1.json 2> gpurun_out/r2i_cfg1.err; echo "rc=$?"
for pf in 0 148 296 444 888; do
FB_ROWS_PF=$pf timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu --no-one-gpu > gpurun_out/r2i_bench_pf$pf.json 2> gpurun_out/r2i_bench_pf$pf.err; echo "rc=$?"
done
timeout 600 python tools/ncu_rsd.py 1024 > gpurun_out/r2i_rsd_1024.log 2>&1
FB_X_TMA=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_x_c2r_tma -c 1 -o gpurun_out/prof_r2_x_tma python bench.py --steps 1 --warmup 1 --no-cpu --no-one-gpu > gpurun_out/r2i_ncu_x.log 2>&1; echo "ncu rc=$?"
