#!/bin/bash
# round 2, call k (1 GPU): L2 prefetch distance sweeps (first pass, x pass, per-thread y pass)
set -x
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-one-gpu --no-e2e"
for pf in 37 74 111 148 185 222; do
FB_ROWS_PF=$pf timeout 300 $B > gpurun_out/r2k_rows_pf$pf.json 2> gpurun_out/r2k_rows_pf$pf.err
done
for pf in 74 148 296 444; do
FB_ROWS_PF=148 FB_X_PF=$pf timeout 300 $B > gpurun_out/r2k_x_pf$pf.json 2> gpurun_out/r2k_x_pf$pf.err
done
for pf in 0 74 148 296; do
FB_ROWS_PF=148 FB_COLS_TMA=0 FB_COLS_PF=$pf timeout 300 $B > gpurun_out/r2k_cols_pf$pf.json 2> gpurun_out/r2k_cols_pf$pf.err
done
FB_ROWS_PF=148 FB_X_PF=148 timeout 600 python tools/bench_all.py 1024 > gpurun_out/r2k_all_1024_pf.json 2> gpurun_out/r2k_all_1024_pf.err
timeout 600 python tools/bench_all.py 1024 > gpurun_out/r2k_all_1024_nopf.json 2> gpurun_out/r2k_all_1024_nopf.err
# TMA x passes, single buffer, two CTAs per SM
timeout 600 python -m pytest tests/test_gpu_passes.py -m gpu -q -x -k "tma" > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2k_pytest.log
for c in 1 2 3; do
FB_ROWS_PF=148 FB_X_TMA=1 FB_X_TMA_CTAS=$c timeout 300 $B > gpurun_out/r2k_xtma_c$c.json 2> gpurun_out/r2k_xtma_c$c.err
done
