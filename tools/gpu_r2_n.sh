#!/bin/bash
# round 2, call n (1 GPU): tile prefetch through the bulk-tensor engine in the x passes, contiguous prefetch in the
# beam x pass: parity with the prefetch on, distance sweeps
set -x
mkdir -p gpurun_out
FB_X_PF=32 FB_BEAM_PF=32 timeout 900 python -m pytest tests/test_gpu_passes.py tests/test_gpu_pipeline.py -m gpu -q -x -k "x_real or beam or roundtrip" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -3 gpurun_out/r2n_pytest.log
B="python bench.py --steps 20 --warmup 3 --no-cpu --no-one-gpu --no-e2e"
for pf in 0 16 32 64 128 256; do
FB_X_PF=$pf timeout 300 $B > gpurun_out/r2n_x_pf$pf.json 2> gpurun_out/r2n_x_pf$pf.err
done
for pf in 0 16 32 64 128; do
FB_BEAM_PF=$pf timeout 300 python tools/ncu_beam.py 1024 3 > gpurun_out/r2n_beam_pf$pf.log 2>&1
done
FB_X_PF=32 timeout 600 python tools/bench_all.py 1024 > gpurun_out/r2n_all_1024_xpf32.json 2> gpurun_out/r2n_all_1024.err
