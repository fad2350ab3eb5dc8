#!/bin/bash
# round 2, call g (1 GPU): beam convolution with a cached spectrum: parity + config-3 bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_shim.py -m gpu -q -x -k "beam" > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -5 gpurun_out/r2g_pytest.log
timeout 600 python bench.py --config filter_beam_poles_1024 --steps 5 --warmup 2 --no-cpu > gpurun_out/r2g_cfg3.json 2> gpurun_out/r2g_cfg3.err; echo "rc=$?"
timeout 600 python tools/ncu_beam.py > gpurun_out/r2g_beam_plain.log 2>&1; echo "rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2g_beam_launches.csv python tools/ncu_beam.py > gpurun_out/r2g_ncu_beam.log 2>&1; echo "rc=$?"
