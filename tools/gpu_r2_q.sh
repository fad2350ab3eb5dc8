#!/bin/bash
# round 2, call q (1 GPU): plane-stride probe for the x passes
set -x
mkdir -p gpurun_out
timeout 600 python tools/x_stride_probe.py 1024 > gpurun_out/r2q_xstride.log 2>&1; echo "rc=$?"
grep pad gpurun_out/r2q_xstride.log
