#!/bin/bash
# round 2 (1 GPU): every hot-path row of SURVEY section 8 at 1024^3 (tools/bench_all.py)
mkdir -p gpurun_out
timeout 200 python tools/bench_all.py 1024 > gpurun_out/r2s_all_1024.txt 2> gpurun_out/r2s_all_1024.err; echo "all rc=$?"
