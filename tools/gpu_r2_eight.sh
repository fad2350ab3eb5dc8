#!/bin/bash
# round 2, 8 GPUs: bench.py --gpus 8 (2048^3, Philox, exchange inside the library), exchange mechanisms compared in one run
set -x
mkdir -p gpurun_out
FB_DIST_COMPARE=${FB_DIST_COMPARE:-3,2:8,2} timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2e_bench_8gpu.json 2> gpurun_out/r2e_bench_8gpu.err; echo "bench8 rc=$?"
tail -c 2500 gpurun_out/r2e_bench_8gpu.json
tail -5 gpurun_out/r2e_bench_8gpu.err
