#!/bin/bash
# round 2, 4 GPUs: bench.py --gpus 4, copy-kernel CTA counts and the bulk-copy pusher compared under load in one run
set -x
mkdir -p gpurun_out
FB_DIST_COMPARE=${FB_DIST_COMPARE:-2:5,2:20,3,3:20,2} timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2f_bench_4gpu.json 2> gpurun_out/r2f_bench_4gpu.err; echo "bench4 rc=$?"
tail -c 600 gpurun_out/r2f_bench_4gpu.err
