#!/bin/bash
# round 2, call p (8 GPUs): copy kernel alone (CTA sweep), chunk counts, CTAs per peer, NCCL path for comparison
set -x
mkdir -p gpurun_out
run8() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-one-gpu > gpurun_out/r2p_b8_$name.json 2> gpurun_out/r2p_b8_$name.err; echo "b8 $name rc=$?"
}
run8 x2_c8_sweep FB_PUSH_SWEEP=1
run8 x2_c16 FB_CHUNKS=16
run8 x2_c16_p4 FB_CHUNKS=16 FB_DIST_PUSH_CTAS=4
run8 x2_c8_p3 FB_DIST_PUSH_CTAS=3
run8 nccl_c8 FB_DIST_MODE=nccl FB_CHUNKS=8
tail -c 600 gpurun_out/r2p_b8_x2_c8_sweep.err
