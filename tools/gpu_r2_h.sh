#!/bin/bash
# round 2, call h (8 GPUs): 2048^3 over 8 GPUs, exchange inside the library (peer stores / copy engine)
set -x
mkdir -p gpurun_out
run8() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-one-gpu > gpurun_out/r2h_b8_$name.json 2> gpurun_out/r2h_b8_$name.err; echo "b8 $name rc=$?"
}
run8 x0_c4 FB_DIST_XMODE=0 FB_CHUNKS=4
run8 x1_c4 FB_DIST_XMODE=1 FB_CHUNKS=4
run8 x0_c8 FB_DIST_XMODE=0 FB_CHUNKS=8
tail -c 600 gpurun_out/r2h_b8_x0_c4.err
