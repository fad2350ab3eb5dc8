#!/bin/bash
# round 2, call m (8 GPUs): 2048^3 over 8 GPUs, exchange by the high-priority copy kernel (xmode 2)
set -x
mkdir -p gpurun_out
run8() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-one-gpu > gpurun_out/r2m_b8_$name.json 2> gpurun_out/r2m_b8_$name.err; echo "b8 $name rc=$?"
}
run8 x2_c4 FB_DIST_XMODE=2 FB_CHUNKS=4
run8 x2_c8 FB_DIST_XMODE=2 FB_CHUNKS=8
run8 x2_c8_p12 FB_DIST_XMODE=2 FB_CHUNKS=8 FB_DIST_PUSH_CTAS=12
run8 x2_c2 FB_DIST_XMODE=2 FB_CHUNKS=2
tail -c 600 gpurun_out/r2m_b8_x2_c4.err
