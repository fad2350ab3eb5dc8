#!/bin/bash
# round 2, call s (2 GPUs): bulk-copy-engine pusher (xmode 3): emulated-rank parity, real 2-process check, bench vs xmode 2
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
tail -5 gpurun_out/r2s_pytest.log
run2() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-one-gpu > gpurun_out/r2s_b2_$name.json 2> gpurun_out/r2s_b2_$name.err; echo "b2 $name rc=$?"
}
run2 x3 FB_DIST_XMODE=3 FB_PUSH_SWEEP=1
run2 x2 FB_DIST_XMODE=2
tail -c 400 gpurun_out/r2s_b2_x3.err
