#!/bin/bash
# round 2, call j (2 GPUs): exchange by the copy kernel (xmode 2): emulated-rank parity, real 2-process check, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -5 gpurun_out/r2j_pytest.log
run2() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-one-gpu > gpurun_out/r2j_b2_$name.json 2> gpurun_out/r2j_b2_$name.err; echo "b2 $name rc=$?"
}
run2 x2_c4 FB_DIST_XMODE=2 FB_CHUNKS=4
run2 x2_c8 FB_DIST_XMODE=2 FB_CHUNKS=8
run2 x2_c8_p24 FB_DIST_XMODE=2 FB_CHUNKS=8 FB_DIST_PUSH_CTAS=24
run2 x0_c4 FB_DIST_XMODE=0 FB_CHUNKS=4
tail -c 400 gpurun_out/r2j_b2_x2_c4.err
