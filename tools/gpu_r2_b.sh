#!/bin/bash
# round 2, second GPU call (2 GPUs): full GPU suite, TMA y pass vs per-thread y pass, exchange variants
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
for v in "0 8" "1 8" "1 4"; do
  set -- $v
  FB_COLS_TMA=$1 FB_CZ_TMA=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-one-gpu > gpurun_out/r2b_bench_tma$1_cz$2.json 2> gpurun_out/r2b_bench_tma$1_cz$2.err; echo "tma $v rc=$?"
done
run2() { # name, env...
  name=$1; shift
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-one-gpu > gpurun_out/r2b_b2_$name.json 2> gpurun_out/r2b_b2_$name.err; echo "b2 $name rc=$?"
}
run2 x0_cz8 FB_DIST_XMODE=0 FB_DIST_CZ=8
run2 x0_cz4 FB_DIST_XMODE=0 FB_DIST_CZ=4
run2 x1_c4 FB_DIST_XMODE=1 FB_CHUNKS=4
run2 x1_c8 FB_DIST_XMODE=1 FB_CHUNKS=8
tail -c 400 gpurun_out/r2b_b2_x1_c4.err
