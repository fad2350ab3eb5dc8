#!/bin/bash
# round 2, first GPU call: full GPU test-suite (2 GPUs visible -> real multi-process exchange tests run),
# headline bench, config pipelines, 2-GPU bench in both exchange modes.
set -x
mkdir -p gpurun_out
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; nvidia-smi -L >> gpurun_out/host.txt
nvidia-smi topo -m >> gpurun_out/host.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench1.json 2> gpurun_out/r2a_bench1.err; echo "bench rc=$?"
for c in lognormal_rsd_512 filter_beam_poles_1024 halos_cross_1024; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 2 > gpurun_out/r2a_cfg_$c.json 2> gpurun_out/r2a_cfg_$c.err; echo "cfg $c rc=$?"
done
for mode in p2p nccl; do
  FB_DIST_MODE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2a_bench2_$mode.json 2> gpurun_out/r2a_bench2_$mode.err; echo "bench2 $mode rc=$?"
done
tail -c 600 gpurun_out/r2a_bench2_p2p.err
