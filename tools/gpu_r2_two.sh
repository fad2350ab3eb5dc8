#!/bin/bash
# round 2, 2 GPUs: real two-process parity (CUDA IPC peer stores and NCCL), the new RSD 'nearest' test, bench --gpus 2
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_pipeline.py -m gpu -q -rs -k "multiprocess or nearest or emulated" > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest.log
tail -5 gpurun_out/r2t_pytest.log
for mode in p2p nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/dist_check.py --mode $mode > gpurun_out/r2t_dist_check_$mode.log 2>&1; echo "dist_check $mode rc=$?"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2t_bench_2gpu.json 2> gpurun_out/r2t_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 600 gpurun_out/r2t_bench_2gpu.json
