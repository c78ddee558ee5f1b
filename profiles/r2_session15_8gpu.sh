#!/bin/bash
# Round 2, GPU session 15 (8 GPUs): the driver's SCALE command at N = 8, once, as a dry run of what it will launch.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/s15_gpus.txt
( time timeout 540 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 > $O/s15_bench_8gpu.json 2> $O/s15_bench_8gpu.err ) 2> $O/s15_bench_8gpu.time
tail -5 $O/s15_bench_8gpu.err
