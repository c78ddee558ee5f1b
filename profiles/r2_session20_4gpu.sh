#!/bin/bash
# Round 2, GPU session 20 (4 GPUs): the default command at N = 4 on the final build (the N = 1 / 2 / 8 points are sessions
# 18 / 17 / 15).
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
( time timeout 540 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 > $O/s20_bench_4gpu.json 2> $O/s20_bench_4gpu.err ) 2> $O/s20_bench_4gpu.time
tail -3 $O/s20_bench_4gpu.err
