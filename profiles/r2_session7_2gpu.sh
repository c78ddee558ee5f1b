#!/bin/bash
# Round 2, GPU session 7 (2 GPUs): library-owned NCCL in the bundle adjuster (exact + iterative), one candidate list
# sharded over two ranks in bench.py.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/s7_gpus.txt
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -x 2>&1 | tail -8 > $O/s7_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --seq-images 1000 --pairs 20000 > $O/s7_bench_2gpu.json 2> $O/s7_bench_2gpu.err
timeout 600 python bench.py --gpus 1 --steps 2 --warmup 1 --seq-images 1000 --pairs 20000 --no-cpu > $O/s7_bench_1gpu.json 2> $O/s7_bench_1gpu.err
B2_BENCH_BA_HOOK=torch timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 1 --seq-images 100 --seq-cand 10 --pairs -1 --no-cpu --no-e2e > $O/s7_ba_2gpu_hook.json 2> $O/s7_ba_2gpu_hook.err
ls -la $O | tail -6
