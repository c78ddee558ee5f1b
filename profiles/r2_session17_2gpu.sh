#!/bin/bash
# Round 2, GPU session 17 (2 GPUs): sharded retrieval leg (all-gather of the word ids over NCCL), the new C3-chain and
# concurrency tests, the default command under torchrun.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_pipeline_gpu.py tests/test_concurrency_gpu.py tests/test_retrieval_gpu.py -m gpu -q 2>&1 | tail -12 ) > $O/s17_pytest.log 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 > $O/s17_bench_2gpu.json 2> $O/s17_bench_2gpu.err ) 2> $O/s17_bench_2gpu.time
tail -5 $O/s17_bench_2gpu.err
