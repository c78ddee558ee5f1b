#!/bin/bash
# Round 2, GPU session 12 (2 GPUs): warp-parallel sampler in the verifier, e2e leg with a warm-up pass and breakdown,
# block-cyclic sharding of the pair list; 1-GPU default line, then the same command on 2 GPUs.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_verify_gpu.py tests/test_pipeline_gpu.py tests/test_two_view_shim.py tests/test_multigpu.py -m gpu -q 2>&1 | tail -8 ) > $O/s12_pytest.log 2>&1
B2_VERIFY_PROFILE=1 timeout 300 python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 --seq-images 1000 --pairs -1 --ba "" --retrieval-words 0 > $O/s12_verify_1k.json 2> $O/s12_verify_1k.err
( time timeout 1500 python bench.py > $O/s12_bench_default.json 2> $O/s12_bench_default.err ) 2> $O/s12_bench_default.time
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 > $O/s12_bench_2gpu.json 2> $O/s12_bench_2gpu.err ) 2> $O/s12_bench_2gpu.time
ls -la $O | tail -8
