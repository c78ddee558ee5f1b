#!/bin/bash
# Round 2, GPU session 14: tcgen05 word search with 16 epilogue warps (two column halves per TMEM quadrant, merge per item);
# two-thread / two-handle concurrency test.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_retrieval_gpu.py tests/test_concurrency_gpu.py -m gpu -q 2>&1 | tail -12 ) > $O/s14_pytest.log 2>&1
timeout 600 python bench.py --steps 1 --warmup 1 --pairs -1 --ba '' --ba-c5 '' --no-e2e > $O/s14_bench_retrieval.json 2> $O/s14_bench_retrieval.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'word_knn_tc' -c 1 -o $O/s14_retrieval_full -f \
  python bench.py --steps 1 --warmup 0 --seq-images 1000 --pairs -1 --ba '' --ba-c5 '' --no-e2e --no-cpu > $O/s14_retrieval_ncu.log 2>&1
ls -la $O | tail -6
