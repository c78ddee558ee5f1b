#!/bin/bash
# Round 2, GPU session 11: pinned / double-buffered output path of the match -> verify chain, 65 536-pair chunks, the
# reworked tcgen05 word-search epilogue; ncu of the word search and the query kernel.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_retrieval_gpu.py tests/test_pipeline_gpu.py -m gpu -q 2>&1 | tail -15 ) > $O/s11_pytest.log 2>&1
( time timeout 1500 python bench.py > $O/s11_bench_default.json 2> $O/s11_bench_default.err ) 2> $O/s11_bench_default.time
timeout 600 python bench.py --chunk-pairs 16384 --no-cpu --no-e2e --pairs -1 --ba '' --retrieval-words 0 > $O/s11_bench_chunk16k.json 2> $O/s11_bench_chunk16k.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'word_knn_tc|query_kernel' -c 2 -o $O/s11_retrieval_full -f \
  python bench.py --steps 1 --warmup 0 --seq-images 1000 --pairs -1 --ba '' --no-e2e --no-cpu > $O/s11_retrieval_ncu.log 2>&1
ls -la $O | tail -8
