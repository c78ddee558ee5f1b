#!/bin/bash
# Round 2, GPU session 8: full GPU test suite incl. the new retrieval tests; the DEFAULT bench line (what the driver
# runs at round end: C3 at 5000 images) and the reference arm, both wall-timed; ncu of the retrieval kernels.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > $O/s8_gpus.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > $O/s8_pytest.log 2>&1
( time timeout 1500 python bench.py > $O/s8_bench_default.json 2> $O/s8_bench_default.err ) 2> $O/s8_bench_default.time
( time timeout 900 python bench.py --impl reference > $O/s8_bench_ref.json 2> $O/s8_bench_ref.err ) 2> $O/s8_bench_ref.time
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'word_knn|query_kernel|sort_words' -c 3 -o $O/s8_retrieval_full -f \
  python bench.py --steps 1 --warmup 0 --seq-images 1000 --pairs -1 --ba '' --no-e2e --no-cpu > $O/s8_retrieval_ncu.log 2>&1
ls -la $O | tail -12
