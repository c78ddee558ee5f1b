#!/bin/bash
# Round 2, GPU session 10: retrieval with the tcgen05 word search (tests, bench leg, ncu); C1 mode; the default bench line
# with the staged verifier; launch list of the default command for profiles/.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_retrieval_gpu.py -m gpu -q 2>&1 | tail -15 ) > $O/s10_pytest_retrieval.log 2>&1
( time timeout 600 python bench.py --c1 --steps 3 --warmup 3 > $O/s10_bench_c1.json 2> $O/s10_bench_c1.err ) 2> $O/s10_bench_c1.time
( time timeout 1500 python bench.py > $O/s10_bench_default.json 2> $O/s10_bench_default.err ) 2> $O/s10_bench_default.time
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'word_knn_tc|query_kernel|project_kernel' -c 3 -o $O/s10_retrieval_full -f \
  python bench.py --steps 1 --warmup 0 --seq-images 1000 --pairs -1 --ba '' --no-e2e --no-cpu > $O/s10_retrieval_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/s10_launches.csv \
  python bench.py --steps 1 --warmup 1 --seq-images 1000 --pairs 20000 --no-cpu > $O/s10_launches_bench.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > $O/s10_pytest_all.log 2>&1
ls -la $O | tail -12
