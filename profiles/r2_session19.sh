#!/bin/bash
# Round 2, GPU session 19: the C3-chain test after its seed fix; launch list of the match kernels with the coalesced fix-up.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_pipeline_gpu.py tests/test_match_gpu.py -m gpu -q > $O/s19_pytest.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'match_|fill_items|pair_items' -c 120 --csv --log-file $O/s19_match_launches.csv \
  python bench.py --steps 1 --warmup 1 --seq-images 1000 --pairs 20000 --ba '' --ba-c5 '' --retrieval-words 0 --no-cpu --no-e2e > $O/s19_launches_bench.log 2>&1
tail -3 $O/s19_pytest.log
