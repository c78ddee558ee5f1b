#!/bin/bash
# Round 2, GPU session 3: task-graph Cholesky (ba_chol.cu v2) -- parity, timing, ncu; source-level capture of the verifier.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_ba_gpu.py tests/test_ba_shim.py tests/test_pba_shim.py -m gpu -q -x --durations=5 2>&1 | tail -15 > $O/s3_pytest_ba.log
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 --images 64 --pairs 2000 --verify-pairs 0"
timeout 300 $B --ba 500,100000,10 > $O/s3_ba_c4.json 2> $O/s3_ba_c4.err
timeout 300 $B --ba 1000,200000,10 --ba-solver exact > $O/s3_ba_1k_exact.json 2> $O/s3_ba_1k_exact.err
timeout 300 $B --ba 2000,400000,10 --ba-solver exact > $O/s3_ba_2k_exact.json 2> $O/s3_ba_2k_exact.err
N="python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --images 64 --pairs 2000 --verify-pairs 0 --ba 500,100000,10"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/s3_ba_c4_launches.csv $N > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"solve_graph_kernel" --launch-skip 3 -c 1 -o $O/s3_ba_chol_full -f $N > /dev/null 2>&1
V="python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --images 64 --pairs 2000 --ba '' --verify-pairs 6000"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"verify_pairs_kernel" --launch-skip 1 -c 1 -o $O/s3_verify_full -f python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --images 64 --pairs 2000 --ba "" --verify-pairs 6000 > $O/s3_verify_ncu.log 2>&1
ls -la $O | tail -12
