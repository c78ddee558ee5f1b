#!/bin/bash
# Round 2, GPU session 2: the rebuilt exact BA step (ba_fused.cu + ba_chol.cu) -- parity suite, C4 timing against the
# staged path, launch list + ncu --set full of the new kernels; verification identity after the device-order oracle.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -rxXs --durations=15 2>&1 | tail -45 > $O/s2_pytest_gpu.log
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 3 --images 64 --pairs 2000 --verify-pairs 0"
timeout 300 $B --ba 500,100000,10 > $O/s2_ba_c4_fused.json 2> $O/s2_ba_c4_fused.err
B2_BA_EXACT=staged timeout 300 $B --ba 500,100000,10 > $O/s2_ba_c4_staged.json 2> $O/s2_ba_c4_staged.err
timeout 300 $B --ba 2000,400000,10 --ba-solver exact > $O/s2_ba_2k_exact.json 2> $O/s2_ba_2k_exact.err
timeout 300 $B --ba 1000,200000,10 --ba-solver exact > $O/s2_ba_1k_exact.json 2> $O/s2_ba_1k_exact.err
N="python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --images 64 --pairs 2000 --verify-pairs 0 --ba 500,100000,10"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/s2_ba_c4_launches.csv $N > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"camera_terms_kernel|schur_points_kernel|schur_window_kernel|backsub_kernel" --launch-skip 8 -c 4 -o $O/s2_ba_fused_full -f $N > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"update_kernel|panel_kernel" --launch-skip 40 -c 2 -o $O/s2_ba_chol_full -f $N > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"solve_kernel" --launch-skip 2 -c 1 -o $O/s2_ba_solve_full -f $N > /dev/null 2>&1
ls -la $O
