#!/bin/bash
# Round 2, GPU session 5: verifier at 2 CTAs/SM (lane workspace in local memory, 128 registers) vs 1 CTA/SM; prefetched
# scoring loop; blocked 4-pivot factorisation in the Cholesky task graph (trace + timing).
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_verify_gpu.py tests/test_pipeline_gpu.py tests/test_ba_gpu.py -m gpu -q -x --durations=5 2>&1 | tail -15 > $O/s5_pytest.log
for v in 0 1; do
  B2_VERIFY_VARIANT=$v B2_VERIFY_PROFILE=1 timeout 300 python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 1000 --pairs -1 --ba "" --verify-pairs 20000 > $O/s5_verify_v$v.json 2> $O/s5_verify_v$v.err
done
B2_BA_CHOL_TRACE=$O/s5_chol_trace.txt timeout 300 python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 100 --seq-cand 10 --pairs -1 --ba 500,100000,10 > $O/s5_ba_c4.json 2> $O/s5_ba_c4.err
timeout 300 python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 100 --seq-cand 10 --pairs -1 --ba 2000,400000,10 --ba-solver exact > $O/s5_ba_2k.json 2> $O/s5_ba_2k.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"verify_pairs_kernel" --launch-skip 1 -c 1 -o $O/s5_verify_full -f python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 300 --pairs -1 --ba "" > $O/s5_verify_ncu.log 2>&1
ls -la $O | tail -8
