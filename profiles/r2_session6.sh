#!/bin/bash
# Round 2, GPU session 6: launch shapes of the verifier (A/B of 4), sampler vector in shared memory; Cholesky trace with
# phase marks; window kernel with padded shared-memory rows.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_verify_gpu.py tests/test_pipeline_gpu.py -m gpu -q -x 2>&1 | tail -5 > $O/s6_pytest.log
for v in 0 1 2 3; do
  B2_VERIFY_VARIANT=$v B2_VERIFY_PROFILE=1 timeout 300 python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 1000 --pairs -1 --ba "" --verify-pairs 20000 > $O/s6_verify_v$v.json 2> $O/s6_verify_v$v.err
done
B2_BA_CHOL_TRACE=$O/s6_chol_trace.txt timeout 300 python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 100 --seq-cand 10 --pairs -1 --ba 500,100000,10 > $O/s6_ba_c4.json 2> $O/s6_ba_c4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"verify_pairs_kernel" --launch-skip 1 -c 1 -o $O/s6_verify_full -f python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 300 --pairs -1 --ba "" > $O/s6_verify_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"schur_window_kernel" --launch-skip 4 -c 1 -o $O/s6_window_full -f python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 100 --seq-cand 10 --pairs -1 --ba 500,100000,10 > /dev/null 2>&1
ls -la $O | tail -8
