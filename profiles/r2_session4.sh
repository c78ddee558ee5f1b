#!/bin/bash
# Round 2, GPU session 4: verifier with compact call graph + division-free scoring (A/B against the reference-expression
# instance), Cholesky task trace, first run of the C3 pipeline bench (reduced size) and of the SiftFeatureMatcher test.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_verify_gpu.py tests/test_pipeline_gpu.py tests/test_two_view_shim.py -m gpu -q -x --durations=5 2>&1 | tail -15 > $O/s4_pytest.log
B="python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 --seq-images 200 --pairs -1 --ba ''"
for v in 0 1; do
  B2_VERIFY_VARIANT=$v B2_VERIFY_PROFILE=1 timeout 300 python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 100 --seq-cand 10 --pairs -1 --ba "" --verify-pairs 20000 > $O/s4_verify_v$v.json 2> $O/s4_verify_v$v.err
done
B2_BA_CHOL_TRACE=$O/s4_chol_trace.txt timeout 300 python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 100 --seq-cand 10 --pairs -1 --ba 500,100000,10 > $O/s4_ba_c4.json 2> $O/s4_ba_c4.err
timeout 900 python bench.py --steps 2 --warmup 1 --seq-images 1000 --pairs 20000 --ba "" > $O/s4_pipeline_1k.json 2> $O/s4_pipeline_1k.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"verify_pairs_kernel" --launch-skip 1 -c 1 -o $O/s4_verify_full -f python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 100 --seq-cand 10 --pairs -1 --ba "" --verify-pairs 6000 > $O/s4_verify_ncu.log 2>&1
ls -la $O | tail -12
