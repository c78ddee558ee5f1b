#!/bin/bash
# Round 2, GPU session 9: the staged verifier (E / F / H / decision kernels) -- parity tests, A/B of the CTAs-per-SM
# of the three RANSAC stages, ncu of the H stage; retrieval tests after the top-k tie fix.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1200 python -m pytest tests/test_retrieval_gpu.py tests/test_verify_gpu.py tests/test_pipeline_gpu.py tests/test_two_view_shim.py -m gpu -q 2>&1 | tail -15 ) > $O/s9_pytest.log 2>&1
for v in 111 112 122 222 212; do
  B2_VERIFY_BPS=$v B2_VERIFY_PROFILE=1 timeout 300 python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 --seq-images 1000 --pairs -1 --ba "" --retrieval-words 0 > $O/s9_verify_$v.json 2> $O/s9_verify_$v.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"verify_stage_kernel" --launch-skip 4 -c 4 -o $O/s9_verify_full -f \
  python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 300 --pairs -1 --ba "" --retrieval-words 0 > $O/s9_verify_ncu.log 2>&1
ls -la $O | tail -12
