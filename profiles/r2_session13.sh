#!/bin/bash
# Round 2, GPU session 13: shuffle-free 9x9 Jacobi sums in the verifier, coalesced uploads, C5 BA leg in the default line,
# smoke().
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/s13_smoke.log 2>&1
( time timeout 900 python -m pytest tests/test_verify_gpu.py tests/test_pipeline_gpu.py tests/test_two_view_shim.py tests/test_match_gpu.py tests/test_retrieval_gpu.py -m gpu -q 2>&1 | tail -8 ) > $O/s13_pytest.log 2>&1
B2_VERIFY_PROFILE=1 timeout 300 python bench.py --no-cpu --no-e2e --steps 2 --warmup 1 --seq-images 1000 --pairs -1 --ba "" --ba-c5 "" --retrieval-words 0 > $O/s13_verify_1k.json 2> $O/s13_verify_1k.err
( time timeout 1500 python bench.py > $O/s13_bench_default.json 2> $O/s13_bench_default.err ) 2> $O/s13_bench_default.time
ls -la $O | tail -6
