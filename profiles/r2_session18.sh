#!/bin/bash
# Round 2, GPU session 18: the whole GPU suite on the final build (no -x), the two new tests with full tracebacks, the
# default line with the coalesced match fix-up.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest "tests/test_pipeline_gpu.py::test_c3_chain_retrieval_to_match_to_verify_on_one_descriptor_pool" tests/test_concurrency_gpu.py -m gpu -q > $O/s18_pytest_new.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > $O/s18_pytest_all.log 2>&1
( time timeout 1500 python bench.py > $O/s18_bench_default.json 2> $O/s18_bench_default.err ) 2> $O/s18_bench_default.time
ls -la $O | tail -5
