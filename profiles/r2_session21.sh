#!/bin/bash
# Round 2, GPU session 21: the retrieval GPU tests incl. the pin against the reference's vendored FLANN (golden vectors and,
# through the oracle, the live oracle/_ref/libflann_ref.so).
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_retrieval_gpu.py tests/test_oracle_retrieval.py -q > gpurun_out/s21_pytest.log 2>&1
tail -3 gpurun_out/s21_pytest.log
