#!/bin/bash
# Round 2, GPU session 23: the retrieval leg at the full C3 size with 4 096 descriptors per image (20.5 M descriptors).
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 150 python bench.py --seq-kp 4096 --steps 1 --warmup 0 --no-cpu --no-e2e --pairs -1 --ba '' --ba-c5 '' > gpurun_out/s23_bench_retr4096.json 2> gpurun_out/s23_bench_retr4096.err
tail -2 gpurun_out/s23_bench_retr4096.err
