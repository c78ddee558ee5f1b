#!/bin/bash
# Round 2, GPU session 22: the C3 pipeline with SURVEY 8d's literal image size (4 096 keypoints per image instead of the
# default line's 2 048), device-resident and end to end.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 230 python bench.py --seq-kp 4096 --steps 2 --warmup 1 --no-cpu --pairs -1 --ba '' --ba-c5 '' --retrieval-words 0 > gpurun_out/s22_bench_kp4096.json 2> gpurun_out/s22_bench_kp4096.err
tail -2 gpurun_out/s22_bench_kp4096.err
