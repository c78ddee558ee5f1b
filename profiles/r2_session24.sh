#!/bin/bash
# Round 2, GPU session 24: a miniature of the default command, to see the final bench.py print its line end to end.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
timeout 75 python bench.py --seq-images 120 --seq-kp 512 --seq-cand 10 --steps 1 --warmup 1 --pairs 500 --images 40 --desc 512 --ba 30,3000,5 --ba-c5 '' --retrieval-words 512 --cpu-sample 16 > gpurun_out/s24_bench_mini.json 2> gpurun_out/s24_bench_mini.err
echo rc=$?
tail -2 gpurun_out/s24_bench_mini.err
