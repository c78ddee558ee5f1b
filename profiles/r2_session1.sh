#!/bin/bash
# Round 2, GPU session 1: state of the tree on a B200 (full -m gpu suite), the box's FP64 / FP32 / i8 peaks, and the
# first timings of everything round 1 built after its GPU budget ran out (flag-gated kernel variants, ITERATIVE_SCHUR,
# guided matching, relative pose).
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > gpurun_out/s1_gpu.txt
timeout 900 python -m pytest tests -m gpu -q -x -rxXs 2>&1 | tail -30 > gpurun_out/s1_pytest_gpu.log
timeout 120 python profiles/measure_peaks.py > gpurun_out/s1_peaks.json 2> gpurun_out/s1_peaks.err
B="python bench.py --no-cpu --no-e2e --steps 3 --warmup 3"
# ITERATIVE_SCHUR at C4, 2 000 and 10 000 images (C5 shape)
timeout 300 $B --pairs 2000 --verify-pairs 0 --ba 500,100000,10 --ba-solver iterative > gpurun_out/s1_ba_iter_500.json 2> gpurun_out/s1_ba_iter_500.err
timeout 300 $B --pairs 2000 --verify-pairs 0 --ba 2000,400000,10 --ba-solver iterative > gpurun_out/s1_ba_iter_2k.json 2> gpurun_out/s1_ba_iter_2k.err
timeout 600 $B --pairs 2000 --verify-pairs 0 --ba 10000,2000000,10 --ba-solver iterative > gpurun_out/s1_ba_iter_10k.json 2> gpurun_out/s1_ba_iter_10k.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/s1_ba_iter_launches.csv \
  $B --pairs 2000 --verify-pairs 0 --steps 1 --warmup 1 --ba 2000,400000,10 --ba-solver iterative > /dev/null 2>&1
# guided matching, relative pose
timeout 300 $B --pairs 2000 --ba "" --verify-pairs 0 --guided-pairs 4000 > gpurun_out/s1_guided.json 2> gpurun_out/s1_guided.err
timeout 300 $B --pairs 2000 --ba "" --verify-pose > gpurun_out/s1_verify_pose.json 2> gpurun_out/s1_verify_pose.err
# A/B of the flag-gated instances
for v in 0 1; do
  B2_VERIFY_VARIANT=$v B2_VERIFY_PROFILE=1 timeout 300 $B --pairs 2000 --ba "" > gpurun_out/s1_ab_verify_$v.json 2> gpurun_out/s1_ab_verify_$v.err
done
B2_BA_CAMTERMS=image timeout 300 $B --pairs 2000 --verify-pairs 0 > gpurun_out/s1_ab_ba_camterms_image.json 2> gpurun_out/s1_ab_ba_camterms_image.err
for s in atomics blocks; do
  B2_BA_SCHUR=$s timeout 300 $B --pairs 2000 --verify-pairs 0 > gpurun_out/s1_ab_ba_$s.json 2> gpurun_out/s1_ab_ba_$s.err
done
B2_BA_SCHUR=blocks B2_BA_CAMTERMS=image timeout 300 $B --pairs 2000 --verify-pairs 0 > gpurun_out/s1_ab_ba_both.json 2> gpurun_out/s1_ab_ba_both.err
ls -la gpurun_out
