#!/bin/bash
# First GPU session of the next round: confirm what was built after round 1's GPU budget ran out, then time the
# experimental kernel instances against the production ones (all flag-gated, default off).
set -x
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_zz_guided_gpu.py tests/test_two_view_shim.py -m gpu -q -rxXs 2>&1 | tail -25 > gpurun_out/zz_first_run.log
# what the CUDA emulator cannot see (tests/cuda_emu/README.md): races between warps, kernels reading host memory. One
# pass of the never-run kernels under compute-sanitizer, small cases only (each tool slows the kernels 10-100x)
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_zz_guided_gpu.py -m gpu -q -x \
    -k "relative_pose or mean_reprojection or camera_model_is_normalised or iterative_schur_matches and kw1 or general_camera and camera2" \
    > gpurun_out/sanitizer_$tool.log 2>&1
  echo "compute-sanitizer $tool exit $?" >> gpurun_out/zz_first_run.log
done
# ITERATIVE_SCHUR (ba_iterative.cu) and relative pose (verify_pose.cu) were written after the round-1 GPU budget was spent:
# time the inner solve at 2 000 and 10 000 images (C5 shape), and take the launch list + one full capture of the matvec pair
python bench.py --pairs 2000 --verify-pairs 0 --no-cpu --no-e2e --steps 3 --warmup 3 --ba 2000,400000,10 --ba-solver iterative \
  > gpurun_out/ba_iter_2k.json 2> gpurun_out/ba_iter_2k.err
python bench.py --pairs 2000 --verify-pairs 0 --no-cpu --no-e2e --steps 3 --warmup 3 --ba 10000,2000000,10 --ba-solver iterative \
  > gpurun_out/ba_iter_10k.json 2> gpurun_out/ba_iter_10k.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/ba_iter_launches.csv \
  python bench.py --pairs 2000 --verify-pairs 0 --no-cpu --no-e2e --steps 1 --warmup 1 --ba 2000,400000,10 --ba-solver iterative > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'image_pass_kernel|matvec_point_kernel' -s 40 -c 2 \
  -o gpurun_out/ba_iter_matvec python bench.py --pairs 2000 --verify-pairs 0 --no-cpu --no-e2e --steps 1 --warmup 1 \
  --ba 2000,400000,10 --ba-solver iterative > /dev/null 2>&1
python bench.py --pairs 2000 --ba "" --verify-pairs 0 --no-cpu --no-e2e --steps 3 --warmup 3 --guided-pairs 4000 > gpurun_out/guided.json 2> gpurun_out/guided.err
python bench.py --pairs 2000 --ba "" --no-cpu --no-e2e --steps 3 --warmup 3 --verify-pose > gpurun_out/verify_pose.json 2> gpurun_out/verify_pose.err
for v in 0 1; do
  B2_VERIFY_VARIANT=$v B2_VERIFY_PROFILE=1 python bench.py --pairs 20000 --ba "" --no-cpu --no-e2e --steps 3 --warmup 3 \
    > gpurun_out/ab_verify_$v.json 2> gpurun_out/ab_verify_$v.err
done
B2_BA_CAMTERMS=image python bench.py --pairs 20000 --verify-pairs 0 --no-cpu --no-e2e --steps 3 --warmup 3 \
  > gpurun_out/ab_ba_camterms_image.json 2> gpurun_out/ab_ba_camterms_image.err
for s in atomics blocks; do
  B2_BA_SCHUR=$s python bench.py --pairs 20000 --verify-pairs 0 --no-cpu --no-e2e --steps 3 --warmup 3 \
    > gpurun_out/ab_ba_$s.json 2> gpurun_out/ab_ba_$s.err
done
python - <<'PY'
import json
for v in (0, 1):
    d = json.load(open(f"gpurun_out/ab_verify_{v}.json"))["verify"]
    print("verify variant", v, d["pairs_per_s_kernel"])
for s in ("atomics", "blocks"):
    d = json.load(open(f"gpurun_out/ab_ba_{s}.json"))["ba"]
    print("ba schur", s, d["lm_iter_per_s"], d["roofline"]["avg_ms_per_iteration"])
PY
