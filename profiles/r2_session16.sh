#!/bin/bash
# Round 2, GPU session 16: the round-end sequence on the final build -- full GPU suite, smoke(), the default bench line and
# the reference arm -- plus the opt-in legs, a launch list restricted to this library's kernels and ncu captures of the
# verifier's stage kernels and the BA kernels as they are now.
set -x
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 ) > $O/s16_pytest_all.log 2>&1
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > $O/s16_smoke.log 2>&1
( time timeout 1500 python bench.py > $O/s16_bench_default.json 2> $O/s16_bench_default.err ) 2> $O/s16_bench_default.time
( time timeout 900 python bench.py --impl reference > $O/s16_bench_ref.json 2> $O/s16_bench_ref.err ) 2> $O/s16_bench_ref.time
timeout 600 python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 300 --pairs -1 --ba '' --ba-c5 '' --retrieval-words 0 --guided-pairs 4000 --verify-pose --verify-pairs 20000 > $O/s16_bench_optin.json 2> $O/s16_bench_optin.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:'b2::' -c 400 --csv --log-file $O/s16_launches.csv \
  python bench.py --steps 1 --warmup 1 --seq-images 1000 --pairs 20000 --ba-c5 '' --no-cpu --no-e2e > $O/s16_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'verify_stage_kernel' --launch-skip 4 -c 4 -o $O/s16_verify_full -f \
  python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 300 --pairs -1 --ba "" --ba-c5 "" --retrieval-words 0 > $O/s16_verify_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k regex:'camera_terms_kernel|schur_points_kernel|schur_window_kernel|solve_graph_kernel|backsub_kernel' --launch-skip 10 -c 5 -o $O/s16_ba_full -f \
  python bench.py --no-cpu --no-e2e --steps 1 --warmup 1 --seq-images 100 --seq-cand 10 --pairs -1 --ba-c5 "" --retrieval-words 0 > $O/s16_ba_ncu.log 2>&1
ls -la $O | tail -12
