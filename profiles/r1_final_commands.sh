set -x
cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1
timeout 500 python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_all_r1b.csv python bench.py --pairs 20000 --verify-pairs 1200 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_all_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:verify_pairs -c 1 -f -o gpurun_out/prof_verify_r1b python bench.py --pairs 2000 --verify-pairs 1200 --ba "" --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_verify_b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:schur_kernel --launch-skip 4 -c 1 -f -o gpurun_out/prof_ba_r1b python bench.py --pairs 2000 --verify-pairs 0 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_ba_b.log 2>&1
cat gpurun_out/pytest_gpu_final.log gpurun_out/smoke_final.log; cat gpurun_out/bench_r1g.json
