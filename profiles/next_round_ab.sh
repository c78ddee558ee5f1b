#!/bin/bash
# First GPU session of the next round: confirm what was built after round 1's GPU budget ran out, then time the
# experimental kernel instances against the production ones (all flag-gated, default off).
set -x
cd ${GRAFT_REPO_ROOT:-.}
python -m pytest tests/test_zz_guided_gpu.py -m gpu -q -rxXs 2>&1 | tail -15 > gpurun_out/zz_first_run.log
for v in 0 1; do
  B2_VERIFY_VARIANT=$v B2_VERIFY_PROFILE=1 python bench.py --pairs 20000 --ba "" --no-cpu --no-e2e --steps 3 --warmup 3 \
    > gpurun_out/ab_verify_$v.json 2> gpurun_out/ab_verify_$v.err
done
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
