"""Host-side helpers of bench.py that can be checked without a GPU."""
import json

import numpy as np

import bench
from dagsfm_b200.verification import RESULT_DTYPE


def test_verify_work_rates_counts_hypothesis_match_evaluations():
    res = np.zeros(3, dtype=RESULT_DTYPE)
    res["E_num_trials"], res["F_num_trials"], res["H_num_trials"] = [10, 0, 5], [100, 200, 0], [1000, 2000, 30]
    r = bench.verify_work_rates(res, [0, 400, 900, 1000], 0.5)
    ef, h = 110 * 400 + 200 * 500 + 5 * 100, 1000 * 400 + 2000 * 500 + 30 * 100
    assert r["hypothesis_match_evals_per_s_min"] == (ef + h) / 0.5
    assert abs(r["fp64_gflops_min"] - (33 * ef + 25 * h) / 0.5 / 1e9) < 1e-12
    json.dumps(r)   # must be serialisable: it goes into the bench line


def test_effective_cores_is_positive_and_bounded():
    import os
    n = bench.effective_cores()
    assert 1 <= n <= (os.cpu_count() or 1)


def test_all_pairs_enumerates_the_upper_triangle():
    p = bench.all_pairs(5)
    assert p.shape == (10, 2) and (p[:, 0] < p[:, 1]).all() and len({tuple(x) for x in p}) == 10
    assert bench.all_pairs(5, limit=3).shape == (3, 2)
