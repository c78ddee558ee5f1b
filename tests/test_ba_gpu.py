"""Parity of the CUDA bundle adjuster (through the C ABI) with the FP64 oracle.

north_star tolerance: converged reprojection RMS within 1e-6 px of the reference path.  The
comparison is made at convergence (tight tolerances) because the step sequences of two LM
implementations need not coincide (hard parts)."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.ba_scene import copy_problem, make_ba_problem, reprojection_rms

pytestmark = pytest.mark.gpu
RMS_TOL = 1e-6  # px, BASELINE.json north_star


def gpu_solve(prob, **kw):
    from dagsfm_b200 import BundleAdjuster, BundleAdjustmentOptions
    o = BundleAdjustmentOptions.default()
    for k, v in kw.items():
        setattr(o, k, v)
    ba = BundleAdjuster(o)
    s = ba.Solve(prob)
    ba.close()
    return s


def test_structure_counts_and_constant_blocks():
    # bundle_adjustment_test.cc:186-233 (400 / 309) and the CheckConstant* macros (:41-107)
    prob = make_ba_problem(n_img=2, n_pts=100, track_len=2, seed=1)
    s = gpu_solve(prob, max_num_iterations=2)
    assert s.num_residuals_reduced == 400 and s.num_effective_parameters_reduced == 309
    prob = make_ba_problem(n_img=6, n_pts=120, track_len=4, seed=2, n_const_pts=15)
    prob["cam_const"][2] = 1
    before = copy_problem(prob)
    gpu_solve(prob, max_num_iterations=10)
    assert (prob["qvec"][0] == before["qvec"][0] / np.linalg.norm(before["qvec"][0])).all()
    assert (prob["tvec"][0] == before["tvec"][0]).all()
    assert prob["tvec"][1][0] == before["tvec"][1][0] and (prob["tvec"][1][1:] != before["tvec"][1][1:]).all()
    c = before["pt_const"].astype(bool)
    assert (prob["xyz"][c] == before["xyz"][c]).all() and (prob["xyz"][~c] != before["xyz"][~c]).any()
    assert (prob["cam_params"][2] == before["cam_params"][2]).all()
    assert (prob["cam_params"][:, 1:3] == before["cam_params"][:, 1:3]).all()


@pytest.mark.parametrize("kw", [
    dict(n_img=8, n_pts=150, track_len=4, seed=5),
    dict(n_img=24, n_pts=600, track_len=6, seed=3),
    dict(n_img=12, n_pts=300, track_len=5, seed=4, shared_camera=True),
    dict(n_img=16, n_pts=400, track_len=40, seed=6),          # tracks longer than one Schur tile
    dict(n_img=30, n_pts=500, track_len=7, seed=7, n_const_pts=60),
])
def test_converged_rms_matches_oracle(kw):
    if kw.get("track_len", 0) > kw["n_img"]:
        kw = dict(kw, n_img=48)
    p_gpu = make_ba_problem(**kw)
    p_cpu = copy_problem(p_gpu)
    tight = dict(max_num_iterations=200, gradient_tolerance=1e-9, function_tolerance=1e-16)
    s_gpu = gpu_solve(p_gpu, **tight)
    s_cpu = orc.ba_solve(p_cpu, **tight)
    rms_gpu, rms_cpu = reprojection_rms(p_gpu), reprojection_rms(p_cpu)
    print(f"\n{kw}: rms gpu {rms_gpu:.9f} cpu {rms_cpu:.9f}  iters {s_gpu.num_iterations}/{s_cpu.num_successful_steps + s_cpu.num_unsuccessful_steps}")
    assert abs(rms_gpu - rms_cpu) < RMS_TOL
    assert abs(np.sqrt(2 * s_gpu.final_cost / len(p_gpu["obs_img"])) - rms_gpu) < 1e-9
    assert s_gpu.initial_cost == pytest.approx(s_cpu.initial_cost, rel=1e-12)
    # same minimum: parameters agree far below the noise level
    assert np.abs(p_gpu["xyz"] - p_cpu["xyz"]).max() < 1e-5
    assert np.abs(p_gpu["cam_params"] - p_cpu["cam_params"]).max() < 1e-3


def test_default_options_follow_the_same_path_as_oracle():
    # reference final-BA options (gradient_tolerance 1.0, 50 iterations): same accept / reject
    # sequence and the same stopping iterate as the oracle restatement
    p_gpu = make_ba_problem(n_img=40, n_pts=2000, track_len=8, seed=11)
    p_cpu = copy_problem(p_gpu)
    s_gpu = gpu_solve(p_gpu)
    s_cpu = orc.ba_solve(p_cpu)
    assert s_gpu.num_successful_steps == s_cpu.num_successful_steps
    assert s_gpu.num_unsuccessful_steps == s_cpu.num_unsuccessful_steps
    assert s_gpu.termination_type == s_cpu.termination
    assert abs(reprojection_rms(p_gpu) - reprojection_rms(p_cpu)) < RMS_TOL
    assert s_gpu.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-9)


def test_invalid_problem_is_rejected():
    from dagsfm_b200 import B2Error
    prob = make_ba_problem(n_img=4, n_pts=20, track_len=3, seed=1)
    prob["obs_pt"] = np.ascontiguousarray(prob["obs_pt"][::-1])
    with pytest.raises(B2Error):
        gpu_solve(prob)


# ---------------------------------------------------------------- fused exact path (ba_fused.cu + ba_chol.cu)
def test_tiled_cholesky_seam_on_device():
    """The linear solver of the exact Schur step is the library's own tiled Cholesky (no cuSOLVER): dense and banded
    SPD systems against LAPACK on the host, and a non-positive pivot is reported."""
    from dagsfm_b200 import BundleAdjuster
    from tests.test_emu_ba import spd_cases
    adj = BundleAdjuster()
    try:
        rng = np.random.default_rng(9)
        M = rng.normal(size=(2000, 2000))
        big = (M @ M.T + 2000 * np.eye(2000), rng.normal(size=2000), None)
        for A, b, band in spd_cases() + [big]:
            x, info, n_tiles = adj.debug_cholesky_solve(A, b)
            nt = (len(b) + 63) // 64
            assert info == 0
            assert n_tiles == nt * (nt + 1) // 2 if band is None else n_tiles < nt * (nt + 1) // 2
            ref = np.linalg.solve(A, b)
            assert np.abs(x - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()) * np.linalg.cond(A)
        A = np.eye(70)
        A[3, 3] = -1.0
        assert adj.debug_cholesky_solve(A, np.ones(70))[1] != 0
    finally:
        adj.close()


@pytest.mark.parametrize("kw", [
    dict(n_img=8, n_pts=150, track_len=4, seed=5),
    dict(n_img=12, n_pts=300, track_len=5, seed=4, shared_camera=True),
    dict(n_img=60, n_pts=3000, track_len=12, seed=8),
    dict(n_img=60, n_pts=1500, track_len=16, seed=2),
    dict(n_img=30, n_pts=500, track_len=7, seed=7, n_const_pts=60),
])
def test_fused_and_staged_exact_paths_agree_on_device(kw, monkeypatch):
    p_f, p_s, p_cpu = make_ba_problem(**kw), make_ba_problem(**kw), make_ba_problem(**kw)
    s_f = gpu_solve(p_f)
    monkeypatch.setenv("B2_BA_EXACT", "staged")
    s_s = gpu_solve(p_s)
    monkeypatch.delenv("B2_BA_EXACT")
    s_cpu = orc.ba_solve(p_cpu)
    assert (s_f.exact_path_used, s_s.exact_path_used) == (2, 1)
    for s in (s_f, s_s):
        assert (s.num_successful_steps, s.num_unsuccessful_steps, s.termination_type) == \
               (s_cpu.num_successful_steps, s_cpu.num_unsuccessful_steps, s_cpu.termination)
        assert s.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-9)
    assert abs(reprojection_rms(p_f) - reprojection_rms(p_cpu)) < RMS_TOL
    assert np.abs(p_f["xyz"] - p_s["xyz"]).max() < 1e-7


@pytest.mark.timeout(900)
def test_c4_shape_rms_within_tolerance_of_the_fp64_oracle():
    """BASELINE configs[3] at full size (500 cams / 100k pts / 1M obs), reference final-BA options: the GPU solve and the
    FP64 oracle (all host threads) stop at the same iterate, reprojection RMS within 1e-6 px (north_star)."""
    kw = dict(n_img=500, n_pts=100000, track_len=10, seed=1)
    p_gpu, p_cpu = make_ba_problem(**kw), make_ba_problem(**kw)
    s_gpu = gpu_solve(p_gpu)
    assert s_gpu.exact_path_used == 2
    s_cpu = orc.ba_solve(p_cpu)
    rms_gpu, rms_cpu = reprojection_rms(p_gpu), reprojection_rms(p_cpu)
    print(f"\nC4: rms gpu {rms_gpu:.12f} cpu {rms_cpu:.12f} |d| {abs(rms_gpu - rms_cpu):.3e}; steps "
          f"{s_gpu.num_successful_steps}+{s_gpu.num_unsuccessful_steps} / {s_cpu.num_successful_steps}+{s_cpu.num_unsuccessful_steps}")
    assert abs(rms_gpu - rms_cpu) < RMS_TOL
    assert (s_gpu.num_successful_steps, s_gpu.num_unsuccessful_steps, s_gpu.termination_type) == \
           (s_cpu.num_successful_steps, s_cpu.num_unsuccessful_steps, s_cpu.termination)
    assert s_gpu.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-9)


def test_gradient_max_norm_goes_through_the_quaternion_plus_on_device():
    from tests.test_emu_ba import rotation_only_problem
    for solver in (1, 2):
        s = gpu_solve(rotation_only_problem(), gradient_tolerance=2.0, linear_solver_type=solver)
        assert (s.num_iterations, s.termination_type) == (0, 0)
        p = rotation_only_problem()
        p["tvec_const"][:] = 0
        assert gpu_solve(p, gradient_tolerance=2.0, linear_solver_type=solver).num_iterations > 0
