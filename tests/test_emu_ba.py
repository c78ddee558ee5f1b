"""The bundle adjuster's CUDA sources (ba_kernels.cu, ba_api.cu, common.cu) compiled for the HOST against
tests/cuda_emu/cuda_emu.h -- kernels run as fibers, block by block, cuSOLVER replaced by a plain Cholesky --
and driven through the same C ABI and Python wrapper as on the GPU.  Checks the device code paths that the
GPU suite covers (and the ones added after the last GPU session: robust losses, reduced-residual counts)
against the oracle without a GPU.  This is a TEST of the CUDA code, not a fallback: the product library
(dagsfm_b200/libdagsfm_b200.so) is not involved and still refuses to run without a device."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.ba_scene import copy_problem, make_ba_problem, reprojection_rms


@pytest.fixture(scope="module")
def emu(request):
    from tests.cuda_emu.build_emu import BA_SOURCES, VERIFY_SOURCES, build
    import dagsfm_b200.bundle_adjustment as ba
    L = C.CDLL(str(build("ba", BA_SOURCES)))
    vp, P = C.c_void_p, C.POINTER
    L.b2_ba_default_options.argtypes = [P(ba.BundleAdjustmentOptions)]
    L.b2_ba_default_options.restype = None
    L.b2_ba_create.argtypes = [C.c_int, P(vp)]
    L.b2_ba_destroy.argtypes = [vp]
    L.b2_ba_set_allreduce.argtypes = [vp, ba.ALLREDUCE_FN, vp]
    L.b2_ba_solve.argtypes = [vp, P(ba.BaProblem), P(ba.BundleAdjustmentOptions), P(ba.BaSummary)]
    L.b2_ba_debug_cholesky_solve.argtypes = [vp, C.c_int64, vp, vp, vp, P(C.c_int32), P(C.c_int32)]
    L.b2_last_error.restype = C.c_char_p
    saved = (ba._L, ba.check)

    def check(rc):
        if rc != 0:
            raise RuntimeError(f"emulated library error {rc}: {L.b2_last_error().decode()}")
    ba._L, ba.check = (lambda: L), check
    yield ba
    ba._L, ba.check = saved


def emu_solve(ba, prob, **kw):
    o = ba.BundleAdjustmentOptions()
    ba._L().b2_ba_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    adj = ba.BundleAdjuster(o)
    try:
        return adj.Solve(prob)
    finally:
        adj.close()


TIGHT = dict(max_num_iterations=200, gradient_tolerance=1e-9, function_tolerance=1e-16)


@pytest.mark.parametrize("kw", [
    dict(n_img=6, n_pts=60, track_len=4, seed=5),
    dict(n_img=8, n_pts=80, track_len=5, seed=3, shared_camera=True),
    dict(n_img=40, n_pts=30, track_len=36, seed=6),            # tracks longer than one Schur tile
    dict(n_img=8, n_pts=90, track_len=4, seed=7, n_const_pts=20),
])
def test_emulated_kernels_converge_to_the_oracle_optimum(emu, kw):
    p_dev = make_ba_problem(**kw)
    p_cpu = copy_problem(p_dev)
    s_dev = emu_solve(emu, p_dev, **TIGHT)
    s_cpu = orc.ba_solve(p_cpu, **TIGHT)
    assert abs(reprojection_rms(p_dev) - reprojection_rms(p_cpu)) < 1e-6
    assert s_dev.initial_cost == pytest.approx(s_cpu.initial_cost, rel=1e-12)
    assert s_dev.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-9)
    assert np.abs(p_dev["xyz"] - p_cpu["xyz"]).max() < 1e-5
    assert (s_dev.num_residuals_reduced, s_dev.num_effective_parameters_reduced) == (s_cpu.num_residuals, s_cpu.num_effective_parameters)


def test_default_options_follow_the_oracle_path(emu):
    p_dev = make_ba_problem(n_img=10, n_pts=150, track_len=5, seed=11)
    p_cpu = copy_problem(p_dev)
    s_dev, s_cpu = emu_solve(emu, p_dev), orc.ba_solve(p_cpu)
    assert (s_dev.num_successful_steps, s_dev.num_unsuccessful_steps, s_dev.termination_type) == \
           (s_cpu.num_successful_steps, s_cpu.num_unsuccessful_steps, s_cpu.termination)
    assert s_dev.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-9)


@pytest.mark.parametrize("loss_type,scale", [(1, 1.0), (2, 1.0), (1, 2.5)])
def test_robust_loss_kernel_instances(emu, loss_type, scale):
    p_dev = make_ba_problem(n_img=8, n_pts=120, track_len=5, seed=4, noise_px=1.0)
    rng = np.random.default_rng(1)
    idx = rng.choice(len(p_dev["obs_xy"]), 30, replace=False)
    p_dev["obs_xy"][idx] += rng.normal(0, 40, (30, 2))
    p_cpu = copy_problem(p_dev)
    kw = dict(max_num_iterations=200, gradient_tolerance=1e-9, function_tolerance=1e-12)
    s_dev = emu_solve(emu, p_dev, loss_function_type=loss_type, loss_function_scale=scale, **kw)
    s_cpu = orc.ba_solve(p_cpu, loss_type=loss_type, loss_scale=scale, **kw)
    assert s_dev.initial_cost == pytest.approx(s_cpu.initial_cost, rel=1e-12)
    assert s_dev.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-9)
    assert abs(reprojection_rms(p_dev) - reprojection_rms(p_cpu)) < 1e-6


def test_reference_config_case_with_dropped_residual_blocks(emu):
    from dagsfm_b200.ba_config import BundleAdjustmentConfig, pack_problem
    from tests.test_ba_config import generate_reconstruction
    r = generate_reconstruction(3, 40)
    vp, cp = r.images[2]["points2D"][1][2], r.images[2]["points2D"][2][2]
    r.delete_observation(2, 0)
    c = BundleAdjustmentConfig()
    c.AddImage(0); c.AddImage(1); c.SetConstantPose(0); c.SetConstantPose(1)
    c.AddVariablePoint(vp); c.AddConstantPoint(cp)
    p_dev, _ = pack_problem(r, c)
    p_cpu = {k: v.copy() for k, v in p_dev.items()}
    kw = dict(max_num_iterations=100, function_tolerance=0.0, gradient_tolerance=0.0, parameter_tolerance=0.0)
    s_dev = emu_solve(emu, p_dev, **kw)
    s_cpu = orc.ba_solve(p_cpu, **kw)
    assert (s_dev.num_residuals_reduced, s_dev.num_effective_parameters_reduced) == (2 * 2 * 40 + 2, 10)
    assert (s_cpu.num_residuals, s_cpu.num_effective_parameters) == (162, 10)
    assert s_dev.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-9)
    for k, const in (("qvec", "pose_const"), ("cam_params", "cam_const"), ("xyz", "pt_const")):
        assert (p_dev[k][p_dev[const] == 1] == p_cpu[k][p_cpu[const] == 1]).all()


def test_invalid_loss_is_rejected(emu):
    p = make_ba_problem(n_img=4, n_pts=20, track_len=3, seed=1)
    with pytest.raises(RuntimeError):
        emu_solve(emu, p, loss_function_type=5)
    with pytest.raises(RuntimeError):
        emu_solve(emu, p, loss_function_scale=0.0)


@pytest.mark.parametrize("kw", [
    dict(n_img=6, n_pts=60, track_len=4, seed=5),
    dict(n_img=8, n_pts=80, track_len=5, seed=3, shared_camera=True),      # intrinsics columns shared by all blocks
    dict(n_img=40, n_pts=30, track_len=36, seed=6),
    dict(n_img=8, n_pts=90, track_len=4, seed=7, n_const_pts=20),
])
def test_pair_major_schur_variant(emu, kw, monkeypatch):
    """B2_BA_SCHUR=blocks: counting-sorted (image, image) blocks, W / Y per observation, one atomic per block
    entry.  Must follow the same LM path as the production kernel and the oracle."""
    p_pm = make_ba_problem(**kw)
    p_ref, p_cpu = copy_problem(p_pm), copy_problem(p_pm)
    s_ref = emu_solve(emu, p_ref, **TIGHT)
    monkeypatch.setenv("B2_BA_SCHUR", "blocks")
    s_pm = emu_solve(emu, p_pm, **TIGHT)
    monkeypatch.delenv("B2_BA_SCHUR")
    s_cpu = orc.ba_solve(p_cpu, **TIGHT)
    assert s_pm.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-9) and s_pm.final_cost == pytest.approx(s_ref.final_cost, rel=1e-9)
    assert abs(reprojection_rms(p_pm) - reprojection_rms(p_cpu)) < 1e-6
    assert np.abs(p_pm["xyz"] - p_ref["xyz"]).max() < 1e-5
    # with the reference's own stopping rule (gradient_tolerance 1.0) the accept / reject sequence is the oracle's;
    # at the 1e-16 noise floor above it depends on the summation order inside S
    q_pm = make_ba_problem(**kw)
    q_cpu = copy_problem(q_pm)
    monkeypatch.setenv("B2_BA_SCHUR", "blocks")
    t_pm = emu_solve(emu, q_pm)
    monkeypatch.delenv("B2_BA_SCHUR")
    t_cpu = orc.ba_solve(q_cpu)
    assert (t_pm.num_successful_steps, t_pm.num_unsuccessful_steps, t_pm.termination_type) == \
           (t_cpu.num_successful_steps, t_cpu.num_unsuccessful_steps, t_cpu.termination)
    assert t_pm.final_cost == pytest.approx(t_cpu.final_cost, rel=1e-9)


def test_pair_major_with_two_observations_of_one_image_in_a_track(emu, monkeypatch):
    p_pm = make_ba_problem(n_img=7, n_pts=70, track_len=4, seed=9)
    for pt in range(0, 70, 5):                       # point pt is seen twice by the image of its first observation
        o = np.searchsorted(p_pm["obs_pt"], pt)
        p_pm["obs_img"][o + 1] = p_pm["obs_img"][o]
        p_pm["obs_xy"][o + 1] = p_pm["obs_xy"][o] + 0.3
    p_ref, p_cpu = copy_problem(p_pm), copy_problem(p_pm)
    s_ref = emu_solve(emu, p_ref, **TIGHT)
    monkeypatch.setenv("B2_BA_SCHUR", "blocks")
    s_pm = emu_solve(emu, p_pm, **TIGHT)
    monkeypatch.delenv("B2_BA_SCHUR")
    s_cpu = orc.ba_solve(p_cpu, **TIGHT)
    assert s_ref.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-9)
    assert s_pm.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-9)
    assert np.abs(p_pm["xyz"] - p_cpu["xyz"]).max() < 1e-5


def test_edge_cases_nothing_to_optimise(emu):
    p = make_ba_problem(n_img=4, n_pts=20, track_len=3, seed=1)
    q = copy_problem(p)
    q["pose_const"][:] = 1; q["cam_const"][:] = 1; q["pt_const"][:] = 1          # everything constant
    before = {k: q[k].copy() for k in ("qvec", "tvec", "cam_params", "xyz")}
    s = emu_solve(emu, q)
    assert s.num_effective_parameters_reduced == 0 and s.num_residuals_reduced == 0
    assert all((q[k] == before[k]).all() for k in ("tvec", "cam_params", "xyz"))
    assert np.abs(q["qvec"] - before["qvec"]).max() < 1e-15          # image.NormalizeQvec() (bundle_adjustment.cc:345) still runs
    e = copy_problem(p)                                                           # no observations at all
    for k in ("obs_img", "obs_pt"):
        e[k] = e[k][:0].copy()
    e["obs_xy"] = e["obs_xy"][:0].copy()
    s = emu_solve(emu, e)
    assert s.num_residuals_reduced == 0 and s.initial_cost == 0.0
    # every camera-side block constant with the EXACT solver (reduced system of size 0: only the points move), then a
    # single free intrinsics block and nothing else on the camera side (reduced system of size 2)
    for free_cam in (False, True):
        d0 = make_ba_problem(n_img=5, n_pts=40, track_len=3, seed=3, shared_camera=True)
        d0["pose_const"][:] = 1
        d0["cam_const"][:] = 0 if free_cam else 1
        dc = copy_problem(d0)
        s, sc = emu_solve(emu, d0, linear_solver_type=1), orc.ba_solve(dc, linear_solver=0)
        assert s.num_effective_parameters_reduced == sc.num_effective_parameters == 3 * 40 + (2 if free_cam else 0)
        assert (s.num_successful_steps, s.num_unsuccessful_steps) == (sc.num_successful_steps, sc.num_unsuccessful_steps)
        assert s.final_cost == pytest.approx(sc.final_cost, rel=1e-9) and s.final_cost < s.initial_cost
        assert np.abs(d0["xyz"] - dc["xyz"]).max() < 1e-7 and (d0["tvec"] == dc["tvec"]).all()
    bad = copy_problem(p)
    bad["obs_img"][0] = 99                                                        # observation of an unknown image
    with pytest.raises(RuntimeError):
        emu_solve(emu, bad)
    unsorted = copy_problem(p)
    unsorted["obs_pt"][0], unsorted["obs_pt"][-1] = unsorted["obs_pt"][-1], unsorted["obs_pt"][0]
    with pytest.raises(RuntimeError):                                             # observations must be sorted by point
        emu_solve(emu, unsorted)


# ---------------------------------------------------------------- ITERATIVE_SCHUR (ba_iterative.cu)
ITER = dict(linear_solver_type=2)


@pytest.mark.parametrize("kw", [
    dict(n_img=6, n_pts=60, track_len=4, seed=5),
    dict(n_img=8, n_pts=80, track_len=5, seed=3, shared_camera=True),   # one intrinsics block shared by all images
    dict(n_img=14, n_pts=20, track_len=11, seed=6),                      # more observations per point than lanes per point
    dict(n_img=8, n_pts=90, track_len=4, seed=7, n_const_pts=20),
    dict(n_img=150, n_pts=40, track_len=3, seed=8),                      # images with zero / one observation
])
def test_iterative_schur_kernels_converge_to_the_oracle_optimum(emu, kw):
    p_dev = make_ba_problem(**kw)
    p_it, p_ex = copy_problem(p_dev), copy_problem(p_dev)
    s_dev = emu_solve(emu, p_dev, **TIGHT, **ITER)
    s_it = orc.ba_solve(p_it, linear_solver=1, **TIGHT)
    s_ex = orc.ba_solve(p_ex, **TIGHT)
    assert s_dev.linear_solver_type_used == 2 and s_dev.num_linear_solver_iterations > 0
    assert s_dev.initial_cost == pytest.approx(s_it.initial_cost, rel=1e-12)
    for ref in (p_it, p_ex):   # the inexact-step oracle and the exact-step oracle share the optimum
        assert abs(reprojection_rms(p_dev) - reprojection_rms(ref)) < 1e-6
    assert s_dev.final_cost == pytest.approx(s_it.final_cost, rel=1e-9)
    assert np.abs(p_dev["xyz"] - p_it["xyz"]).max() < 1e-5
    assert (s_dev.num_residuals_reduced, s_dev.num_effective_parameters_reduced) == (s_it.num_residuals, s_it.num_effective_parameters)


@pytest.mark.parametrize("kw", [
    dict(n_img=10, n_pts=150, track_len=5, seed=11),
    dict(n_img=30, n_pts=400, track_len=6, seed=12),
    dict(n_img=12, n_pts=200, track_len=5, seed=13, shared_camera=True),
])
def test_iterative_schur_follows_the_oracle_iteration_by_iteration(emu, kw):
    """Default (final-BA) options.  Over the first LM steps the q-tolerance test of the conjugate-gradient loop fires
    at exactly the same inner iteration as in the oracle's restatement of Ceres' loop; over the whole solve the LM
    path -- accepted / rejected steps, termination -- is the same and the inner-iteration total stays close (the
    test `zeta < 0.1` is discontinuous, so after several hundred inner iterations rounding may move one firing)."""
    p_dev = make_ba_problem(**kw)
    p_cpu = copy_problem(p_dev)
    s_dev = emu_solve(emu, p_dev, max_num_iterations=3, **ITER)
    s_cpu = orc.ba_solve(p_cpu, linear_solver=1, max_num_iterations=3)
    assert s_dev.num_linear_solver_iterations == s_cpu.num_linear_iterations > 3
    assert s_dev.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-11)
    p_dev = make_ba_problem(**kw)
    p_cpu = copy_problem(p_dev)
    s_dev, s_cpu = emu_solve(emu, p_dev, **ITER), orc.ba_solve(p_cpu, linear_solver=1)
    assert (s_dev.num_successful_steps, s_dev.num_unsuccessful_steps, s_dev.termination_type) == \
           (s_cpu.num_successful_steps, s_cpu.num_unsuccessful_steps, s_cpu.termination)
    assert abs(s_dev.num_linear_solver_iterations - s_cpu.num_linear_iterations) <= 0.25 * s_cpu.num_linear_iterations
    assert s_dev.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-7)
    assert abs(reprojection_rms(p_dev) - reprojection_rms(p_cpu)) < 1e-6


def test_iterative_schur_inner_iteration_cap_and_robust_loss(emu):
    p_dev = make_ba_problem(n_img=12, n_pts=200, track_len=5, seed=21, noise_px=1.0)
    p_cpu = copy_problem(p_dev)
    kw = dict(max_num_iterations=8, gradient_tolerance=1e-12)
    s_dev = emu_solve(emu, p_dev, max_linear_solver_iterations=3, loss_function_type=1, **kw, **ITER)
    s_cpu = orc.ba_solve(p_cpu, linear_solver=1, max_linear_solver_iterations=3, loss_type=1, **kw)
    assert s_dev.num_linear_solver_iterations == s_cpu.num_linear_iterations <= 3 * 8
    assert s_dev.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-9)


def test_solver_choice_follows_the_reference_rule(emu):
    """bundle_adjustment.cc:274-284: direct solvers up to 1000 images, ITERATIVE_SCHUR above."""
    small = make_ba_problem(n_img=1000, n_pts=60, track_len=3, seed=2)
    large = make_ba_problem(n_img=1001, n_pts=60, track_len=3, seed=2)
    o = dict(max_num_iterations=1)
    assert emu_solve(emu, small, **o).linear_solver_type_used == 1
    s = emu_solve(emu, large, **o)
    assert s.linear_solver_type_used == 2 and s.num_linear_solver_iterations >= 1
    assert emu_solve(emu, make_ba_problem(n_img=6, n_pts=40, track_len=3, seed=2), linear_solver_type=2, **o).linear_solver_type_used == 2
    with pytest.raises(RuntimeError):
        emu_solve(emu, small, linear_solver_type=3)


def test_iterative_schur_allreduce_hook_two_shards(emu):
    """Two 'ranks' run one after the other is not possible (the hook is synchronous), so the hook is exercised with
    the identity on one rank: every collective call site of the CG loop is reached and the result is unchanged."""
    calls = []
    p_a = make_ba_problem(n_img=8, n_pts=100, track_len=4, seed=31)
    p_b = copy_problem(p_a)
    o = emu.BundleAdjustmentOptions()
    emu._L().b2_ba_default_options(C.byref(o))
    o.linear_solver_type = 2
    adj = emu.BundleAdjuster(o)
    adj.set_allreduce(lambda ptr, n, op: calls.append((n, op)))
    s_a = adj.Solve(p_a)
    adj.close()
    s_b = emu_solve(emu, p_b, **ITER)
    assert s_a.final_cost == s_b.final_cost and s_a.num_linear_solver_iterations == s_b.num_linear_solver_iterations
    D = s_a.num_effective_parameters_reduced - 3 * 100
    assert (D, 0) in calls and (4 * D, 0) in calls and (3 * D, 0) in calls and (1, 1) in calls


@pytest.mark.parametrize("solver", ["auto", "iterative"])
def test_bench_ba_leg_runs_on_the_emulated_library(emu, solver):
    """bench.py's BA leg (the code the driver runs on the GPU box) end to end on the emulated library: both solver
    settings produce a complete result dictionary with a finite roofline entry."""
    import sys
    import types
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        import bench
    finally:
        sys.argv = argv
    a = types.SimpleNamespace(ba="8,60,4", ba_solver=solver, no_cpu=True)
    out = bench.bench_ba(a, 0, 0, 1, 1, lambda: None, (6490.5, "test"))
    assert out["rms_px_final"] < out["rms_px_initial"] and out["termination"] == 0
    assert np.isfinite(out["roofline"]["achieved"]) and out["roofline"]["frac"] > 0
    assert ("ITERATIVE" in out["linear_solver"]) == (solver == "iterative")
    if solver == "iterative":
        assert out["cg_iterations"] > 0


def test_one_handle_serves_alternating_solver_types_and_problem_sizes(emu):
    """b2_ba_solve frees everything it allocated, so one handle can alternate between the exact and the iterative path and
    between problem sizes; every solve equals a solve on a fresh handle."""
    o = emu.BundleAdjustmentOptions()
    emu._L().b2_ba_default_options(C.byref(o))
    adj = emu.BundleAdjuster(o)
    try:
        for k, (solver, kw) in enumerate([(2, dict(n_img=7, n_pts=70, track_len=4, seed=1)), (1, dict(n_img=5, n_pts=40, track_len=3, seed=2)),
                                          (2, dict(n_img=9, n_pts=30, track_len=5, seed=3)), (1, dict(n_img=7, n_pts=70, track_len=4, seed=1))]):
            p_a = make_ba_problem(**kw)
            p_b = copy_problem(p_a)
            adj.options.linear_solver_type = solver
            s_a = adj.Solve(p_a)
            s_b = emu_solve(emu, p_b, linear_solver_type=solver)
            assert s_a.linear_solver_type_used == solver
            assert (s_a.final_cost, s_a.num_iterations, s_a.num_linear_solver_iterations) == \
                   (s_b.final_cost, s_b.num_iterations, s_b.num_linear_solver_iterations), k
            assert (p_a["xyz"] == p_b["xyz"]).all() and (p_a["qvec"] == p_b["qvec"]).all()
    finally:
        adj.close()


def test_iterative_schur_degenerate_inputs(emu):
    """No observations (BundleAdjuster::Solve returns false: nothing moves), every camera-side block constant (D = 0: only
    the points move), a single point, and a point observed twice by the same image (cross terms inside one block)."""
    p = make_ba_problem(n_img=4, n_pts=10, track_len=3, seed=1)
    for k in ("obs_img", "obs_pt"):
        p[k] = p[k][:0].copy()
    p["obs_xy"] = p["obs_xy"][:0].copy()
    q0 = p["qvec"].copy()
    s = emu_solve(emu, p, **ITER)
    assert s.num_iterations == 0 and (p["qvec"] == q0).all()
    # all poses and cameras constant
    p = make_ba_problem(n_img=5, n_pts=40, track_len=3, seed=3)
    p["pose_const"][:] = 1
    p["cam_const"][:] = 1
    pc = copy_problem(p)
    s, sc = emu_solve(emu, p, **ITER), orc.ba_solve(pc, linear_solver=1)
    assert s.num_effective_parameters_reduced == sc.num_effective_parameters == 3 * 40
    assert s.final_cost == pytest.approx(sc.final_cost, rel=1e-9) and s.final_cost < s.initial_cost
    assert s.num_linear_solver_iterations == 0
    # one point only
    p = make_ba_problem(n_img=4, n_pts=1, track_len=4, seed=5)
    pc = copy_problem(p)
    s, sc = emu_solve(emu, p, max_num_iterations=5, **ITER), orc.ba_solve(pc, max_num_iterations=5, linear_solver=1)
    assert s.final_cost == pytest.approx(sc.final_cost, rel=1e-6, abs=1e-12)
    # the same image twice in one track
    p = make_ba_problem(n_img=6, n_pts=50, track_len=4, seed=9)
    p["obs_img"][1] = p["obs_img"][0]
    p["obs_img"][5] = p["obs_img"][4]
    pc = copy_problem(p)
    s, sc = emu_solve(emu, p, max_num_iterations=4, **ITER), orc.ba_solve(pc, max_num_iterations=4, linear_solver=1)
    assert (s.num_successful_steps, s.num_unsuccessful_steps) == (sc.num_successful_steps, sc.num_unsuccessful_steps)
    assert s.num_linear_solver_iterations == sc.num_linear_iterations
    assert s.final_cost == pytest.approx(sc.final_cost, rel=1e-8)


# ---------------------------------------------------------------- general camera models (wide Jacobian layout)
GENERAL_CAMERAS = [
    (3, [1200.0, 500, 500, 0.05, -0.01]),                                              # RADIAL
    (4, [1180.0, 1210.0, 505.0, 495.0, -0.12, 0.03, 0.001, -0.0015]),                  # OPENCV
    (5, [1200.0, 1200.0, 500, 500, 0.02, -0.005, 0.001, 0.0]),                         # OPENCV_FISHEYE
    (7, [1200.0, 1200.0, 500, 500, 0.3]),                                              # FOV
    (8, [1200.0, 500, 500, 0.04]),                                                     # SIMPLE_RADIAL_FISHEYE
    (9, [1200.0, 500, 500, 0.04, -0.01]),                                              # RADIAL_FISHEYE
]


@pytest.mark.parametrize("camera", GENERAL_CAMERAS)
@pytest.mark.parametrize("solver", [1, 2])
def test_general_camera_models_follow_the_oracle(emu, camera, solver):
    """Problems with a camera model beyond SIMPLE_PINHOLE / PINHOLE / SIMPLE_RADIAL run on the wide Jacobian layout
    (ObsJacW: 12 intrinsics slots, dual-number derivatives of WorldToImage) through both linear solvers: same LM path
    and cost as the oracle with the final-BA options, same optimum when run to convergence."""
    kw = dict(n_img=8, n_pts=120, track_len=5, seed=3, camera=camera, noise_px=0.5)
    p_dev = make_ba_problem(**kw)
    p_cpu = copy_problem(p_dev)
    s_dev = emu_solve(emu, p_dev, linear_solver_type=solver, max_num_iterations=6)
    s_cpu = orc.ba_solve(p_cpu, linear_solver=solver - 1, max_num_iterations=6)
    assert (s_dev.num_residuals_reduced, s_dev.num_effective_parameters_reduced) == (s_cpu.num_residuals, s_cpu.num_effective_parameters)
    assert s_dev.initial_cost == pytest.approx(s_cpu.initial_cost, rel=1e-12)
    assert (s_dev.num_successful_steps, s_dev.num_unsuccessful_steps) == (s_cpu.num_successful_steps, s_cpu.num_unsuccessful_steps)
    assert s_dev.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-7)
    if solver == 2:
        assert abs(s_dev.num_linear_solver_iterations - s_cpu.num_linear_iterations) <= 2
    n = len(camera[1])
    if solver == 1:   # (the truncated inner solves of ITERATIVE_SCHUR leave the unobservable high-order coefficients free)
        assert np.allclose(p_dev["cam_params"][:, :n], p_cpu["cam_params"][:, :n], rtol=1e-6, atol=1e-8)
    assert (p_dev["cam_params"][:, n:] == 0).all()                      # unused slots stay untouched
    if solver == 1:   # (the inexact steps reach the same optimum, shown for the 4-slot layout above and by the oracle tests)
        p_dev = make_ba_problem(**kw)
        start = reprojection_rms(p_dev)
        emu_solve(emu, p_dev, linear_solver_type=solver, **TIGHT)
        assert reprojection_rms(p_dev) < 0.36 < 4 < start


def test_mixed_camera_models_and_shared_general_camera(emu):
    """One problem with SIMPLE_RADIAL and OPENCV cameras side by side (everything moves to the wide layout), and a single
    OPENCV camera shared by all images (an 8 x 8 preconditioner block fed by every observation)."""
    cam = (4, [1180.0, 1210.0, 505.0, 495.0, -0.12, 0.03, 0.001, -0.0015])
    p = make_ba_problem(n_img=8, n_pts=100, track_len=4, seed=5, camera=cam, noise_px=0.5)
    # images 0..3 keep the OPENCV camera; 4..7 get a SIMPLE_RADIAL with the same (undistorted-ish) parameters and their
    # observations re-projected accordingly
    for c in range(4, 8):
        p["cam_model"][c] = 2
        p["cam_params"][c] = 0
        p["cam_params"][c, :4] = [1195.0, 505.0, 495.0, 0.0]
    for solver in (1, 2):
        a, b = copy_problem(p), copy_problem(p)
        s_dev = emu_solve(emu, a, linear_solver_type=solver, max_num_iterations=5)
        s_cpu = orc.ba_solve(b, linear_solver=solver - 1, max_num_iterations=5)
        assert s_dev.num_effective_parameters_reduced == s_cpu.num_effective_parameters == 3 * 100 + (6 * 8 - 7) + 4 * 6 + 4 * 2
        assert (s_dev.num_successful_steps, s_dev.num_unsuccessful_steps) == (s_cpu.num_successful_steps, s_cpu.num_unsuccessful_steps)
        assert s_dev.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-7)
    q = make_ba_problem(n_img=8, n_pts=100, track_len=4, seed=6, camera=cam, noise_px=0.5, shared_camera=True)
    for solver in (1, 2):
        a, b = copy_problem(q), copy_problem(q)
        s_dev = emu_solve(emu, a, linear_solver_type=solver, max_num_iterations=5)
        s_cpu = orc.ba_solve(b, linear_solver=solver - 1, max_num_iterations=5)
        assert s_dev.num_effective_parameters_reduced == s_cpu.num_effective_parameters == 3 * 100 + (6 * 8 - 7) + 6
        assert s_dev.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-7)


def test_camera_params_stride_is_checked(emu):
    p = make_ba_problem(n_img=4, n_pts=20, track_len=3, seed=1)
    p["cam_model"][:] = 4                      # OPENCV needs 8 parameters, the array has 4 per camera
    with pytest.raises(RuntimeError):
        emu_solve(emu, p)
    p["cam_model"][:] = 11
    with pytest.raises(RuntimeError):
        emu_solve(emu, p)


def test_mean_reprojection_error_metric(emu):
    """b2_ba_reprojection_errors (Reconstruction::ComputeMeanReprojectionError + Point3D errors) on the
    emulated kernel against the oracle's restatement and an independent numpy evaluation; points behind a camera are
    skipped but still count in the denominators, as in the reference."""
    from tests.ba_scene import mean_reprojection_error
    L = emu._L()
    L.b2_ba_reprojection_errors.argtypes = [C.c_void_p, C.POINTER(emu.BaProblem), C.c_void_p, C.POINTER(C.c_double)]
    adj = emu.BundleAdjuster(emu.BundleAdjustmentOptions())
    try:
        for cam in (None, (4, [1180.0, 1210.0, 505.0, 495.0, -0.12, 0.03, 0.001, -0.0015]), (8, [1200.0, 500, 500, 0.04])):
            p = make_ba_problem(n_img=9, n_pts=150, track_len=5, seed=12, camera=cam)
            p["qvec"][3] *= 1.7                        # the metric normalises quaternions itself
            mean, err = adj.ComputeMeanReprojectionError(p)
            omean, oerr = orc.ba_mean_reprojection_error(p)
            assert mean == pytest.approx(omean, rel=1e-13) and np.allclose(err, oerr, rtol=1e-13, atol=1e-15)
            p["qvec"][3] /= 1.7
            assert mean == pytest.approx(mean_reprojection_error(p), rel=1e-10)
            assert err.shape == (150,) and (err > 0).all()
        p = make_ba_problem(n_img=6, n_pts=40, track_len=4, seed=2)
        p["xyz"][5] = [0, 0, -50.0]                    # far behind every camera on the ring? not all: use the oracle as reference
        mean, err = adj.ComputeMeanReprojectionError(p)
        omean, oerr = orc.ba_mean_reprojection_error(p)
        assert mean == pytest.approx(omean, rel=1e-13) and np.allclose(err, oerr, rtol=1e-13, atol=1e-15)
        p["obs_img"] = p["obs_img"][:0].copy(); p["obs_pt"] = p["obs_pt"][:0].copy(); p["obs_xy"] = p["obs_xy"][:0].copy()
        assert adj.ComputeMeanReprojectionError(p)[0] == 0.0
    finally:
        adj.close()


@pytest.mark.parametrize("kw", [
    dict(n_img=6, n_pts=60, track_len=4, seed=5),
    dict(n_img=8, n_pts=80, track_len=5, seed=3, shared_camera=True),      # intrinsics shared: atomics across image blocks
    dict(n_img=150, n_pts=40, track_len=3, seed=8),                        # images without observations
    dict(n_img=8, n_pts=90, track_len=4, seed=7, n_const_pts=20),
    dict(n_img=8, n_pts=80, track_len=4, seed=3, camera=(4, [1180.0, 1210.0, 505.0, 495.0, -0.12, 0.03, 0.001, -0.0015]), noise_px=0.5),
])
def test_image_major_camera_terms_variant(emu, kw, monkeypatch):
    """B2_BA_CAMTERMS=image (exact path): U = Jc'Jc, g_c and diag_c summed per image in registers and written once per
    image instead of ~75 atomics per observation.  Same LM path as the production kernel and the oracle."""
    p_v = make_ba_problem(**kw)
    p_ref, p_cpu = copy_problem(p_v), copy_problem(p_v)
    s_ref = emu_solve(emu, p_ref, linear_solver_type=1)
    monkeypatch.setenv("B2_BA_CAMTERMS", "image")
    s_v = emu_solve(emu, p_v, linear_solver_type=1)
    monkeypatch.delenv("B2_BA_CAMTERMS")
    s_cpu = orc.ba_solve(p_cpu)
    for s in (s_ref, s_cpu):
        assert (s_v.num_successful_steps, s_v.num_unsuccessful_steps) == (s.num_successful_steps, s.num_unsuccessful_steps)
        assert s_v.final_cost == pytest.approx(s.final_cost, rel=1e-9)
    assert np.abs(p_v["xyz"] - p_ref["xyz"]).max() < 1e-8 and np.abs(p_v["qvec"] - p_ref["qvec"]).max() < 1e-10


# ---------------------------------------------------------------- fused exact path (ba_fused.cu + ba_chol.cu)
def spd_cases():
    rng = np.random.default_rng(5)
    out = []
    for D, band in ((37, None), (150, None), (300, 70), (450, 64)):
        M = rng.normal(size=(D, D))
        A = M @ M.T + D * np.eye(D)
        if band is not None:                         # banded camera graph: zero 64 x 64 tiles away from the diagonal
            i, j = np.indices((D, D))
            A[np.abs(i - j) > band] = 0.0
            A += np.eye(D) * np.abs(A).sum(1).max()
        out.append((A, rng.normal(size=D), band))
    return out


def test_tiled_cholesky_seam_solves_dense_and_banded_systems(emu):
    adj = emu.BundleAdjuster(emu.BundleAdjustmentOptions())
    try:
        for A, b, band in spd_cases():
            x, info, n_tiles = adj.debug_cholesky_solve(A, b)
            nt = (len(b) + 63) // 64
            assert info == 0
            assert n_tiles == nt * (nt + 1) // 2 if band is None else n_tiles < nt * (nt + 1) // 2
            ref = np.linalg.solve(A, b)
            assert np.abs(x - ref).max() <= 1e-12 * max(1.0, np.abs(ref).max()) * np.linalg.cond(A)
        A = np.eye(70)
        A[3, 3] = -1.0                               # not positive definite: reported, not hidden
        assert adj.debug_cholesky_solve(A, np.ones(70))[1] != 0
    finally:
        adj.close()


@pytest.mark.parametrize("kw", [
    dict(n_img=6, n_pts=60, track_len=4, seed=5),
    dict(n_img=8, n_pts=80, track_len=5, seed=3, shared_camera=True),
    dict(n_img=30, n_pts=200, track_len=12, seed=8),                    # 12-slot windows, several tiles
    dict(n_img=30, n_pts=120, track_len=16, seed=2),                    # 16-slot windows
    dict(n_img=8, n_pts=90, track_len=4, seed=7, n_const_pts=20),
])
def test_fused_and_staged_exact_paths_agree(emu, kw, monkeypatch):
    """The default exact step is the fused one (exact_path_used == 2); B2_BA_EXACT=staged selects the first-generation
    kernels.  Same LM path (reference options) and the same optimum as the oracle on both."""
    p_f, p_s, p_cpu = make_ba_problem(**kw), make_ba_problem(**kw), make_ba_problem(**kw)
    s_f = emu_solve(emu, p_f)
    monkeypatch.setenv("B2_BA_EXACT", "staged")
    s_s = emu_solve(emu, p_s)
    monkeypatch.delenv("B2_BA_EXACT")
    s_cpu = orc.ba_solve(p_cpu)
    assert (s_f.exact_path_used, s_s.exact_path_used) == (2, 1)
    assert s_f.linear_solve_seconds >= 0 and s_f.reduced_system_bytes > 0
    for s in (s_f, s_s):
        assert (s.num_successful_steps, s.num_unsuccessful_steps, s.termination_type) == \
               (s_cpu.num_successful_steps, s_cpu.num_unsuccessful_steps, s_cpu.termination)
        assert s.final_cost == pytest.approx(s_cpu.final_cost, rel=1e-9)
    assert abs(reprojection_rms(p_f) - reprojection_rms(p_cpu)) < 1e-6
    assert np.abs(p_f["xyz"] - p_s["xyz"]).max() < 1e-7


def test_tracks_that_do_not_fit_a_window_take_the_staged_path(emu):
    p = make_ba_problem(n_img=40, n_pts=30, track_len=36, seed=6)
    s = emu_solve(emu, p, **TIGHT)
    assert s.exact_path_used == 1


def rotation_only_problem():
    p = make_ba_problem(n_img=6, n_pts=80, track_len=4, seed=12, noise_px=3.0)
    p["tvec_const"][:] = 7
    p["cam_const"][:] = 1
    p["pt_const"][:] = 1
    p["pose_const"][:] = 0
    return p


def test_gradient_max_norm_goes_through_the_quaternion_plus(emu):
    """Ceres' gradient_max_norm is |x - Plus(x, -g)|_inf (trust_region_minimizer.cc): for a rotation block the
    quaternion's displacement under QuaternionParameterization::Plus, bounded by 2 whatever the gradient.  With only
    rotations free and gradient_tolerance = 2 the solve therefore converges at iteration 0 although |g|_inf is in the
    thousands; with the tvec free as well, the plain |g| of its columns keeps the solve going."""
    for solver in (1, 2):
        p, q = rotation_only_problem(), rotation_only_problem()
        s = emu_solve(emu, p, gradient_tolerance=2.0, linear_solver_type=solver)
        sc = orc.ba_solve(q, gradient_tolerance=2.0, linear_solver=0 if solver == 1 else 1)
        assert (s.num_iterations, s.termination_type) == (0, 0)
        assert (sc.num_successful_steps + sc.num_unsuccessful_steps, sc.termination) == (0, 0)
        p, q = rotation_only_problem(), rotation_only_problem()
        p["tvec_const"][:] = 0
        q["tvec_const"][:] = 0
        s = emu_solve(emu, p, gradient_tolerance=2.0, linear_solver_type=solver)
        sc = orc.ba_solve(q, gradient_tolerance=2.0, linear_solver=0 if solver == 1 else 1)
        assert s.num_iterations > 0 and sc.num_successful_steps > 0
        assert (s.num_successful_steps, s.num_unsuccessful_steps) == (sc.num_successful_steps, sc.num_unsuccessful_steps)
