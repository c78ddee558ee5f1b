"""Pins of the BA oracle: the residual goldens of the reference (cost_functions_test.cc:41-98),
the Jacobian against central differences, the structural counts the reference's BA tests
assert (bundle_adjustment_test.cc:186-233), and convergence on a synthetic scene."""
import numpy as np

from oracle import pyoracle as orc
from tests.ba_scene import copy_problem, make_ba_problem, reprojection_rms


def test_cost_function_goldens():
    # cost_functions_test.cc:41-67 (SIMPLE_PINHOLE, identity pose): params {f, cx, cy}
    q, t = [1, 0, 0, 0], [0, 0, 0]
    r = orc.ba_evaluate(0, q, t, [0, 0, 1], [1, 0, 0], [0, 0])[0]
    assert r.tolist() == [0, 0]
    r = orc.ba_evaluate(0, q, t, [0, 0, 2], [1, 0, 0], [0, 0])[0]   # "1,1,2"-style checks:
    assert r.tolist() == [0, 0]
    r = orc.ba_evaluate(0, q, t, [-1, 1, 1], [2, 0, 0], [0, 0])[0]  # X=(-1,1,1), f=2 -> (-2, 2)
    assert r.tolist() == [-2, 2]
    r = orc.ba_evaluate(0, q, [0, 0, 1], [1, 1, 1], [1, 0, 0], [0, 0])[0]
    assert np.allclose(r, [0.5, 0.5])


def test_jacobian_vs_central_differences():
    rng = np.random.default_rng(0)
    for model, k in ((2, [1200.0, 500, 510, 0.03]), (0, [900.0, 400, 300, 0]), (1, [900.0, 950, 400, 300])):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        t = rng.normal(size=3) + [0, 0, 6]
        X = rng.normal(size=3)
        obs = rng.uniform(0, 1000, 2)
        r, Jq, Jt, JX, Jk = orc.ba_evaluate(model, q, t, X, k, obs)
        h = 1e-6
        for c in range(3):
            d = np.zeros(3); d[c] = h
            rp = orc.ba_evaluate(model, orc.ba_quat_plus(q, d), t, X, k, obs)[0]
            rm = orc.ba_evaluate(model, orc.ba_quat_plus(q, -d), t, X, k, obs)[0]
            assert np.allclose((rp - rm) / (2 * h), Jq[:, c], rtol=1e-5, atol=1e-4)
            rp = orc.ba_evaluate(model, q, t + d, X, k, obs)[0]; rm = orc.ba_evaluate(model, q, t - d, X, k, obs)[0]
            assert np.allclose((rp - rm) / (2 * h), Jt[:, c], rtol=1e-5, atol=1e-4)
            rp = orc.ba_evaluate(model, q, t, X + d, k, obs)[0]; rm = orc.ba_evaluate(model, q, t, X - d, k, obs)[0]
            assert np.allclose((rp - rm) / (2 * h), JX[:, c], rtol=1e-5, atol=1e-4)
        for c in range(4 if model else 3):
            d = np.zeros(4); d[c] = h * max(1.0, abs(k[c]))
            rp = orc.ba_evaluate(model, q, t, X, np.array(k) + d, obs)[0]
            rm = orc.ba_evaluate(model, q, t, X, np.array(k) - d, obs)[0]
            assert np.allclose((rp - rm) / (2 * d[c]), Jk[:, c], rtol=1e-5, atol=1e-5)


def test_structure_counts_two_view():
    # bundle_adjustment_test.cc:186-233 TestTwoView: 100 points seen by 2 images ->
    # 400 residuals, 309 effective parameters (300 + 5 [pose of image 1 minus tvec.x] + 2 * 2 intrinsics)
    prob = make_ba_problem(n_img=2, n_pts=100, track_len=2, seed=1)
    s = orc.ba_solve(prob, max_num_iterations=2)
    assert s.num_residuals == 400
    assert s.num_effective_parameters == 309


def test_constant_blocks_stay_bit_identical():
    # CheckConstant* macros of bundle_adjustment_test.cc:41-107
    prob = make_ba_problem(n_img=6, n_pts=120, track_len=4, seed=2, n_const_pts=15)
    prob["cam_const"][2] = 1
    before = copy_problem(prob)
    orc.ba_solve(prob, max_num_iterations=10)
    assert (prob["qvec"][0] == before["qvec"][0] / np.linalg.norm(before["qvec"][0])).all()
    assert (prob["tvec"][0] == before["tvec"][0]).all()
    assert prob["tvec"][1][0] == before["tvec"][1][0] and (prob["tvec"][1][1:] != before["tvec"][1][1:]).all()
    c = before["pt_const"].astype(bool)
    assert (prob["xyz"][c] == before["xyz"][c]).all() and (prob["xyz"][~c] != before["xyz"][~c]).any()
    assert (prob["cam_params"][2] == before["cam_params"][2]).all()
    assert (prob["cam_params"][:, 1:3] == before["cam_params"][:, 1:3]).all()      # cx, cy never refined
    assert (prob["cam_params"][[0, 1, 3], 0] != before["cam_params"][[0, 1, 3], 0]).all()


def test_converges_to_noise_floor():
    prob = make_ba_problem(n_img=24, n_pts=600, track_len=6, seed=3)
    rms0 = reprojection_rms(prob)
    s = orc.ba_solve(prob, max_num_iterations=50)
    rms1 = reprojection_rms(prob)
    # U(-2,2) noise has sigma 1.155 px per coordinate -> RMS ~ 1.63 px minus what the fit absorbs
    assert rms0 > 5 and 1.2 < rms1 < 1.7
    assert abs(np.sqrt(2 * s.final_cost / len(prob["obs_img"])) - rms1) < 1e-9
    assert s.num_successful_steps >= 3
    # shared intrinsics: one camera for all images
    prob = make_ba_problem(n_img=12, n_pts=300, track_len=5, seed=4, shared_camera=True)
    s = orc.ba_solve(prob, max_num_iterations=50)
    assert 1.2 < reprojection_rms(prob) < 1.7
    assert s.num_effective_parameters == 300 * 3 + 11 * 6 - 1 + 2


def test_iterative_schur_restatement_reaches_the_exact_schur_optimum():
    """ITERATIVE_SCHUR + SCHUR_JACOBI (what BundleAdjuster::Solve selects above 1000 images,
    bundle_adjustment.cc:274-284): the matrix-free Schur product, the block-Jacobi preconditioner and Ceres'
    CG loop are restated independently of the dense exact-step branch; both must stop at the same optimum, the
    inexact steps must never increase the cost, and the inner-iteration cap must hold."""
    T = dict(max_num_iterations=200, gradient_tolerance=1e-9, function_tolerance=1e-16)
    for kw in (dict(n_img=12, n_pts=300, track_len=5, seed=4), dict(n_img=8, n_pts=80, track_len=5, seed=3, shared_camera=True),
               dict(n_img=9, n_pts=120, track_len=4, seed=8, n_const_pts=30)):
        p_ex = make_ba_problem(**kw)
        p_it = copy_problem(p_ex)
        s_ex, s_it = orc.ba_solve(p_ex, **T), orc.ba_solve(p_it, linear_solver=1, **T)
        assert s_it.num_linear_iterations > 0
        assert s_it.final_cost == __import__("pytest").approx(s_ex.final_cost, rel=1e-10)
        assert abs(reprojection_rms(p_ex) - reprojection_rms(p_it)) < 1e-9
        assert np.abs(p_ex["xyz"] - p_it["xyz"]).max() < 1e-6
    costs = []
    for n in range(1, 6):   # cost after n LM iterations is non-increasing; at most 4 inner iterations per step
        p = make_ba_problem(n_img=12, n_pts=300, track_len=5, seed=4)
        s = orc.ba_solve(p, linear_solver=1, max_num_iterations=n, max_linear_solver_iterations=4)
        assert s.num_linear_iterations <= 4 * n
        costs.append(s.final_cost)
    assert all(b <= a for a, b in zip(costs, costs[1:])) and costs[-1] < costs[0]


def test_jacobian_of_every_camera_model_vs_central_differences():
    """BundleAdjustmentCostFunction on all eleven models: the three hand-derived ones and the eight evaluated on dual
    numbers (what Ceres' autodiff does with the reference's templated WorldToImage)."""
    from tests.camera_cases import CAMERA_CASES, NUM_PARAMS
    rng = np.random.default_rng(0)
    for model, params in CAMERA_CASES:
        K = NUM_PARAMS[model]
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        t = rng.normal(size=3) + [0, 0, 6]
        X = rng.normal(size=3)
        obs = rng.uniform(0, 800, 2)
        k = np.array(params, float)
        r, Jq, Jt, JX, Jk = orc.ba_evaluate(model, q, t, X, k, obs)

        def res(q_, t_, X_, k_):
            return orc.ba_evaluate(model, q_, t_, X_, k_, obs)[0]
        h = 1e-6
        for c in range(3):
            e = np.zeros(3); e[c] = h
            assert np.allclose(Jt[:, c], (res(q, t + e, X, k) - res(q, t - e, X, k)) / (2 * h), rtol=1e-5, atol=1e-5)
            assert np.allclose(JX[:, c], (res(q, t, X + e, k) - res(q, t, X - e, k)) / (2 * h), rtol=1e-5, atol=1e-5)
            qp, qm = orc.ba_quat_plus(q, e), orc.ba_quat_plus(q, -e)
            assert np.allclose(Jq[:, c], (res(qp, t, X, k) - res(qm, t, X, k)) / (2 * h), rtol=1e-5, atol=1e-4)
        for c in range(K):
            if model == 7 and abs(k[4] ** 2 - 1e-4) < 1e-9:
                continue                                   # omega^2 == kEpsilon: the difference would straddle two branches
            hp = 1e-6 * max(1.0, abs(k[c]))
            e = np.zeros(K); e[c] = hp
            num = (res(q, t, X, k + e) - res(q, t, X, k - e)) / (2 * hp)
            assert np.allclose(Jk[:, c], num, rtol=2e-5, atol=1e-5 * max(1.0, np.abs(num).max())), (model, c)


def test_general_camera_models_exact_and_iterative_reach_the_same_optimum():
    T = dict(max_num_iterations=100, gradient_tolerance=1e-9, function_tolerance=1e-14)
    for cam in ((3, [1200.0, 500, 500, 0.05, -0.01]), (4, [1180.0, 1210.0, 505.0, 495.0, -0.12, 0.03, 0.001, -0.0015]),
                (5, [1200.0, 1200.0, 500, 500, 0.02, -0.005, 0.001, 0.0]), (7, [1200.0, 1200.0, 500, 500, 0.3]),
                (8, [1200.0, 500, 500, 0.04])):
        p = make_ba_problem(n_img=10, n_pts=200, track_len=5, seed=3, camera=cam, noise_px=0.5)
        a, b = copy_problem(p), copy_problem(p)
        s_ex, s_it = orc.ba_solve(a, **T), orc.ba_solve(b, linear_solver=1, **T)
        n_extra = len(cam[1]) - (4 if cam[0] in (4, 5, 7) else 3)
        n_focal = 2 if cam[0] in (4, 5, 7) else 1
        assert s_ex.num_effective_parameters == 3 * 200 + (6 * 10 - 7) + 10 * (n_focal + n_extra)   # focal + extra refined, pp fixed
        assert reprojection_rms(a) < 0.36 < 5 < reprojection_rms(p)
        assert abs(reprojection_rms(a) - reprojection_rms(b)) < 1e-6
        assert s_it.final_cost == __import__("pytest").approx(s_ex.final_cost, rel=1e-8)
        assert (a["cam_params"][:, len(cam[1]):] == 0).all()


def test_reprojection_error_reference_cases():
    """base/projection_test.cc:95-124 (CalculateSquaredReprojectionError: 0 at the exact projection, 2 one pixel off in
    x and y) and base/pose_test.cc:154-172 (QuaternionRotatePoint normalises its quaternion) through the B8 metric."""
    rng = np.random.default_rng(0)
    X = np.abs(rng.uniform(-1, 1, 3))
    base = dict(qvec=np.array([[1.0, 0, 0, 0]]), tvec=np.zeros((1, 3)), img_cam=np.zeros(1, np.int32), pose_const=np.zeros(1, np.uint8),
                tvec_const=np.zeros(1, np.uint8), cam_model=np.zeros(1, np.int32), cam_params=np.array([[1.0, 0, 0, 0]]),
                cam_const=np.zeros(1, np.uint8), xyz=X[None].copy(), pt_const=np.zeros(1, np.uint8), obs_img=np.zeros(1, np.int32),
                obs_pt=np.zeros(1, np.int32), obs_xy=(X[:2] / X[2])[None].copy())
    mean, err = orc.ba_mean_reprojection_error(base)
    assert mean == 0 and err[0] == 0
    off = {k: v.copy() for k, v in base.items()}
    off["obs_xy"] += 1
    assert orc.ba_mean_reprojection_error(off)[0] ** 2 == __import__("pytest").approx(2, rel=1e-8)
    scaled = {k: v.copy() for k, v in base.items()}
    scaled["qvec"] *= 0.1                                  # Vector4d(0.1, 0, 0, 0) rotates like the identity
    assert orc.ba_mean_reprojection_error(scaled)[0] == 0
    behind = {k: v.copy() for k, v in off.items()}
    behind["xyz"][0, 2] = -1.0                             # not in front of the camera: skipped, the track still counts
    assert orc.ba_mean_reprojection_error(behind) [0] == 0
    flip = {k: v.copy() for k, v in base.items()}          # rotation by pi about x maps (1, 1, z) to (1, -1, -z)
    flip["qvec"][0] = [0, 1, 0, 0]
    flip["xyz"][0] = [1, 1, -2]
    flip["obs_xy"][0] = [0.5, -0.5]
    assert orc.ba_mean_reprojection_error(flip)[0] < 1e-15
