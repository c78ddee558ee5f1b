"""The two-view verification CUDA sources (verify_kernel.cu, verify_solvers.cuh, verify_api.cu) compiled for the
HOST against tests/cuda_emu/cuda_emu.h (warps as groups of fibers, shuffles / ballots / __syncwarp emulated)
and driven through the C ABI: the same checks as tests/test_verify_gpu.py at small sizes, without a GPU.
TEST of the CUDA code -- the product library is not involved."""
import ctypes as C

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.tv_scene import scene


@pytest.fixture(scope="module")
def ver():
    from tests.cuda_emu.build_emu import BA_SOURCES, VERIFY_SOURCES, build
    import dagsfm_b200.verification as vm
    L = C.CDLL(str(build("verify", VERIFY_SOURCES)))
    L.b2_last_error.restype = C.c_char_p
    saved = (vm._L, vm.check, vm._bound)
    vm._bound = False
    real_lib = vm.lib

    def check(rc):
        if rc != 0:
            raise RuntimeError(f"emulated library error {rc}: {L.b2_last_error().decode()}")
    vm.lib = lambda: L            # _L() binds argtypes on whatever lib() returns
    vm.check = check
    v = vm.TwoViewGeometryVerifier(0)
    yield v
    v.close()
    vm.lib, vm.check, vm._bound = real_lib, saved[1], False


@pytest.mark.parametrize("total,k", [(50, 7), (7, 7), (200, 5), (33, 4), (20, 1)])
def test_sample_stream_bit_exact(ver, total, k):
    for seed in (0, 12345):
        assert (ver.debug_sample_stream(seed, total, k, 40) == orc.sample_stream(seed, total, k, 40)).all()


def _match(got, exp, tol):
    assert len(got) == len(exp)
    for g, e in zip(got, exp):
        assert min(np.abs(g - e).max(), np.abs(g + e).max()) < tol * max(1.0, np.abs(e).max())


def test_minimal_and_local_solvers(ver):
    rng = np.random.default_rng(0)
    for it in range(4):
        p1, p2 = scene(rng, 40, 0, noise=0.5)
        _match(ver.debug_solve(1, p1[:7], p2[:7]), orc.f7(p1[:7], p2[:7]), 1e-7)
        _match(ver.debug_solve(2, p1[:4], p2[:4]), [orc.h_dlt(p1[:4], p2[:4])], 1e-7)
        n1, n2 = (p1 - 500) / 1200, (p2 - 500) / 1200
        _match(ver.debug_solve(0, n1[:5], n2[:5]), orc.e5(n1[:5], n2[:5]), 1e-6)
        _match(ver.debug_solve(3, p1, p2), [orc.eight_point(p1, p2)], 1e-7)          # warp QR + Jacobi on R
        _match(ver.debug_solve(2, p1, p2), [orc.h_dlt(p1, p2)], 1e-7)
        _match(ver.debug_solve(0, n1, n2), orc.e5(n1, n2), 1e-5)


def test_score_models_bit_exact(ver):
    rng = np.random.default_rng(2)
    p1, p2 = scene(rng, 70, 30, noise=1.0)
    F = orc.eight_point(p1[:70], p2[:70])
    for typ, models, thr in ((1, np.stack([F, F * 3.0, rng.normal(size=(3, 3))]), 16.0),
                             (2, np.stack([orc.h_dlt(p1[:30], p2[:30]), np.eye(3)]), 16.0)):
        counts, sums, masks = ver.score_models(typ, p1, p2, models, thr)
        for k, M in enumerate(models):
            r = orc.residuals(typ, p1, p2, M)
            m = r <= thr
            s = 0.0
            for v in r[m]:
                s += v
            assert counts[k] == m.sum() and (masks[k] == m).all() and sums[k] == s


def test_two_view_decisions_equal_oracle(ver):
    from dagsfm_b200 import Camera, TwoViewOptions
    rng = np.random.default_rng(9)
    specs = [(60, 12, False, True), (50, 10, True, False), (48, 20, False, False), (9, 0, False, True)]
    cams, kps, pairs, offs, ms, priors = [], [], [], [0], [], []
    for i, (n_in, n_out, planar, prior) in enumerate(specs):
        p1, p2 = scene(rng, n_in, n_out, planar=planar, noise=0.3)
        perm = np.random.default_rng(100 + i).permutation(len(p2))
        kps += [p1, p2[perm]]
        ms.append(np.stack([np.arange(len(p1)), np.argsort(perm)], 1))
        cams += [Camera.make(prior_focal=prior), Camera.make(prior_focal=prior)]
        priors.append(prior)
        pairs.append((2 * i, 2 * i + 1))
        offs.append(offs[-1] + len(p1))
    ver.set_images(cams, kps)
    seeds = np.arange(len(specs), dtype=np.uint32) + 77
    opt = TwoViewOptions.default()
    opt.max_num_trials = 300          # keep the emulated H-RANSAC of the non-planar pairs short
    oopt = orc.tv_default_options()
    oopt.max_num_trials = 300
    res, inl = ver.verify_pairs(pairs, offs, np.concatenate(ms), opt, seeds)
    cfgs = []
    for i in range(len(specs)):
        c = orc.make_camera(prior=priors[i])
        # the oracle evaluated in the kernel's own operation order (its "device-order" solver stack): one floating-point
        # stack on both sides, so the model matrices must agree bit for bit, not only the decisions
        with orc.solver_stack(1):
            r1, oi1 = orc.two_view(c, kps[2 * i], c, kps[2 * i + 1], ms[i], oopt, seed=int(seeds[i]))
        g = res[i]
        for m in ("E", "F", "H"):
            assert np.array_equal(np.array(getattr(r1, m)[:]).view(np.uint64), g[m].view(np.uint64)), (i, m)
        assert (g["config"], g["n_inliers"], g["E_num_trials"], g["F_num_trials"], g["H_num_trials"]) == \
               (r1.config, r1.n_inliers, r1.E_trials, r1.F_trials, r1.H_trials)
        assert inl[offs[i]:offs[i] + r1.n_inliers].tolist() == oi1.tolist()
        # ... and in its independent stack (sequential sums): the same decisions
        r, oi = orc.two_view(c, kps[2 * i], c, kps[2 * i + 1], ms[i], oopt, seed=int(seeds[i]))
        assert (g["config"], g["n_inliers"], g["E_num_inliers"], g["F_num_inliers"], g["H_num_inliers"]) == \
               (r.config, r.n_inliers, r.E_inl, r.F_inl, r.H_inl), i
        assert (g["E_num_trials"], g["F_num_trials"], g["H_num_trials"]) == (r.E_trials, r.F_trials, r.H_trials)
        assert inl[offs[i]:offs[i] + r.n_inliers].tolist() == oi.tolist()
        cfgs.append(int(g["config"]))
    assert cfgs[3] == 1 and 6 in cfgs and (2 in cfgs or 3 in cfgs)


def test_estimate_multiple_through_the_c_abi(ver):
    """b2_verify_pairs_multiple (rounds of the batched Estimate, verify_multiple.h) on the emulated kernel."""
    from dagsfm_b200 import Camera, TwoViewOptions
    from tests.test_host_multiple import two_motion_pair
    rng = np.random.default_rng(3)
    kps, pairs, offs, ms = [], [], [0], []
    for k, (na, nb, no) in enumerate([(60, 50, 10), (70, 0, 20), (10, 0, 0)]):
        if nb:
            p1, p2, m = two_motion_pair(rng, na, nb, no)
        else:
            p1, p2 = scene(rng, na, no, noise=0.3)
            m = np.stack([np.arange(len(p1))] * 2, 1).astype(np.uint32)
        kps += [p1, p2]; pairs.append((2 * k, 2 * k + 1)); ms.append(m); offs.append(offs[-1] + len(m))
    ver.set_images([Camera.make(prior_focal=False)] * len(kps), kps)
    seeds = np.arange(3, dtype=np.uint32) + 40
    opt = TwoViewOptions.default()
    opt.max_num_trials = 300
    oopt = orc.tv_default_options()
    oopt.max_num_trials = 300
    res, inl = ver.verify_pairs_multiple(pairs, offs, np.concatenate(ms), opt, seeds)
    cam = orc.make_camera(prior=False)
    cfgs = []
    for k, (i, j) in enumerate(pairs):
        cfg, geos, exp_inl = orc.two_view_multiple(cam, kps[i], cam, kps[j], ms[k], oopt, seed=int(seeds[k]))
        cfgs.append(cfg)
        assert res["config"][k] == cfg and res["n_inliers"][k] == len(exp_inl)
        assert inl[offs[k]:offs[k] + len(exp_inl)].tolist() == exp_inl.tolist()
    assert cfgs[0] == 8 and cfgs[2] == 1


def test_stage_launch_shapes_keep_every_decision(ver, monkeypatch):
    """B2_VERIFY_BPS=222 (two CTAs per SM for the E, F and H stage kernels: the instances that score two hypotheses per pass
    instead of four): decisions, inlier lists and trial counts must stay those of the production shapes and the oracle."""
    from dagsfm_b200 import Camera, TwoViewOptions
    rng = np.random.default_rng(19)
    specs = [(70, 15, False, True), (55, 12, True, False), (50, 25, False, False), (64, 0, False, True)]
    cams, kps, pairs, offs, ms, priors = [], [], [], [0], [], []
    for i, (n_in, n_out, planar, prior) in enumerate(specs):
        p1, p2 = scene(rng, n_in, n_out, planar=planar, noise=0.3)
        perm = np.random.default_rng(200 + i).permutation(len(p2))
        kps += [p1, p2[perm]]
        ms.append(np.stack([np.arange(len(p1)), np.argsort(perm)], 1))
        cams += [Camera.make(prior_focal=prior), Camera.make(prior_focal=prior)]
        priors.append(prior); pairs.append((2 * i, 2 * i + 1)); offs.append(offs[-1] + len(p1))
    ver.set_images(cams, kps)
    seeds = np.arange(len(specs), dtype=np.uint32) + 5
    opt = TwoViewOptions.default(); opt.max_num_trials = 300
    oopt = orc.tv_default_options(); oopt.max_num_trials = 300
    base, inl0 = ver.verify_pairs(pairs, offs, np.concatenate(ms), opt, seeds)
    monkeypatch.setenv("B2_VERIFY_BPS", "222")
    res, inl = ver.verify_pairs(pairs, offs, np.concatenate(ms), opt, seeds)
    monkeypatch.delenv("B2_VERIFY_BPS")
    assert res.tobytes() == base.tobytes() and (inl == inl0).all()        # bit-identical to the production instance
    for i in range(len(specs)):
        c = orc.make_camera(prior=priors[i])
        r, oi = orc.two_view(c, kps[2 * i], c, kps[2 * i + 1], ms[i], oopt, seed=int(seeds[i]))
        assert (res["config"][i], res["n_inliers"][i], res["E_num_trials"][i], res["F_num_trials"][i], res["H_num_trials"][i]) == \
               (r.config, r.n_inliers, r.E_trials, r.F_trials, r.H_trials)
        assert inl[offs[i]:offs[i] + r.n_inliers].tolist() == oi.tolist()


def test_pair_with_more_matches_than_the_shared_memory_sampler_holds(ver):
    """A pair with more than 2 048 matches (kIdxSmem): the sampler's persistent index vector lives in the global scratch
    instead of shared memory -- the warp-parallel generator, the swaps and every decision must still equal the oracle's, bit
    for bit (images of 4 096 keypoints produce such pairs in the C3 workload)."""
    from dagsfm_b200 import Camera, TwoViewOptions
    rng = np.random.default_rng(77)
    kps, pairs, offs, ms, priors = [], [], [0], [], []
    for i, (n_in, n_out, planar, prior) in enumerate([(1900, 300, False, True), (1500, 700, True, False)]):
        p1, p2 = scene(rng, n_in, n_out, planar=planar, noise=0.4)
        kps += [p1, p2]
        ms.append(np.stack([np.arange(len(p1))] * 2, 1).astype(np.uint32))
        priors.append(prior); pairs.append((2 * i, 2 * i + 1)); offs.append(offs[-1] + len(p1))
    assert all(len(m) > 2048 for m in ms)
    cams = [Camera.make(prior_focal=p) for p in priors for _ in range(2)]
    ver.set_images(cams, kps)
    seeds = np.array([123, 456], np.uint32)
    opt = TwoViewOptions.default(); opt.max_num_trials = 200
    oopt = orc.tv_default_options(); oopt.max_num_trials = 200
    res, inl = ver.verify_pairs(pairs, offs, np.concatenate(ms), opt, seeds)
    for i in range(2):
        c = orc.make_camera(prior=priors[i])
        with orc.solver_stack(1):
            r, oi = orc.two_view(c, kps[2 * i], c, kps[2 * i + 1], ms[i], oopt, seed=int(seeds[i]))
        g = res[i]
        for m in ("E", "F", "H"):
            assert np.array_equal(np.array(getattr(r, m)[:]).view(np.uint64), g[m].view(np.uint64)), (i, m)
        assert (g["config"], g["n_inliers"], g["E_num_trials"], g["F_num_trials"], g["H_num_trials"]) == \
               (r.config, r.n_inliers, r.E_trials, r.F_trials, r.H_trials)
        assert inl[offs[i]:offs[i] + r.n_inliers].tolist() == oi.tolist()
        assert g["n_inliers"] > 1000


def test_edge_cases_empty_tiny_and_invalid_pairs(ver):
    from dagsfm_b200 import Camera, TwoViewOptions
    rng = np.random.default_rng(4)
    p1, p2 = scene(rng, 40, 5, noise=0.3)
    kps = [p1, p2, np.zeros((0, 2)), p1[:3]]
    ver.set_images([Camera.make(prior_focal=False)] * 4, kps)
    full = np.stack([np.arange(45)] * 2, 1).astype(np.uint32)
    pairs = [(0, 1), (0, 1), (0, 1), (3, 1)]
    ms = [full, full[:0], full[:6], full[:3]]                    # all, none, fewer than any minimal sample needs, 3
    offs = np.r_[0, np.cumsum([len(m) for m in ms])]
    opt = TwoViewOptions.default(); opt.max_num_trials = 200
    oopt = orc.tv_default_options(); oopt.max_num_trials = 200
    seeds = np.array([1, 2, 3, 4], np.uint32)
    res, inl = ver.verify_pairs(pairs, offs, np.concatenate(ms), opt, seeds)
    cam = orc.make_camera(prior=False)
    for k, (i, j) in enumerate(pairs):
        r, oi = orc.two_view(cam, kps[i], cam, kps[j], ms[k], oopt, seed=int(seeds[k]))
        assert (res["config"][k], res["n_inliers"][k]) == (r.config, r.n_inliers)
    assert res["config"][1] == 1 and res["config"][2] == 1 and res["config"][3] == 1      # DEGENERATE
    with pytest.raises(RuntimeError):                                                     # image id outside the store
        ver.verify_pairs([(0, 9)], [0, 45], full, opt, seeds[:1])
    with pytest.raises(RuntimeError):                                                     # TwoViewGeometry::Options::Check
        bad = TwoViewOptions.default(); bad.min_num_inliers = -1
        ver.verify_pairs(pairs[:1], offs[:2], full, bad, seeds[:1])
    r0, _ = ver.verify_pairs([], [0], np.zeros((0, 2), np.uint32), opt, np.zeros(0, np.uint32))
    assert len(r0) == 0


# ---------------------------------------------------------------- relative pose (verify_pose.cu)
def test_relative_pose_kernel_equals_the_oracle(ver):
    """b2_verify_relative_pose on the emulated warp kernel against the oracle's EstimateWithRelativePose
    restatement (pinned to the reference's essential / homography / triangulation unit tests in
    tests/test_oracle_relative_pose.py): same candidate, same surviving points, quaternion / translation equal to
    rounding, the median triangulation angle the same element -- for E-based, H-based (planar), panoramic,
    uncalibrated-camera and degenerate pairs."""
    from tests.pose_cases import check_relative_pose_against_oracle
    check_relative_pose_against_oracle(ver)


def test_relative_pose_rejects_inconsistent_input(ver):
    from dagsfm_b200.verification import RESULT_DTYPE
    res = np.zeros(1, dtype=RESULT_DTYPE)
    res["config"], res["n_inliers"] = 2, 50          # more inliers than matches
    with pytest.raises(RuntimeError):
        ver.relative_pose([(0, 1)], [0, 10], res, np.zeros((10, 2), np.uint32))


def test_relative_pose_device_flags_inconsistent_inlier_lists(ver):
    """The device entry cannot be validated on the host: the kernel itself refuses an inlier list longer than the pair's
    slot or naming a keypoint the image does not have (default pose for that pair, B2_ERR_INVALID for the call), and
    still serves the consistent pairs of the same call."""
    from dagsfm_b200.verification import POSE_DTYPE, RESULT_DTYPE, Camera
    rng = np.random.default_rng(8)
    p1, p2 = scene(rng, 40, 0, noise=0.0)
    ver.set_images([Camera.make(prior_focal=True)] * 2, [p1, p2])
    pairs = np.array([(0, 1), (0, 1), (0, 1)], np.uint32)
    offs = np.array([0, 40, 80, 120], np.int64)
    inl = np.tile(np.stack([np.arange(40)] * 2, 1), (3, 1)).astype(np.uint32)
    res = np.zeros(3, RESULT_DTYPE)
    res["config"] = 2
    res["n_inliers"] = [40, 41, 40]                 # pair 1: one more inlier than its slot holds
    res["E"][:] = np.array([0, 0, 0, 0, 0, -1, 0, 1, 0], float)
    pose = np.zeros(3, POSE_DTYPE)
    with pytest.raises(RuntimeError, match="inlier list inconsistent"):
        ver.relative_pose_device(3, pairs.ctypes.data, offs.ctypes.data, res.ctypes.data, inl.ctypes.data, pose.ctypes.data)
    res["n_inliers"] = 40
    inl[85, 1] = len(p2)                             # pair 2: a keypoint index one past the image's last
    with pytest.raises(RuntimeError, match="inlier list inconsistent"):
        ver.relative_pose_device(3, pairs.ctypes.data, offs.ctypes.data, res.ctypes.data, inl.ctypes.data, pose.ctypes.data)
    assert pose["n_points3D"][2] == 0 and (pose["qvec"][2] == 0).all()
    inl[85, 1] = 5
    ver.relative_pose_device(3, pairs.ctypes.data, offs.ctypes.data, res.ctypes.data, inl.ctypes.data, pose.ctypes.data)


def test_relative_pose_random_geometries(ver):
    """Seeded differential fuzz of relative_pose_kernel against the oracle on hand-made verification results: exact and
    perturbed essential matrices, homographies of planes and of pure rotations, inlier counts around the warp size (the
    ballot compaction and the even / odd median), points behind a camera, every camera layout."""
    from dagsfm_b200.verification import Camera, RESULT_DTYPE
    rng = np.random.default_rng(21)

    def euler(a):
        cx, sx, cy, sy, cz, sz = np.cos(a[0]), np.sin(a[0]), np.cos(a[1]), np.sin(a[1]), np.cos(a[2]), np.sin(a[2])
        return (np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @
                np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]]))
    cams, ocams, kps, pairs, offs, inl_all, res = [], [], [], [], [0], [], []
    models = [(2, [1200.0, 500.0, 500.0, 0.02]), (1, [1180.0, 1220.0, 510.0, 490.0]), (0, [1200.0, 500.0, 500.0]),
              (4, [1190.0, 1205.0, 500.0, 505.0, -0.05, 0.01, 0.001, -0.001]), (9, [1200.0, 500.0, 500.0, 0.03, -0.01])]
    counts = [0, 1, 2, 3, 15, 31, 32, 33, 63, 64, 65, 100, 129]
    n_cases = 48
    for k in range(n_cases):
        model, params = models[k % len(models)]
        n = counts[k % len(counts)]
        kind = ["E", "E_noisy", "H_plane", "H_rot", "E_behind"][k % 5]
        R = euler(rng.uniform(-0.3, 0.3, 3))
        t = rng.normal(size=3); t /= np.linalg.norm(t)
        X = rng.uniform(-1, 1, (max(n, 1), 3)) * [2, 2, 1] + [0, 0, 7]
        if kind == "H_plane":
            X[:, 2] = 7 + 0.1 * X[:, 0] - 0.05 * X[:, 1]
        if kind == "E_behind":
            X[::3, 2] *= -1           # a third of the points behind the first camera
        if kind == "H_rot":
            t = np.zeros(3)
        Xc = X @ R.T + t
        ocam = orc.make_camera(model=model, params=params, prior=(k % 7 != 6))
        cam = Camera.make(model=model, params=params, prior_focal=(k % 7 != 6))
        p1 = orc.world_to_image(ocam, X[:, :2] / X[:, 2:])
        p2 = orc.world_to_image(ocam, Xc[:, :2] / Xc[:, 2:])
        r = np.zeros(1, RESULT_DTYPE)
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        E = tx @ R
        if kind == "E_noisy":
            E = E + rng.normal(0, 1e-3, (3, 3))
        K = np.array([[params[0], 0, params[2 if model in (1, 4) else 1]], [0, params[1 if model in (1, 4) else 0], params[3 if model in (1, 4) else 2]], [0, 0, 1.0]])
        if kind == "H_plane":
            # plane n.X = d in camera 1: z = 7 + 0.1 x - 0.05 y  <=>  (-0.1, 0.05, 1).X = 7
            nrm = np.array([-0.1, 0.05, 1.0])
            Hn = R + np.outer(t, nrm) / 7.0
            H = K @ Hn @ np.linalg.inv(K)
        else:
            H = K @ R @ np.linalg.inv(K)
        r["E"][0], r["H"][0] = E.ravel(), H.ravel()
        r["config"][0] = {"E": 2, "E_noisy": 3, "H_plane": 6, "H_rot": 6, "E_behind": 2}[kind] if k % 11 != 10 else 7
        r["n_inliers"][0] = n
        perm = rng.permutation(max(n, 1))
        kp2 = np.empty_like(p2); kp2[perm] = p2
        inl = np.stack([np.arange(n), perm[:n]], 1).astype(np.uint32) if n else np.zeros((0, 2), np.uint32)
        cams += [cam, cam]; ocams += [ocam, ocam]; kps += [p1, kp2]
        pairs.append((2 * k, 2 * k + 1))
        pad = int(rng.integers(0, 4))                         # the inlier list is a prefix of the pair's match slice
        inl_all.append(np.vstack([inl, np.zeros((pad, 2), np.uint32)]))
        offs.append(offs[-1] + n + pad)
        res.append(r)
    res = np.concatenate(res)
    inl_all = np.concatenate(inl_all)
    ver.set_images(cams, kps)
    poses = ver.relative_pose(pairs, offs, res, inl_all)
    n_checked = 0
    for k in range(n_cases):
        a, n = offs[k], int(res["n_inliers"][k])
        if not ocams[2 * k].has_prior_focal:
            assert poses["qvec"][k].tolist() == [0, 0, 0, 0] and poses["config"][k] == res["config"][k]
            continue
        exp = orc.relative_pose(ocams[2 * k], kps[2 * k], ocams[2 * k + 1], kps[2 * k + 1], int(res["config"][k]),
                                res["E"][k].reshape(3, 3), res["H"][k].reshape(3, 3), inl_all[a:a + n])
        assert poses["config"][k] == exp.config and poses["n_points3D"][k] == exp.n_points3D, k
        assert np.abs(poses["qvec"][k] - np.array(exp.qvec)).max() < 1e-11, k
        assert np.abs(poses["tvec"][k] - np.array(exp.tvec)).max() < 1e-11, k
        assert poses["tri_angle"][k] == pytest.approx(exp.tri_angle, abs=1e-12), k
        n_checked += 1
    assert n_checked >= 35 and (poses["config"] == 5).any() and (poses["config"] == 4).any()


def test_device_pointer_chain_verify_then_relative_pose(ver):
    """b2_verify_pairs_device -> b2_verify_relative_pose_device on the same buffers (device pointers are host pointers on
    the emulator) equals the host-buffer calls: the device-resident chain of SURVEY 8f rank 1 extended by the pose step."""
    from dagsfm_b200.verification import POSE_DTYPE, RESULT_DTYPE, Camera, TwoViewOptions
    rng = np.random.default_rng(4)
    kps, pairs, offs, matches = [], [], [0], []
    for k in range(4):
        p1, p2 = scene(rng, 60 + 15 * k, 12, planar=(k == 2), noise=0.4)
        kps += [p1, p2]
        pairs.append((2 * k, 2 * k + 1))
        matches.append(np.stack([np.arange(len(p1))] * 2, 1))
        offs.append(offs[-1] + len(p1))
    pairs = np.array(pairs, np.uint32)
    offs = np.array(offs, np.int64)
    matches = np.concatenate(matches).astype(np.uint32)
    seeds = np.arange(4, dtype=np.uint32) + 40
    ver.set_images([Camera.make(prior_focal=True)] * 8, kps)
    vo = TwoViewOptions.default()
    res_h, inl_h = ver.verify_pairs(pairs, offs, matches, vo, seeds)
    pose_h = ver.relative_pose(pairs, offs, res_h, inl_h)
    res_d = np.zeros(4, RESULT_DTYPE)
    inl_d = np.zeros_like(matches)
    pose_d = np.zeros(4, POSE_DTYPE)
    ver.verify_pairs_device(4, pairs.ctypes.data, offs.ctypes.data, matches.ctypes.data, vo, seeds.ctypes.data,
                            res_d.ctypes.data, inl_d.ctypes.data)
    ver.relative_pose_device(4, pairs.ctypes.data, offs.ctypes.data, res_d.ctypes.data, inl_d.ctypes.data, pose_d.ctypes.data)
    assert res_d.tobytes() == res_h.tobytes() and pose_d.tobytes() == pose_h.tobytes()
    assert (pose_d["config"] == [2, 2, 4, 2]).all() and (pose_d["n_points3D"] > 50).all()


def test_bench_verification_leg_runs_on_the_emulated_library(ver):
    """bench.py's verification leg, with the opt-in relative-pose timing, end to end on the emulated library."""
    import sys
    import types
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        import bench
    finally:
        sys.argv = argv
    a = types.SimpleNamespace(verify_pairs=6, no_cpu=True, verify_pose=True)
    out = bench.bench_verify(a, 0, 0, 1, 1, lambda: None)
    assert out["pairs"] == 6 and out["pairs_per_s_e2e"] > 0 and "error" not in out["roofline"]
    assert "error" not in out["relative_pose"] and out["relative_pose"]["pairs_per_s_e2e"] > 0
    assert 0 <= out["relative_pose"]["pairs_with_pose"] <= 6

