"""Parity of the CUDA two-view verifier (through the C ABI) with the CPU oracle.

Levels: (1) the sampler's PRNG index stream -- bit exact; (2) minimal / local
solvers -- bit-identical to the oracle's device-order solver stack (one floating-point stack on
both sides: oracle/twoview_oracle.cc "second solver stack"), and within 1e-7 of its independent
stack (sequential sums, pinned to the reference's Matlab goldens); (3) model scoring -- counts,
masks and the ordered residual sum bit exact; (4) the whole TwoViewGeometry::Estimate decision
for seeded pairs -- configuration, every inlier / trial count, the inlier match list AND the
E / F / H matrices bit for bit identical to the oracle (device-order stack) on EVERY pair,
noise-free and noisy; the agreement rate with the independent stack is printed beside it."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.tv_scene import scene

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ver():
    from dagsfm_b200 import TwoViewGeometryVerifier
    v = TwoViewGeometryVerifier(0)
    yield v
    v.close()


@pytest.mark.parametrize("total,k", [(50, 7), (7, 7), (1000, 5), (33, 4), (20, 1)])
def test_sample_stream_bit_exact(ver, total, k):
    for seed in (0, 1, 12345, 2**32 - 1):
        got = ver.debug_sample_stream(seed, total, k, 300)
        exp = orc.sample_stream(seed, total, k, 300)
        assert (got == exp).all()


def _match_models(got, exp, tol):
    assert len(got) == len(exp)
    for g, e in zip(got, exp):
        d = min(np.abs(g - e).max(), np.abs(g + e).max())
        assert d < tol * max(1.0, np.abs(e).max()), (d, g, e)


def test_minimal_solvers_vs_oracle(ver):
    rng = np.random.default_rng(0)
    for it in range(20):
        p1, p2 = scene(rng, 8, 0, noise=0.5)
        _match_models(ver.debug_solve(1, p1[:7], p2[:7]), orc.f7(p1[:7], p2[:7]), 1e-7)
        _match_models(ver.debug_solve(2, p1[:4], p2[:4]), [orc.h_dlt(p1[:4], p2[:4])], 1e-7)
        n1, n2 = (p1 - 500) / 1200, (p2 - 500) / 1200
        _match_models(ver.debug_solve(0, n1[:5], n2[:5]), orc.e5(n1[:5], n2[:5]), 1e-6)


def test_reference_golden_vectors_on_device(ver):
    # the same Matlab goldens the oracle is pinned to (fundamental_matrix_test.cc:39-105)
    from tests.test_oracle_twoview import P1_7, P2_7, P1_8, P2_8
    F = ver.debug_solve(1, P1_7, P2_7)[0]
    exp = np.array([[4.81441976, -8.16978909, 6.73133404], [5.16247992, 0.19325606, -2.87239381],
                    [-9.92570126, 3.64159554, 1.0]])
    assert np.allclose(F, exp, rtol=1e-8)
    F8 = ver.debug_solve(3, P1_8, P2_8)[0]
    exp8 = np.array([[-0.217859, 0.419282, -0.0343075], [-0.0717941, 0.0451643, 0.0216073],
                     [0.248062, -0.429478, 0.0221019]])
    assert np.abs(F8 - exp8).max() < 1e-5


def test_local_estimators_vs_oracle(ver):
    rng = np.random.default_rng(1)
    for n in (8, 9, 40, 333):
        p1, p2 = scene(rng, n, 0, noise=0.5)
        _match_models(ver.debug_solve(3, p1, p2), [orc.eight_point(p1, p2)], 1e-7)
        _match_models(ver.debug_solve(2, p1, p2), [orc.h_dlt(p1, p2)], 1e-7)
        n1, n2 = (p1 - 500) / 1200, (p2 - 500) / 1200
        _match_models(ver.debug_solve(0, n1, n2), orc.e5(n1, n2), 1e-5)


def test_score_models_bit_exact(ver):
    rng = np.random.default_rng(2)
    p1, p2 = scene(rng, 700, 300, noise=1.0)
    F = orc.eight_point(p1[:700], p2[:700])
    Fs = np.stack([F, F * 3.0, orc.f7(p1[:7], p2[:7])[0], rng.normal(size=(3, 3))])
    for typ, models, thr in ((1, Fs, 16.0), (2, np.stack([orc.h_dlt(p1[:50], p2[:50]), np.eye(3)]), 16.0)):
        counts, sums, masks = ver.score_models(typ, p1, p2, models, thr)
        for k, M in enumerate(models):
            r = orc.residuals(typ, p1, p2, M)
            m = r <= thr
            s = 0.0
            for v in r[m]:
                s += v            # sequential sum in index order (support_measurement.cc:43-46)
            assert counts[k] == m.sum()
            assert (masks[k] == m).all()
            assert sums[k] == s   # bit exact


def _run_both(ver, cams_prior, scenes, seeds, opt_kw=None, stack=1):
    from dagsfm_b200 import Camera, TwoViewOptions
    n = len(scenes)
    cams, kps, pairs, offs, ms = [], [], [], [0], []
    for i, (p1, p2) in enumerate(scenes):
        cams += [Camera.make(prior_focal=cams_prior[i]), Camera.make(prior_focal=cams_prior[i])]
        perm = np.random.default_rng(100 + i).permutation(len(p2))
        kps += [p1, p2[perm]]
        inv = np.argsort(perm)
        ms.append(np.stack([np.arange(len(p1)), inv], 1))
        pairs.append((2 * i, 2 * i + 1))
        offs.append(offs[-1] + len(p1))
    ver.set_images(cams, kps)
    opt = TwoViewOptions.default()
    oopt = orc.tv_default_options()
    for k, v in (opt_kw or {}).items():
        setattr(opt, k, v)
        setattr(oopt, k, v)
    res, inl = ver.verify_pairs(pairs, offs, np.concatenate(ms), opt, seeds)
    out = []
    with orc.solver_stack(stack):
        for i in range(n):
            c = orc.make_camera(prior=cams_prior[i])
            r, oi = orc.two_view(c, kps[2 * i], c, kps[2 * i + 1], ms[i], oopt, seed=int(seeds[i]))
            gi = inl[offs[i]:offs[i] + res["n_inliers"][i]]
            out.append((res[i], gi, r, oi))
    return out


def _same(g, gi, r, oi):
    return (g["config"] == r.config and g["n_inliers"] == r.n_inliers and g["E_num_inliers"] == r.E_inl and
            g["F_num_inliers"] == r.F_inl and g["H_num_inliers"] == r.H_inl and gi.tolist() == oi.tolist())


def _same_bits(g, gi, r, oi):
    """Everything: decisions, trial counts, inlier list, and the three model matrices bit for bit."""
    return (_same(g, gi, r, oi) and g["E_num_trials"] == r.E_trials and g["F_num_trials"] == r.F_trials and
            g["H_num_trials"] == r.H_trials and
            all(np.array_equal(np.array(getattr(r, m)[:]).view(np.uint64), g[m].view(np.uint64)) for m in ("E", "F", "H")))


def test_two_view_exact_on_well_separated_data(ver):
    # noise-free inliers + gross outliers: the reference's own style of RANSAC test
    # (loransac_test.cc:57-107 expects the exact inlier mask under a fixed seed)
    rng = np.random.default_rng(3)
    scenes, prior = [], []
    for i in range(12):
        scenes.append(scene(rng, 150 + 10 * i, 60 + 5 * i, planar=(i % 4 == 3), noise=0.0))
        prior.append(i % 2 == 0)
    seeds = np.arange(12) * 7 + 1
    out = _run_both(ver, prior, scenes, seeds)
    for k, (g, gi, r, oi) in enumerate(out):
        assert _same(g, gi, r, oi), (k, g, (r.config, r.n_inliers, r.E_inl, r.F_inl, r.H_inl))
        assert g["E_num_trials"] == r.E_trials and g["F_num_trials"] == r.F_trials and g["H_num_trials"] == r.H_trials
        if g["config"] in (2, 3):
            assert g["n_inliers"] >= 150 + 10 * k
    cfgs = [int(o[0]["config"]) for o in out]
    assert 2 in cfgs and 3 in cfgs and 6 in cfgs


def test_two_view_identical_on_noisy_data(ver):
    """north_star: bit-exact inlier masks.  60 noisy seeded pairs: every field, the inlier list and the E / F / H
    matrices are bit-identical to the oracle evaluated in the same operation order (device-order stack)."""
    rng = np.random.default_rng(4)
    scenes = [scene(rng, 120 + (i * 37) % 200, 80 + (i * 13) % 150, planar=(i % 5 == 0), noise=0.7) for i in range(60)]
    prior = [i % 3 != 0 for i in range(60)]
    out = _run_both(ver, prior, scenes, np.arange(60) + 1000)
    bad = [k for k, o in enumerate(out) if not _same_bits(*o)]
    assert not bad, f"pairs {bad} differ from the oracle (device-order stack)"
    # the oracle's independent stack (sequential sums) on the same pairs: informational agreement rate, equivalent answers
    out0 = _run_both(ver, prior, scenes, np.arange(60) + 1000, stack=0)
    same0 = sum(_same(*o) for o in out0)
    print(f"\nverification parity on noisy data: 60/60 bit-identical to the device-order stack; "
          f"{same0}/60 decisions identical to the independent stack")
    assert same0 >= 54
    for g, gi, r, oi in out0:
        assert abs(int(g["n_inliers"]) - r.n_inliers) <= max(3, 0.05 * r.n_inliers)


def test_two_view_identical_on_a_large_seeded_batch(ver):
    """2 000 pairs of the bench's workload family (20-420 matches, inlier ratios 0.1-0.95, planar scenes, with and
    without prior focal length): 100 % identical, matrices included."""
    import ctypes as C
    from dagsfm_b200 import Camera, TwoViewOptions
    from tests.tv_scene import make_pairs
    n = 2000
    w = make_pairs(n, n_in=(20, 220), n_out=(10, 200), seed=5, noise=0.7)
    cams = [Camera.make(params=w["cam_params"], prior_focal=bool(p)) for p in w["prior"]]
    ver.set_images(cams, w["keypoints"])
    seeds = (np.arange(n) * 2654435761 % (2 ** 32)).astype(np.uint32)
    res, inl = ver.verify_pairs(w["pairs"], w["match_offsets"], w["matches"], TwoViewOptions.default(), seeds)
    ocams = (orc.OrcCamera * (2 * n))(*[orc.make_camera(params=w["cam_params"], prior=bool(p)) for p in w["prior"]])
    ptrs = (C.c_void_p * (2 * n))(*[k.ctypes.data for k in w["keypoints"]])
    oo = orc.tv_default_options()
    ores = (orc.OrcTvResult * n)()
    oinl = np.zeros((int(w["match_offsets"][-1]), 2), np.uint32)
    with orc.solver_stack(1):
        orc._tv().orc_two_view_pairs_mt2(C.cast(ocams, C.c_void_p), C.cast(ptrs, C.c_void_p), w["pairs"].ctypes.data, n,
                                         w["match_offsets"].ctypes.data, w["matches"].ctypes.data, C.byref(oo),
                                         seeds.ctypes.data, 8, C.cast(ores, C.c_void_p), oinl.ctypes.data)
    off = w["match_offsets"]
    bad = [i for i in range(n) if not _same_bits(res[i], inl[off[i]:off[i] + max(res["n_inliers"][i], 0)], ores[i],
                                                 oinl[off[i]:off[i] + max(ores[i].n_inliers, 0)])]
    assert not bad, f"{len(bad)} of {n} pairs differ, first {bad[:5]}"
    hist = np.bincount(res["config"], minlength=8)
    assert hist[2] > 0 and hist[3] > 0 and (hist[4] + hist[5] + hist[6]) > 0   # calibrated, uncalibrated and planar paths ran


def test_degenerate_and_watermark_paths(ver):
    rng = np.random.default_rng(5)
    few = scene(rng, 10, 0)                      # < min_num_inliers -> DEGENERATE
    junk = (rng.uniform(0, 1000, (200, 2)), rng.uniform(0, 1000, (200, 2)))  # no geometry
    # pure translation of points that all lie in the image border -> WATERMARK
    b = np.r_[rng.uniform(0, 1000, (60, 1)) * [1], ].ravel()
    wm1 = np.stack([b, rng.uniform(0, 60, 60)], 1)
    wm2 = wm1 + [3.0, 2.0]
    out = _run_both(ver, [True, True, False], [few, junk, (wm1, wm2)], np.array([1, 2, 3]))
    for g, gi, r, oi in out:
        assert _same(g, gi, r, oi)
    assert out[0][0]["config"] == 1   # too few matches; the junk pair only has to agree with the oracle
    assert out[2][0]["config"] == 7

