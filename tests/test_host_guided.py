"""Guided matching on the CPU: the row function the CUDA kernel runs
(dagsfm_b200/csrc/match_guided.cuh) and the integer threshold tables every matcher kernel uses
(match_thresholds.h), compiled for the host (tests/cpp/host_guided.cc), against the oracle's
restatement of MatchGuidedSiftFeaturesCPU / FindBestMatchesOneWay (sift.cc:111-162, :824-875)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.tv_scene import scene

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hg():
    src = os.path.join(HERE, "cpp", "host_guided.cc")
    out = os.path.join(HERE, "cpp", "_host_guided.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    vp = ctypes.c_void_p
    lib.host_guided_match.argtypes = [vp, vp, vp, ctypes.c_int, vp, ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_double,
                                      ctypes.c_float, ctypes.c_float, ctypes.c_int, vp, ctypes.c_int]
    lib.host_threshold_accept.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int]
    return lib


def host_guided(lib, k1, k2, d1, d2, config, F=None, H=None, max_error=4.0, max_ratio=0.8, max_distance=0.7,
                cross_check=True):
    k1 = np.ascontiguousarray(k1, np.float32).reshape(-1, 2)
    k2 = np.ascontiguousarray(k2, np.float32).reshape(-1, 2)
    d1 = np.ascontiguousarray(d1, np.uint8).reshape(-1, 128)
    d2 = np.ascontiguousarray(d2, np.uint8).reshape(-1, 128)
    Fm = np.ascontiguousarray(np.eye(3) if F is None else F, np.float64)
    Hm = np.ascontiguousarray(np.eye(3) if H is None else H, np.float64)
    cap = max(len(d1), 1)
    out = np.zeros((cap, 2), np.uint32)
    n = lib.host_guided_match(k1.ctypes.data, k2.ctypes.data, d1.ctypes.data, len(d1), d2.ctypes.data, len(d2),
                              config, Fm.ctypes.data, Hm.ctypes.data, max_error, max_ratio, max_distance,
                              int(cross_check), out.ctypes.data, cap)
    return None if n == -1 else out[:n]


def _scene_with_descriptors(rng, n_in, n_out, planar):
    """Matched keypoints + outliers.  Half of the outliers of image 2 are DECOYS: they carry the
    descriptor of an inlier (repetitive structure), so the plain ratio test rejects that inlier while
    the guided filter removes the decoy (it is off the epipolar line / away from H x) and keeps it."""
    p1, p2 = scene(rng, n_in, n_out, planar=planar, noise=0.4)
    base = rng.gamma(0.6, 1.0, (n_in, 128)).astype(np.float32)
    d1 = np.stack([orc.l2_normalize_to_u8(v) for v in np.r_[base, rng.gamma(0.6, 1.0, (n_out, 128)).astype(np.float32)]])
    jit = np.maximum(base + rng.normal(0, 0.05, base.shape).astype(np.float32), 0)
    d2 = np.stack([orc.l2_normalize_to_u8(v) for v in np.r_[jit, rng.gamma(0.6, 1.0, (n_out, 128)).astype(np.float32)]])
    n_decoy = min(n_out // 2, n_in)
    d2[n_in:n_in + n_decoy] = d2[:n_decoy]
    perm = rng.permutation(len(p2))
    return p1, p2[perm], d1, d2[perm]


def test_reference_guided_test_cases(hg):
    # sift_test.cc:327-372
    d1 = orc.create_random_descriptors(2)
    d2 = d1[::-1].copy()
    k1 = np.array([[1, 0], [2, 0]], np.float32)
    k2 = np.array([[2, 0], [1, 0]], np.float32)
    for fn in (lambda *a, **k: orc.match_guided(*a, **k), lambda *a, **k: host_guided(hg, *a, **k)):
        assert fn(k1, k2, d1, d2, 6, H=np.eye(3)).tolist() == [[0, 1], [1, 0]]
        k1b = k1.copy()
        k1b[0, 0] = 100
        assert fn(k1b, k2, d1, d2, 6, H=np.eye(3)).tolist() == [[1, 0]]
        e, ek = np.zeros((0, 128), np.uint8), np.zeros((0, 2), np.float32)
        assert len(fn(ek, k2, e, d2, 6)) == 0 and len(fn(k1, ek, d1, e, 6)) == 0 and len(fn(ek, ek, e, e, 6)) == 0
        assert fn(k1, k2, d1, d2, 0) is None and fn(k1, k2, d1, d2, 7) is None   # UNDEFINED / WATERMARK: no filter


@pytest.mark.parametrize("planar", [False, True])
def test_device_row_function_equals_oracle_on_scenes(hg, planar):
    rng = np.random.default_rng(5 + planar)
    for it in range(6):
        k1, k2, d1, d2 = _scene_with_descriptors(rng, 120 + 30 * it, 60 + 10 * it, planar)
        if planar:
            M = orc.h_dlt(*_inlier_pairs(k1, k2, d1, d2))
            kw, cfg = dict(H=M), 4 + (it % 3)
        else:
            M = orc.eight_point(*_inlier_pairs(k1, k2, d1, d2))
            kw, cfg = dict(F=M), 2 + (it % 2)
        for opts in (dict(), dict(cross_check=False), dict(max_error=1.0), dict(max_ratio=0.95, max_distance=1.2)):
            exp = orc.match_guided(k1, k2, d1, d2, cfg, **kw, **opts)
            got = host_guided(hg, k1, k2, d1, d2, cfg, **kw, **opts)
            assert got.tolist() == exp.tolist()
        # the filter matters: the decoys make the unguided ratio test drop inliers the guided one keeps
        n_guided, n_plain = len(orc.match_guided(k1, k2, d1, d2, cfg, **kw)), len(orc.match_sift(d1, d2))
        assert n_guided >= 60 and n_guided > n_plain + 10
        assert len(orc.match_guided(k1, k2, d1, d2, cfg, **kw, max_error=0.01)) < 30


def _inlier_pairs(k1, k2, d1, d2):
    m = orc.match_sift(d1, d2)
    return k1[m[:, 0]].astype(np.float64), k2[m[:, 1]].astype(np.float64)


def test_integer_threshold_tables_equal_float_decisions(hg):
    # accept iff NOT(a(best) > max_distance) and NOT(a(best) >= max_ratio * a(second)), float32 (sift.cc:139-157)
    rng = np.random.default_rng(0)
    kn = np.float32(1.0) / (np.float32(512.0) * np.float32(512.0))
    for ratio, dist in ((0.8, 0.7), (0.6, 0.5), (0.95, 1.3), (1.0, 1.5707964), (0.3, 0.2)):
        r32, d32 = np.float32(ratio), np.float32(dist)
        best = np.r_[rng.integers(1, 300000, 4000), 262143, 262144, 262145, 1, 2]
        second = np.minimum(np.r_[rng.integers(0, 300000, 4000), 262144, 0, 1, 0, 1], best)
        near = rng.integers(1, 262144, 3000)
        best = np.r_[best, near]
        second = np.r_[second, np.maximum(near - rng.integers(0, 3000, 3000), 0)]
        for b, s in zip(best.tolist(), second.tolist()):
            ab = np.arccos(np.minimum(kn * np.float32(b), np.float32(1.0)), dtype=np.float32)
            asn = np.arccos(np.minimum(kn * np.float32(s), np.float32(1.0)), dtype=np.float32)
            exp = (not ab > d32) and (not ab >= r32 * asn)
            assert bool(hg.host_threshold_accept(ratio, dist, b, s)) == bool(exp), (ratio, dist, b, s)


# ----------------------------------------------------------------------------------------------
# The CUDA kernels themselves (match_guided_kernels.cuh), executed on the host block by block and
# thread by thread inside a re-statement of the library's pipeline (tests/cpp/host_guided_kernel.cc).
@pytest.fixture(scope="module")
def hk():
    src = os.path.join(HERE, "cpp", "host_guided_kernel.cc")
    out = os.path.join(HERE, "cpp", "_host_guided_kernel.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    vp, ci = ctypes.c_void_p, ctypes.c_int
    lib.host_guided_pipeline.argtypes = [ci, vp, vp, vp, ci, vp, vp, vp, vp, ctypes.c_double, ctypes.c_float,
                                         ctypes.c_float, ci, ci, vp, vp, ctypes.c_int64]
    return lib


def run_kernels(lib, descs, kps, pairs, geos, grid_blocks, max_error=4.0, max_ratio=0.8, max_distance=0.7, cross_check=True):
    n = len(descs)
    ds = [np.ascontiguousarray(d, np.uint8).reshape(-1, 128) for d in descs]
    ks = [np.ascontiguousarray(k, np.float32).reshape(-1, 2) for k in kps]
    dptr = (ctypes.c_void_p * n)(*[d.ctypes.data for d in ds])
    kptr = (ctypes.c_void_p * n)(*[k.ctypes.data for k in ks])
    cnt = np.array([len(d) for d in ds], np.int32)
    pr = np.ascontiguousarray(pairs, np.uint32).reshape(-1, 2)
    cfg = np.array([g[0] for g in geos], np.int32)
    F = np.ascontiguousarray([np.zeros(9) if g[1] is None else np.asarray(g[1], np.float64).reshape(9) for g in geos])
    H = np.ascontiguousarray([np.zeros(9) if g[2] is None else np.asarray(g[2], np.float64).reshape(9) for g in geos])
    off = np.zeros(len(pr) + 1, np.int64)
    cap = int(sum(len(ds[a]) for a, _ in pr)) + 1
    out = np.zeros((cap, 2), np.uint32)
    total = lib.host_guided_pipeline(n, dptr, kptr, cnt.ctypes.data, len(pr), pr.ctypes.data, cfg.ctypes.data,
                                     F.ctypes.data, H.ctypes.data, max_error, max_ratio, max_distance, int(cross_check),
                                     grid_blocks, off.ctypes.data, out.ctypes.data, cap)
    assert total == off[-1] >= 0
    return off, out


@pytest.mark.parametrize("grid_blocks", [1, 3, 64])
def test_cuda_kernels_on_host_equal_oracle(hk, grid_blocks):
    rng = np.random.default_rng(21)
    kps, descs, pairs, geos = [], [], [], []
    for k, (n_in, n_out) in enumerate([(120, 40), (260, 90), (300, 240), (90, 10)]):   # 160 .. 540 rows: 1-3 supertiles
        planar = k % 2 == 1
        k1, k2, d1, d2 = _scene_with_descriptors(rng, n_in, n_out, planar)
        a, b = _inlier_pairs(k1, k2, d1, d2)
        geos.append((4 + k, None, orc.h_dlt(a, b)) if planar else (2 + (k // 2), orc.eight_point(a, b), None))
        kps += [k1, k2[: len(k2) - 7 * k]]                 # ragged: the two images of a pair differ in size
        descs += [d1, d2[: len(d2) - 7 * k]]
        pairs.append((2 * k, 2 * k + 1))
    kps.append(np.zeros((0, 2), np.float32)); descs.append(np.zeros((0, 128), np.uint8))       # an empty image
    pairs += [(0, 8), (8, 1), (3, 2), (0, 1)]
    geos += [geos[0], geos[0], geos[1], (7, None, None)]    # empty image twice, swapped order, WATERMARK = no filter
    for opts in (dict(), dict(cross_check=False, max_error=2.0), dict(max_ratio=0.95, max_distance=1.2)):
        off, out = run_kernels(hk, descs, kps, pairs, geos, grid_blocks, **opts)
        for p, ((i, j), (cfg, F, H)) in enumerate(zip(pairs, geos)):
            e = orc.match_guided(kps[i], kps[j], descs[i], descs[j], cfg, F=F, H=H, **opts)
            exp = [] if (e is None or len(descs[i]) == 0 or len(descs[j]) == 0) else e.tolist()
            assert out[off[p]:off[p + 1]].tolist() == exp, (p, cfg, opts)
        assert off[4] > 500
