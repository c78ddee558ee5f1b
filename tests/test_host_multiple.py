"""TwoViewGeometry::EstimateMultiple (two_view_geometry.cc:128-167) on the CPU: the round loop the
product runs around the GPU kernel (dagsfm_b200/csrc/verify_multiple.h, compiled by
tests/cpp/host_multiple.cc) with the ORACLE plugged in as the batched estimator, against a direct
restatement of the reference function (oracle.pyoracle.two_view_multiple)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.tv_scene import scene

HERE = os.path.dirname(os.path.abspath(__file__))
CB = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_int64, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64),
                      ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(orc.OrcTvResult),
                      ctypes.POINTER(ctypes.c_uint32))


@pytest.fixture(scope="module")
def hm():
    src = os.path.join(HERE, "cpp", "host_multiple.cc")
    out = os.path.join(HERE, "cpp", "_host_multiple.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O1", "-std=c++17", "-shared", "-fPIC", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    vp = ctypes.c_void_p
    lib.host_estimate_multiple.argtypes = [ctypes.c_int64, vp, vp, vp, ctypes.c_int, CB, vp, vp]
    return lib


def two_motion_pair(rng, n_a, n_b, n_out):
    """Matches of two independently moving rigid parts plus outliers (what multiple_models is for)."""
    a1, a2 = scene(rng, n_a, 0, noise=0.3)
    b1, b2 = scene(rng, n_b, n_out, noise=0.3, ang=-0.25, t=(0.3, 0.8, -0.1))
    p1, p2 = np.r_[a1, b1], np.r_[a2, b2]
    perm = rng.permutation(len(p2))
    matches = np.stack([np.arange(len(p1)), np.argsort(perm)], 1).astype(np.uint32)
    return p1, p2[perm], matches[rng.permutation(len(matches))]


def run_loop(lib, cam, kps, pairs, offs, matches, seeds, ignore_wm=True, opt=None):
    opt = opt or orc.tv_default_options()
    calls = []

    def estimate(na, ids, off, m, sd, res, inl):
        calls.append(na)
        for k in range(na):
            p = ids[k]
            lo, hi = off[k], off[k + 1]
            mk = np.array([[m[2 * i], m[2 * i + 1]] for i in range(lo, hi)], dtype=np.uint32).reshape(-1, 2)
            r, oi = orc.two_view(cam, kps[pairs[p][0]], cam, kps[pairs[p][1]], mk, opt, seed=int(sd[k]))
            res[k] = r
            for i, (x, y) in enumerate(oi):
                inl[2 * (lo + i)], inl[2 * (lo + i) + 1] = int(x), int(y)
        return 0

    n = len(pairs)
    off = np.ascontiguousarray(offs, np.int64)
    mt = np.ascontiguousarray(matches, np.uint32).reshape(-1, 2)
    sd = np.ascontiguousarray(seeds, np.uint32)
    res = (orc.OrcTvResult * n)()
    inl = np.zeros((max(len(mt), 1), 2), np.uint32)
    rc = lib.host_estimate_multiple(n, off.ctypes.data, mt.ctypes.data, sd.ctypes.data, int(ignore_wm), CB(estimate),
                                    ctypes.cast(res, ctypes.c_void_p), inl.ctypes.data)
    assert rc == 0
    return res, inl, calls


def test_round_loop_equals_reference_restatement(hm):
    rng = np.random.default_rng(3)
    cam = orc.make_camera(prior=False)
    kps, pairs, offs, ms = [], [], [0], []
    specs = [(160, 120, 40), (200, 0, 60), (90, 90, 90), (12, 0, 0), (150, 100, 0)]   # two motions, one, two, too few, two
    for k, (na, nb, no) in enumerate(specs):
        if nb:
            p1, p2, m = two_motion_pair(rng, na, nb, no)
        else:
            p1, p2 = scene(rng, na, no, noise=0.3)
            m = np.stack([np.arange(len(p1))] * 2, 1).astype(np.uint32)
        kps += [p1, p2]
        pairs.append((2 * k, 2 * k + 1))
        ms.append(m)
        offs.append(offs[-1] + len(m))
    seeds = np.arange(5, dtype=np.uint32) + 40
    res, inl, calls = run_loop(hm, cam, kps, pairs, offs, np.concatenate(ms), seeds)
    configs = []
    for k, (i, j) in enumerate(pairs):
        cfg, geos, exp_inl = orc.two_view_multiple(cam, kps[i], cam, kps[j], ms[k], seed=int(seeds[k]))
        configs.append(cfg)
        assert res[k].config == cfg and res[k].n_inliers == len(exp_inl)
        assert inl[offs[k]:offs[k] + len(exp_inl)].tolist() == exp_inl.tolist()
        if len(geos) == 1:      # exactly the plain Estimate of round 0
            r0, i0 = orc.two_view(cam, kps[i], cam, kps[j], ms[k], seed=int(seeds[k]))
            assert (res[k].config, res[k].n_inliers, list(res[k].F)) == (r0.config, r0.n_inliers, list(r0.F))
        if cfg == 8:
            assert list(res[k].E) == [0.0] * 9 and list(res[k].F) == [0.0] * 9 and list(res[k].H) == [0.0] * 9
            assert len(exp_inl) > max(g.n_inliers for g in geos)
    assert configs.count(8) >= 2 and configs[3] == 1 and configs[1] in (2, 3)
    assert calls[0] == 5 and len(calls) >= 3 and calls == sorted(calls, reverse=True)   # pairs drop out round by round


def test_watermark_geometries_are_ignored_or_kept(hm):
    rng = np.random.default_rng(5)
    cam = orc.make_camera(prior=False)
    b = rng.uniform(0, 1000, 60)
    wm1 = np.stack([b, rng.uniform(0, 60, 60)], 1)
    wm2 = wm1 + [3.0, 2.0]
    m = np.stack([np.arange(60)] * 2, 1).astype(np.uint32)
    for ignore, exp_cfg in ((True, 1), (False, 7)):
        res, inl, _ = run_loop(hm, cam, [wm1, wm2], [(0, 1)], [0, 60], m, [3], ignore_wm=ignore)
        cfg, geos, exp_inl = orc.two_view_multiple(cam, wm1, cam, wm2, m, seed=3, multiple_ignore_watermark=ignore)
        assert cfg == exp_cfg == res[0].config
        assert res[0].n_inliers == len(exp_inl)
