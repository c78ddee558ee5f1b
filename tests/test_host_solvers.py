"""The minimal solvers of the verification kernel, compiled for the host from the same header
(tests/cpp/host_solvers.cc), against the oracle's restatement of the reference estimators
(fundamental_matrix.cc:43-120, homography_matrix.cc:54-95, essential_matrix.cc:60-178).

The device solvers take the constraint matrix's null space from a Householder QR, the oracle from a
Jacobi SVD like the reference: the models they derive are normalised, so both must agree to solver
tolerance (sign of a homogeneous matrix is free)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.tv_scene import scene

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hs():
    src = os.path.join(HERE, "cpp", "host_solvers.cc")
    out = os.path.join(HERE, "cpp", "_host_solvers.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, src], check=True)
    lib = ctypes.CDLL(out)
    dp = ctypes.POINTER(ctypes.c_double)
    for f in (lib.host_f7, lib.host_e5, lib.host_h4):
        f.argtypes = [dp, dp, dp]
        f.restype = ctypes.c_int
    return lib


def _call(fn, p1, p2):
    p1 = np.ascontiguousarray(p1, np.float64)
    p2 = np.ascontiguousarray(p2, np.float64)
    m = np.zeros(90)
    dp = ctypes.POINTER(ctypes.c_double)
    n = fn(p1.ctypes.data_as(dp), p2.ctypes.data_as(dp), m.ctypes.data_as(dp))
    return [m[9 * i:9 * i + 9].reshape(3, 3).copy() for i in range(n)]


def _match(got, exp, tol):
    assert len(got) == len(exp)
    for g, e in zip(got, exp):
        d = min(np.abs(g - e).max(), np.abs(g + e).max())
        assert d < tol * max(1.0, np.abs(e).max()), (d, g, e)


def _unit(m):
    return m / np.linalg.norm(m)


def test_minimal_solvers_match_oracle(hs):
    rng = np.random.default_rng(0)
    e5_diff = []
    for it in range(200):
        p1, p2 = scene(rng, 8, 0, noise=0.5)
        _match(_call(hs.host_f7, p1[:7], p2[:7]), orc.f7(p1[:7], p2[:7]), 1e-7)
        _match([_unit(m) for m in _call(hs.host_h4, p1[:4], p2[:4])], [_unit(orc.h_dlt(p1[:4], p2[:4]))], 1e-7)
        n1, n2 = (p1 - 500) / 1200, (p2 - 500) / 1200
        got, exp = _call(hs.host_e5, n1[:5], n2[:5]), orc.e5(n1[:5], n2[:5])
        assert len(got) == len(exp)
        e5_diff += [min(np.abs(g - e).max(), np.abs(g + e).max()) for g, e in zip(got, exp)]
    # the 5-point solver goes through a 10x20 elimination and a degree-10 root finder: a few
    # hypotheses per thousand are ill-conditioned (near-double roots) in ANY arithmetic order
    e5_diff = np.array(e5_diff)
    assert np.median(e5_diff) < 1e-12
    assert np.quantile(e5_diff, 0.99) < 1e-6
    assert e5_diff.max() < 1e-3


def test_reference_seven_point_golden(hs):
    from tests.test_oracle_twoview import P1_7, P2_7
    F = _call(hs.host_f7, P1_7, P2_7)[0]
    exp = np.array([[4.81441976, -8.16978909, 6.73133404], [5.16247992, 0.19325606, -2.87239381],
                    [-9.92570126, 3.64159554, 1.0]])
    assert np.allclose(F, exp, rtol=1e-8)


def test_degenerate_samples_do_not_blow_up(hs):
    # repeated / collinear points: rank-deficient constraint matrices
    p = np.array([[10.0, 10], [20, 20], [30, 30], [40, 40], [50, 50], [60, 60], [70, 70]])
    for fn, k in ((hs.host_f7, 7), (hs.host_h4, 4), (hs.host_e5, 5)):
        for m in _call(fn, p[:k], p[:k] + 1.0):
            assert m.shape == (3, 3)
    same = np.tile(np.array([[5.0, 7.0]]), (7, 1))
    for fn, k in ((hs.host_f7, 7), (hs.host_e5, 5)):
        _call(fn, same[:k], same[:k])

