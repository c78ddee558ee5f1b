"""Robust losses in the BA oracle (BundleAdjustmentOptions::CreateLossFunction, bundle_adjustment.cc:53-68;
the mapper's local BA uses SOFT_L1, incremental_mapper_controller.cc:252-253): Ceres 1.14's SoftLOneLoss /
CauchyLoss + Corrector restated in oracle/ba_oracle.cc.  Checked against an independent numpy evaluation
of 1/2 sum rho(|r|^2): reported costs match and the converged point is stationary for THAT function."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests.ba_scene import R_from_quat, make_ba_problem


def robust_cost(p, typ, a):
    R = np.stack([R_from_quat(q / np.linalg.norm(q)) for q in p["qvec"]])
    i = p["obs_img"]
    Xc = np.einsum("nij,nj->ni", R[i], p["xyz"][p["obs_pt"]]) + p["tvec"][i]
    u, v = Xc[:, 0] / Xc[:, 2], Xc[:, 1] / Xc[:, 2]
    k = p["cam_params"][p["img_cam"][i]]
    d = 1 + k[:, 3] * (u * u + v * v)
    s = (k[:, 0] * u * d + k[:, 1] - p["obs_xy"][:, 0]) ** 2 + (k[:, 0] * v * d + k[:, 2] - p["obs_xy"][:, 1]) ** 2
    b = a * a
    rho = s if typ == 0 else 2 * b * (np.sqrt(1 + s / b) - 1) if typ == 1 else b * np.log(1 + s / b)
    return 0.5 * rho.sum()


@pytest.mark.parametrize("typ,scale", [(1, 1.0), (2, 1.0), (1, 2.5), (2, 2.0)])
def test_costs_and_stationarity(typ, scale):
    p = make_ba_problem(n_img=10, n_pts=200, track_len=5, seed=2, noise_px=1.0)
    rng = np.random.default_rng(1)
    idx = rng.choice(len(p["obs_xy"]), 40, replace=False)
    p["obs_xy"][idx] += rng.normal(0, 40, (40, 2))          # gross outliers: what the loss is for
    c0 = robust_cost(p, typ, scale)
    s = orc.ba_solve(p, max_num_iterations=200, gradient_tolerance=1e-10, function_tolerance=1e-12,
                     loss_type=typ, loss_scale=scale)
    c1 = robust_cost(p, typ, scale)
    assert abs(s.initial_cost - c0) <= 1e-10 * c0 and abs(s.final_cost - c1) <= 1e-10 * c1
    assert s.termination == 0 and c1 < 0.5 * c0
    # stationarity of the numpy cost at the oracle's optimum: central differences over point coordinates
    g = []
    for j in rng.choice(len(p["xyz"]), 12, replace=False):
        for a in range(3):
            h = 1e-6
            p["xyz"][j, a] += h
            cp = robust_cost(p, typ, scale)
            p["xyz"][j, a] -= 2 * h
            cm = robust_cost(p, typ, scale)
            p["xyz"][j, a] += h
            g.append((cp - cm) / (2 * h))
    assert np.abs(g).max() < 1e-3 * max(1.0, c1)


def test_robust_loss_resists_outliers():
    errs = {}
    for typ in (0, 1):
        p = make_ba_problem(n_img=10, n_pts=200, track_len=5, seed=2, noise_px=0.5)
        truth = p["xyz"].copy()
        rng = np.random.default_rng(1)
        idx = rng.choice(len(p["obs_xy"]), 60, replace=False)
        p["obs_xy"][idx] += rng.normal(0, 60, (60, 2))
        clean = np.setdiff1d(np.arange(len(p["obs_xy"])), idx)
        orc.ba_solve(p, max_num_iterations=100, loss_type=typ, loss_scale=1.0)
        R = np.stack([R_from_quat(q / np.linalg.norm(q)) for q in p["qvec"]])
        i = p["obs_img"][clean]
        Xc = np.einsum("nij,nj->ni", R[i], p["xyz"][p["obs_pt"][clean]]) + p["tvec"][i]
        u, v = Xc[:, 0] / Xc[:, 2], Xc[:, 1] / Xc[:, 2]
        k = p["cam_params"][p["img_cam"][i]]
        d = 1 + k[:, 3] * (u * u + v * v)
        e = np.hypot(k[:, 0] * u * d + k[:, 1] - p["obs_xy"][clean, 0], k[:, 0] * v * d + k[:, 2] - p["obs_xy"][clean, 1])
        errs[typ] = np.sqrt((e ** 2).mean())
    assert errs[1] < 0.5 * errs[0]        # inlier reprojection RMS: SOFT_L1 ignores the outliers, L2 does not


def test_kernel_loss_function_equals_oracle_corrector():
    """dagsfm_b200/csrc/ba_loss.cuh (what jacobian_kernel<LOSS> applies) compiled for the host, against the
    oracle's LossEvaluate + Corrector: same rho(s), same scaling sqrt(rho'), and the corrector never leaves
    its rho'' <= 0 branch for these two losses (alpha = 0), which is what the kernel relies on."""
    import ctypes
    import os
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    out = os.path.join(here, "cpp", "_host_ba_loss.so")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.run([cxx, "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", out,
                    os.path.join(here, "cpp", "host_ba_loss.cc")], check=True)
    dev = ctypes.CDLL(out)
    dp = ctypes.POINTER(ctypes.c_double)
    dev.host_ba_loss.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, dp, dp]
    dev.host_ba_loss.restype = None
    f = orc.lib().orc_ba_loss
    f.argtypes = [ctypes.c_int, ctypes.c_double, ctypes.c_double, dp, dp, dp, dp]
    f.restype = None
    rng = np.random.default_rng(0)
    for typ in (1, 2):
        for a in (0.5, 1.0, 2.5):
            for s in np.r_[0.0, 1e-300, 10.0 ** rng.uniform(-12, 12, 300)]:
                rho = (ctypes.c_double * 3)()
                rs, sq, al = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
                f(typ, a, float(s), rho, ctypes.byref(rs), ctypes.byref(sq), ctypes.byref(al))
                r0, w = ctypes.c_double(), ctypes.c_double()
                dev.host_ba_loss(typ, a, float(s), ctypes.byref(r0), ctypes.byref(w))
                assert al.value == 0.0 and rs.value == sq.value          # pure scaling
                assert r0.value == rho[0] and w.value == sq.value        # bit-identical
