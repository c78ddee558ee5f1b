"""ctypes access to the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py -- never by dagsfm_b200/.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    srcs = list(HERE.glob("*.cc")) + [HERE / "Makefile"]
    if force or not LIB_PATH.exists() or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in srcs):
        subprocess.run(["make", "-C", str(HERE), "-s"], check=True)
    return LIB_PATH


def _usable_cores() -> int:
    """Affinity mask capped by the cgroup CPU quota (a box may show 128 CPUs and grant 16)."""
    import os
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = Path("/sys/fs/cgroup/cpu.max").read_text().split()
        if q != "max":
            n = max(1, min(n, int(float(q) / float(per) + 0.5)))
    except Exception:
        pass
    return n


def _cap_openmp() -> None:
    # libgomp would otherwise start one spinning thread per visible CPU; under a cgroup quota
    # that turns every parallel region of ba_oracle.cc into a throttled busy-wait.
    import os
    os.environ.setdefault("OMP_NUM_THREADS", str(_usable_cores()))
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        _cap_openmp()
        L = C.CDLL(str(LIB_PATH))
        vp = C.c_void_p
        L.orc_match_sift.argtypes = [vp, C.c_int, vp, C.c_int, C.c_float, C.c_float, C.c_int, vp, C.c_int]
        L.orc_match_guided.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp, C.c_double, C.c_float,
                                       C.c_float, C.c_int, vp, C.c_int]
        L.orc_best_one_way.argtypes = [vp, C.c_int, vp, C.c_int, C.c_float, C.c_float, vp]
        L.orc_best_one_way.restype = None
        L.orc_create_random_descriptors.argtypes = [C.c_int, C.c_uint, vp]
        L.orc_create_random_descriptors.restype = None
        L.orc_l2_normalize_to_u8.argtypes = [vp, vp]
        L.orc_l2_normalize_to_u8.restype = None
        L.orc_match_pairs_mt.argtypes = [vp, vp, vp, C.c_long, C.c_float, C.c_float, C.c_int, C.c_int, vp, vp]
        L.orc_match_pairs_mt.restype = C.c_double
        _lib = L
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    return a.reshape(-1, 128) if a.size else a.reshape(0, 128)


def match_sift(d1, d2, max_ratio=0.8, max_distance=0.7, cross_check=True) -> np.ndarray:
    """MatchSiftFeaturesCPU -> uint32 [n,2]."""
    d1, d2 = _d(d1), _d(d2)
    cap = max(1, min(d1.shape[0], d2.shape[0]) if cross_check else d1.shape[0])
    out = np.empty((cap, 2), dtype=np.uint32)
    n = lib().orc_match_sift(d1.ctypes.data, d1.shape[0], d2.ctypes.data, d2.shape[0],
                             np.float32(max_ratio), np.float32(max_distance), int(cross_check),
                             out.ctypes.data, cap)
    assert n >= 0
    return out[:n].copy()


def match_guided(kp1, kp2, d1, d2, config, F=None, H=None, max_error=4.0, max_ratio=0.8, max_distance=0.7,
                 cross_check=True):
    """MatchGuidedSiftFeaturesCPU (sift.cc:824-875) -> uint32 [n,2], or None when `config` has no
    guided filter (the reference then leaves inlier_matches untouched)."""
    d1, d2 = _d(d1), _d(d2)
    k1 = np.ascontiguousarray(kp1, dtype=np.float32).reshape(-1, 2)
    k2 = np.ascontiguousarray(kp2, dtype=np.float32).reshape(-1, 2)
    assert len(k1) == len(d1) and len(k2) == len(d2)
    Fm = np.ascontiguousarray(np.eye(3) if F is None else F, dtype=np.float64).reshape(9)
    Hm = np.ascontiguousarray(np.eye(3) if H is None else H, dtype=np.float64).reshape(9)
    cap = max(1, min(d1.shape[0], d2.shape[0]) if cross_check else d1.shape[0])
    out = np.empty((cap, 2), dtype=np.uint32)
    n = lib().orc_match_guided(k1.ctypes.data, k2.ctypes.data, d1.ctypes.data, d1.shape[0], d2.ctypes.data,
                               d2.shape[0], int(config), Fm.ctypes.data, Hm.ctypes.data, float(max_error),
                               np.float32(max_ratio), np.float32(max_distance), int(cross_check), out.ctypes.data, cap)
    if n == -1:
        return None
    assert n >= 0
    return out[:n].copy()


def best_one_way(d1, d2, max_ratio=0.8, max_distance=0.7) -> np.ndarray:
    d1, d2 = _d(d1), _d(d2)
    out = np.full(max(d1.shape[0], 1), -1, dtype=np.int32)
    lib().orc_best_one_way(d1.ctypes.data, d1.shape[0], d2.ctypes.data, d2.shape[0],
                           np.float32(max_ratio), np.float32(max_distance), out.ctypes.data)
    return out[: d1.shape[0]]


def create_random_descriptors(n: int, seed: int = 0) -> np.ndarray:
    """sift_test.cc:243-253 fixture."""
    out = np.zeros((n, 128), dtype=np.uint8)
    if n:
        lib().orc_create_random_descriptors(n, seed, out.ctypes.data)
    return out


def l2_normalize_to_u8(row) -> np.ndarray:
    r = np.ascontiguousarray(row, dtype=np.float32).reshape(128)
    out = np.zeros(128, dtype=np.uint8)
    lib().orc_l2_normalize_to_u8(r.ctypes.data, out.ctypes.data)
    return out


def match_pairs_mt(descs: list, pairs, max_ratio=0.8, max_distance=0.7, cross_check=True, n_threads=8):
    """Multi-threaded CPU baseline; returns (seconds, counts, checksum)."""
    ds = [_d(d) for d in descs]
    ptrs = (C.c_void_p * len(ds))(*[d.ctypes.data for d in ds])
    nd = np.array([d.shape[0] for d in ds], dtype=np.int32)
    pr = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
    counts = np.zeros(len(pr), dtype=np.int32)
    cs = C.c_uint32(0)
    t = lib().orc_match_pairs_mt(ptrs, nd.ctypes.data, pr.ctypes.data, len(pr), np.float32(max_ratio),
                                 np.float32(max_distance), int(cross_check), n_threads,
                                 counts.ctypes.data, C.byref(cs))
    return t, counts, cs.value


# ======================================================================== two-view
class OrcCamera(C.Structure):
    _fields_ = [("model", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("has_prior_focal", C.c_int32), ("params", C.c_double * 12)]


class OrcTvOptions(C.Structure):
    """TwoViewGeometry::Options + RANSACOptions with the matcher's defaults
    (src/feature/sift.h:142-158, src/estimators/two_view_geometry.h:80-117)."""
    _fields_ = [("min_num_inliers", C.c_int32), ("detect_watermark", C.c_int32),
                ("min_E_F_inlier_ratio", C.c_double), ("max_H_inlier_ratio", C.c_double),
                ("watermark_min_inlier_ratio", C.c_double), ("watermark_border_size", C.c_double),
                ("max_error", C.c_double), ("min_inlier_ratio", C.c_double), ("confidence", C.c_double),
                ("min_num_trials", C.c_int64), ("max_num_trials", C.c_int64)]


class OrcTvResult(C.Structure):
    _fields_ = [("config", C.c_int32), ("n_inliers", C.c_int32),
                ("E_inl", C.c_int32), ("F_inl", C.c_int32), ("H_inl", C.c_int32),
                ("E_trials", C.c_int32), ("F_trials", C.c_int32), ("H_trials", C.c_int32),
                ("E", C.c_double * 9), ("F", C.c_double * 9), ("H", C.c_double * 9)]


def tv_default_options() -> OrcTvOptions:
    return OrcTvOptions(15, 1, 0.95, 0.8, 0.7, 0.1, 4.0, 0.25, 0.999, 30, 10000)


def make_camera(model=2, width=1000, height=1000, params=(1200.0, 500.0, 500.0, 0.0), prior=True) -> OrcCamera:
    c = OrcCamera()
    c.model, c.width, c.height, c.has_prior_focal = model, width, height, int(prior)
    for i, v in enumerate(params):
        c.params[i] = v
    return c


_tv_ready = False


def _tv():
    global _tv_ready
    L = lib()
    if not _tv_ready:
        vp, d, i32, i64 = C.c_void_p, C.c_double, C.c_int, C.c_int64
        L.orc_f7.argtypes = [vp, vp, vp]
        L.orc_eight_point.argtypes = [i32, vp, vp, i32, vp]
        L.orc_e5.argtypes = [i32, vp, vp, vp, vp, vp]
        L.orc_h_dlt.argtypes = [i32, vp, vp, vp]
        L.orc_e5_system.argtypes = [vp, vp]; L.orc_e5_system.restype = None
        L.orc_e5_det_coeffs.argtypes = [vp, vp]; L.orc_e5_det_coeffs.restype = None
        L.orc_residuals.argtypes = [i32, i32, vp, vp, vp, vp]; L.orc_residuals.restype = None
        L.orc_compute_num_trials.argtypes = [C.c_uint64, C.c_uint64, d, i32]
        L.orc_compute_num_trials.restype = C.c_uint64
        L.orc_poly_roots.argtypes = [i32, vp, vp, vp]
        L.orc_svd.argtypes = [vp, i32, i32, vp, vp]; L.orc_svd.restype = None
        L.orc_sample_stream.argtypes = [C.c_uint, i32, i32, i32, vp]; L.orc_sample_stream.restype = None
        L.orc_image_to_world.argtypes = [C.POINTER(OrcCamera), i32, vp, vp]; L.orc_image_to_world.restype = None
        L.orc_ransac.argtypes = [i32, i32, i32, vp, vp, d, d, d, i64, i64, C.c_uint, vp, vp, vp, vp, vp]
        L.orc_two_view.argtypes = [C.POINTER(OrcCamera), vp, C.POINTER(OrcCamera), vp, vp, i32,
                                   C.POINTER(OrcTvOptions), C.c_uint, C.POINTER(OrcTvResult), vp]
        L.orc_two_view.restype = None
        L.orc_two_view_pairs_mt.argtypes = [vp, vp, vp, C.c_long, vp, vp, C.POINTER(OrcTvOptions), vp, i32, vp]
        L.orc_two_view_pairs_mt.restype = d
        L.orc_two_view_pairs_mt2.argtypes = [vp, vp, vp, C.c_long, vp, vp, C.POINTER(OrcTvOptions), vp, i32, vp, vp]
        L.orc_two_view_pairs_mt2.restype = d
        L.orc_set_solver_stack.argtypes = [i32]; L.orc_set_solver_stack.restype = None
        L.orc_get_solver_stack.argtypes = []; L.orc_get_solver_stack.restype = i32
        _tv_ready = True
    return L


def _p(a):
    return np.ascontiguousarray(a, dtype=np.float64)


EST_E5, EST_F7, EST_H4, EST_T2 = 0, 1, 2, 3


def f7(p1, p2):
    p1, p2 = _p(p1), _p(p2)
    out = np.zeros((3, 3, 3))
    n = _tv().orc_f7(p1.ctypes.data, p2.ctypes.data, out.ctypes.data)
    return out[:n]


def eight_point(p1, p2, essential=False):
    p1, p2 = _p(p1), _p(p2)
    out = np.zeros((3, 3))
    _tv().orc_eight_point(len(p1), p1.ctypes.data, p2.ctypes.data, int(essential), out.ctypes.data)
    return out


def e5(p1, p2, with_system=False):
    p1, p2 = _p(p1), _p(p2)
    out = np.zeros((10, 3, 3)); A = np.zeros((10, 20)); c = np.zeros(11)
    if with_system:   # the 10 x 20 system / determinant coefficients exist in the oracle's own stack only
        n = _tv().orc_e5(len(p1), p1.ctypes.data, p2.ctypes.data, out.ctypes.data, A.ctypes.data, c.ctypes.data)
        return out[:n], A, c
    n = _tv().orc_e5(len(p1), p1.ctypes.data, p2.ctypes.data, out.ctypes.data, None, None)
    return out[:n]


def h_dlt(p1, p2):
    p1, p2 = _p(p1), _p(p2)
    out = np.zeros((3, 3))
    _tv().orc_h_dlt(len(p1), p1.ctypes.data, p2.ctypes.data, out.ctypes.data)
    return out


def e5_system(basis_4x9):
    b = _p(basis_4x9); A = np.zeros((10, 20))
    _tv().orc_e5_system(b.ctypes.data, A.ctypes.data)
    return A


def e5_det_coeffs(B_colmajor39):
    b = _p(B_colmajor39); c = np.zeros(11)
    _tv().orc_e5_det_coeffs(b.ctypes.data, c.ctypes.data)
    return c


def residuals(est_type, p1, p2, M):
    p1, p2, M = _p(p1), _p(p2), _p(M)
    r = np.zeros(len(p1))
    _tv().orc_residuals(est_type, len(p1), p1.ctypes.data, p2.ctypes.data, M.ctypes.data, r.ctypes.data)
    return r


def compute_num_trials(num_inliers, num_samples, confidence, kmin):
    return int(_tv().orc_compute_num_trials(num_inliers, num_samples, confidence, kmin))


def poly_roots(coeffs):
    c = _p(coeffs); re = np.zeros(len(c)); im = np.zeros(len(c))
    n = _tv().orc_poly_roots(len(c), c.ctypes.data, re.ctypes.data, im.ctypes.data)
    return (re[:n], im[:n]) if n >= 0 else None


def svd(A):
    A = _p(A); m, n = A.shape
    s = np.zeros(n); V = np.zeros((n, n))
    _tv().orc_svd(A.ctypes.data, m, n, s.ctypes.data, V.ctypes.data)
    return s, V


def center_and_normalize(pts):
    p = _p(pts); out = np.zeros_like(p); T = np.zeros((3, 3))
    f = _tv().orc_center_and_normalize
    f.argtypes = [C.c_int] + [C.c_void_p] * 3
    f.restype = None
    f(len(p), p.ctypes.data, out.ctypes.data, T.ctypes.data)
    return out, T


def support_evaluate(residuals, max_residual):
    r = _p(residuals); n = C.c_long(0); s = C.c_double(0)
    f = _tv().orc_support_evaluate
    f.argtypes = [C.c_void_p, C.c_int, C.c_double, C.POINTER(C.c_long), C.POINTER(C.c_double)]
    f.restype = None
    f(r.ctypes.data, len(r), float(max_residual), C.byref(n), C.byref(s))
    return n.value, s.value


def support_compare(n1, s1, n2, s2) -> bool:
    f = _tv().orc_support_compare
    f.argtypes = [C.c_long, C.c_double, C.c_long, C.c_double]
    return bool(f(int(n1), float(s1), int(n2), float(s2)))


def translation_estimate(src, dst):
    a, b = _p(src), _p(dst); t = np.zeros(2); r = np.zeros(len(a))
    f = _tv().orc_translation_estimate
    f.argtypes = [C.c_int] + [C.c_void_p] * 4
    f.restype = None
    f(len(a), a.ctypes.data, b.ctypes.data, t.ctypes.data, r.ctypes.data)
    return t, r


def sample_stream(seed, total, k, n_trials):
    out = np.zeros((n_trials, k), dtype=np.int32)
    _tv().orc_sample_stream(seed, total, k, n_trials, out.ctypes.data)
    return out


def image_to_world(cam: OrcCamera, xy):
    xy = _p(xy); out = np.zeros_like(xy)
    _tv().orc_image_to_world(C.byref(cam), len(xy), xy.ctypes.data, out.ctypes.data)
    return out


def world_to_image(cam: OrcCamera, uv):
    uv = _p(uv); out = np.zeros_like(uv)
    _tv().orc_world_to_image(C.byref(cam), len(uv), C.c_void_p(uv.ctypes.data), C.c_void_p(out.ctypes.data))
    return out


def image_to_world_threshold(cam: OrcCamera, thr: float) -> float:
    f = _tv().orc_image_to_world_threshold
    f.restype = C.c_double
    f.argtypes = [C.POINTER(OrcCamera), C.c_double]
    return f(C.byref(cam), float(thr))


def camera_num_params(model: int) -> int:
    return _tv().orc_camera_num_params(int(model))


def ransac(est_type, X, Y, max_error, min_inlier_ratio=0.1, confidence=0.99, min_num_trials=0,
           max_num_trials=2**62, seed=0, use_lo=True):
    X, Y = _p(X), _p(Y)
    n = len(X)
    model = np.zeros((3, 3)); ni = C.c_int32(0); rs = C.c_double(0); nt = C.c_int64(0)
    mask = np.zeros(max(n, 1), dtype=np.uint8)
    ok = _tv().orc_ransac(est_type, int(use_lo), n, X.ctypes.data, Y.ctypes.data, max_error, min_inlier_ratio,
                          confidence, min_num_trials, max_num_trials, seed, model.ctypes.data,
                          C.byref(ni), C.byref(rs), C.byref(nt), mask.ctypes.data)
    return {"success": bool(ok), "model": model, "num_inliers": ni.value, "residual_sum": rs.value,
            "num_trials": nt.value, "mask": mask[:n].astype(bool)}


def set_solver_stack(stack: int) -> None:
    """0 = the oracle's own floating-point solver stack (sequential sums; the default, pinned to the reference's
    goldens); 1 = the same algorithms evaluated in the CUDA path's operation order (host build of
    dagsfm_b200/csrc/verify_solvers.cuh + the warp-order QR / Jacobi restated in twoview_oracle.cc), against which the
    GPU parity tests assert 100 % identity.  Process-wide."""
    _tv().orc_set_solver_stack(int(stack))


def get_solver_stack() -> int:
    return int(_tv().orc_get_solver_stack())


class solver_stack:
    """with solver_stack(1): ... -- scoped selection, restores the previous stack."""
    def __init__(self, stack):
        self.stack = stack

    def __enter__(self):
        self.prev = get_solver_stack()
        set_solver_stack(self.stack)

    def __exit__(self, *a):
        set_solver_stack(self.prev)


def two_view(cam1, pts1, cam2, pts2, matches, opt=None, seed=0):
    opt = opt or tv_default_options()
    pts1, pts2 = _p(pts1), _p(pts2)
    m = np.ascontiguousarray(matches, dtype=np.uint32).reshape(-1, 2)
    res = OrcTvResult()
    inl = np.zeros((max(len(m), 1), 2), dtype=np.uint32)
    _tv().orc_two_view(C.byref(cam1), pts1.ctypes.data, C.byref(cam2), pts2.ctypes.data, m.ctypes.data, len(m),
                       C.byref(opt), seed, C.byref(res), inl.ctypes.data)
    return res, inl[: res.n_inliers].copy()


def two_view_multiple(cam1, pts1, cam2, pts2, matches, opt=None, seed=0, multiple_ignore_watermark=True):
    """TwoViewGeometry::EstimateMultiple (two_view_geometry.cc:128-167) restated on top of `two_view`:
    estimate, drop the inliers BY VALUE (ExtractOutlierMatches, :67-88), repeat until DEGENERATE.
    Round r is seeded with seed + r * 0x9E3779B9 (the reference continues an unseeded thread-local
    PRNG).  Returns (config, list of per-geometry results, inlier matches uint32 [n,2])."""
    remaining = np.ascontiguousarray(matches, dtype=np.uint32).reshape(-1, 2)
    geos, inliers, rnd = [], [], 0
    while True:
        res, inl = two_view(cam1, pts1, cam2, pts2, remaining, opt, seed=(int(seed) + rnd * 0x9E3779B9) & 0xFFFFFFFF)
        rnd += 1
        if res.config == 1:        # DEGENERATE
            break
        if not (multiple_ignore_watermark and res.config == 7):
            geos.append(res)
            inliers.append(inl)
        gone = {(int(a), int(b)) for a, b in inl}
        remaining = np.array([m for m in remaining if (int(m[0]), int(m[1])) not in gone], dtype=np.uint32).reshape(-1, 2)
    if not geos:
        return 1, geos, np.zeros((0, 2), np.uint32)
    if len(geos) == 1:
        return geos[0].config, geos, inliers[0]
    return 8, geos, np.concatenate(inliers)


# ============================================================== bundle adjustment
class OrcBaProblem(C.Structure):
    _fields_ = [("n_img", C.c_int32), ("n_cam", C.c_int32), ("n_pts", C.c_int32), ("n_obs", C.c_int64),
                ("qvec", C.c_void_p), ("tvec", C.c_void_p), ("img_cam", C.c_void_p), ("pose_const", C.c_void_p),
                ("tvec_const", C.c_void_p), ("cam_model", C.c_void_p), ("cam_params", C.c_void_p),
                ("cam_const", C.c_void_p), ("refine_focal", C.c_int32), ("refine_principal", C.c_int32),
                ("refine_extra", C.c_int32), ("xyz", C.c_void_p), ("pt_const", C.c_void_p),
                ("obs_img", C.c_void_p), ("obs_pt", C.c_void_p), ("obs_xy", C.c_void_p), ("cam_stride", C.c_int32)]


class OrcBaOptions(C.Structure):
    _fields_ = [("max_num_iterations", C.c_int32), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double), ("n_threads", C.c_int32),
                ("loss_type", C.c_int32), ("loss_scale", C.c_double),
                ("linear_solver", C.c_int32), ("max_linear_solver_iterations", C.c_int32)]


class OrcBaSummary(C.Structure):
    _fields_ = [("initial_cost", C.c_double), ("final_cost", C.c_double), ("num_successful_steps", C.c_int32),
                ("num_unsuccessful_steps", C.c_int32), ("termination", C.c_int32), ("num_residuals", C.c_int32),
                ("num_effective_parameters", C.c_int32), ("seconds", C.c_double),
                ("num_linear_iterations", C.c_int64)]


def ba_solve(prob: dict, max_num_iterations=50, function_tolerance=0.0, gradient_tolerance=1.0,
             parameter_tolerance=0.0, loss_type=0, loss_scale=1.0, linear_solver=0, max_linear_solver_iterations=100):
    """prob: dict of numpy arrays (see tests/ba_scene.py); qvec/tvec/cam_params/xyz are updated in place.
    linear_solver: 0 exact Schur step (DENSE/SPARSE_SCHUR), 1 ITERATIVE_SCHUR + SCHUR_JACOBI."""
    L = lib()
    L.orc_ba_solve.argtypes = [C.POINTER(OrcBaProblem), C.POINTER(OrcBaOptions), C.POINTER(OrcBaSummary)]
    L.orc_ba_solve.restype = None
    p = OrcBaProblem()
    p.n_img, p.n_cam, p.n_pts, p.n_obs = len(prob["qvec"]), len(prob["cam_params"]), len(prob["xyz"]), len(prob["obs_img"])
    for k in ("qvec", "tvec", "img_cam", "pose_const", "tvec_const", "cam_model", "cam_params", "cam_const", "xyz",
              "pt_const", "obs_img", "obs_pt", "obs_xy"):
        assert prob[k].flags["C_CONTIGUOUS"]
        setattr(p, k, prob[k].ctypes.data)
    p.refine_focal, p.refine_principal, p.refine_extra = prob.get("refine", (1, 0, 1))
    p.cam_stride = prob["cam_params"].shape[1]     # 4 (SIMPLE_PINHOLE / PINHOLE / SIMPLE_RADIAL) or up to 12
    o = OrcBaOptions(max_num_iterations, function_tolerance, gradient_tolerance, parameter_tolerance, 0,
                     int(loss_type), float(loss_scale), int(linear_solver), int(max_linear_solver_iterations))
    s = OrcBaSummary()
    L.orc_ba_solve(C.byref(p), C.byref(o), C.byref(s))
    return s


def ba_mean_reprojection_error(prob: dict):
    """Reconstruction::ComputeMeanReprojectionError over the problem's tracks -> (mean error, per-point errors)."""
    L = lib()
    L.orc_ba_mean_reprojection_error.argtypes = [C.POINTER(OrcBaProblem), C.c_void_p]
    L.orc_ba_mean_reprojection_error.restype = C.c_double
    p = OrcBaProblem()
    p.n_img, p.n_cam, p.n_pts, p.n_obs = len(prob["qvec"]), len(prob["cam_params"]), len(prob["xyz"]), len(prob["obs_img"])
    for k in ("qvec", "tvec", "img_cam", "pose_const", "tvec_const", "cam_model", "cam_params", "cam_const", "xyz",
              "pt_const", "obs_img", "obs_pt", "obs_xy"):
        assert prob[k].flags["C_CONTIGUOUS"]
        setattr(p, k, prob[k].ctypes.data)
    p.cam_stride = prob["cam_params"].shape[1]
    err = np.zeros(max(p.n_pts, 1))
    mean = L.orc_ba_mean_reprojection_error(C.byref(p), err.ctypes.data)
    return mean, err[: p.n_pts]


def ba_evaluate(model, q, t, X, k, obs):
    L = lib()
    L.orc_ba_evaluate.argtypes = [C.c_int] + [C.c_void_p] * 10
    L.orc_ba_evaluate.restype = None
    q, t, X, obs = (_p(a) for a in (q, t, X, obs))
    k = np.concatenate([_p(k), np.zeros(12)])[:12].copy()
    r = np.zeros(2); Jq = np.zeros((2, 3)); Jt = np.zeros((2, 3)); JX = np.zeros((2, 3)); Jk = np.zeros((2, 12))
    L.orc_ba_evaluate(model, q.ctypes.data, t.ctypes.data, X.ctypes.data, k.ctypes.data, obs.ctypes.data,
                      r.ctypes.data, Jq.ctypes.data, Jt.ctypes.data, JX.ctypes.data, Jk.ctypes.data)
    return r, Jq, Jt, JX, Jk[:, : (4 if model <= 2 else 12)].copy()   # 2 x 4 for the three 4-parameter-slot models


def ba_quat_plus(x, d):
    L = lib()
    L.orc_ba_quat_plus.argtypes = [C.c_void_p] * 3
    L.orc_ba_quat_plus.restype = None
    x, d = _p(x), _p(d)
    out = np.zeros(4)
    L.orc_ba_quat_plus(x.ctypes.data, d.ctypes.data, out.ctypes.data)
    return out



# ------------------------------------------------- relative pose
class OrcRelPose(C.Structure):
    _fields_ = [("qvec", C.c_double * 4), ("tvec", C.c_double * 3), ("tri_angle", C.c_double),
                ("config", C.c_int32), ("n_points3D", C.c_int32)]


def decompose_essential(E):
    E = _p(E); R1 = np.zeros((3, 3)); R2 = np.zeros((3, 3)); t = np.zeros(3)
    _tv().orc_decompose_essential(C.c_void_p(E.ctypes.data), C.c_void_p(R1.ctypes.data), C.c_void_p(R2.ctypes.data), C.c_void_p(t.ctypes.data))
    return R1, R2, t


def pose_from_essential(E, p1, p2):
    E, p1, p2 = _p(E), _p(p1), _p(p2)
    R = np.zeros((3, 3)); t = np.zeros(3); X = np.zeros((max(len(p1), 1), 3))
    f = _tv().orc_pose_from_essential
    f.restype = C.c_int
    n = f(C.c_void_p(E.ctypes.data), len(p1), C.c_void_p(p1.ctypes.data), C.c_void_p(p2.ctypes.data), C.c_void_p(R.ctypes.data),
          C.c_void_p(t.ctypes.data), C.c_void_p(X.ctypes.data))
    return R, t, X[:n]


def decompose_homography(H, K1, K2):
    H, K1, K2 = _p(H), _p(K1), _p(K2)
    R = np.zeros((4, 3, 3)); t = np.zeros((4, 3)); nrm = np.zeros((4, 3))
    f = _tv().orc_decompose_homography
    f.restype = C.c_int
    n = f(C.c_void_p(H.ctypes.data), C.c_void_p(K1.ctypes.data), C.c_void_p(K2.ctypes.data), C.c_void_p(R.ctypes.data),
          C.c_void_p(t.ctypes.data), C.c_void_p(nrm.ctypes.data))
    return R[:n], t[:n], nrm[:n]


def pose_from_homography(H, K1, K2, p1, p2):
    H, K1, K2, p1, p2 = _p(H), _p(K1), _p(K2), _p(p1), _p(p2)
    R = np.zeros((3, 3)); t = np.zeros(3); nrm = np.zeros(3); X = np.zeros((max(len(p1), 1), 3))
    f = _tv().orc_pose_from_homography
    f.restype = C.c_int
    n = f(C.c_void_p(H.ctypes.data), C.c_void_p(K1.ctypes.data), C.c_void_p(K2.ctypes.data), len(p1), C.c_void_p(p1.ctypes.data),
          C.c_void_p(p2.ctypes.data), C.c_void_p(R.ctypes.data), C.c_void_p(t.ctypes.data), C.c_void_p(nrm.ctypes.data),
          C.c_void_p(X.ctypes.data))
    return R, t, nrm, X[:n]


def triangulate_point(R, t, p1, p2):
    R, t, p1, p2 = _p(R), _p(t), _p(p1), _p(p2)
    X = np.zeros(3)
    _tv().orc_triangulate_point(C.c_void_p(R.ctypes.data), C.c_void_p(t.ctypes.data), C.c_void_p(p1.ctypes.data),
                                C.c_void_p(p2.ctypes.data), C.c_void_p(X.ctypes.data))
    return X


def check_cheirality(R, t, p1, p2):
    """CheckCheirality (base/pose.cc:225-248): the triangulated points in front of both cameras."""
    R, t, p1, p2 = _p(R), _p(t), _p(p1), _p(p2)
    X = np.zeros((max(len(p1), 1), 3))
    f = _tv().orc_check_cheirality
    f.restype = C.c_int
    n = f(C.c_void_p(R.ctypes.data), C.c_void_p(t.ctypes.data), len(p1), C.c_void_p(p1.ctypes.data), C.c_void_p(p2.ctypes.data),
          C.c_void_p(X.ctypes.data))
    return X[:n]


def triangulation_angles(R, t, points3D):
    R, t, X = _p(R), _p(t), _p(points3D)
    a = np.zeros(len(X))
    _tv().orc_triangulation_angles(C.c_void_p(R.ctypes.data), C.c_void_p(t.ctypes.data), len(X), C.c_void_p(X.ctypes.data),
                                   C.c_void_p(a.ctypes.data))
    return a


def median(v):
    v = _p(v)
    f = _tv().orc_median
    f.restype = C.c_double
    return f(len(v), C.c_void_p(v.ctypes.data))


def rotation_to_quaternion(R):
    R = _p(R); q = np.zeros(4)
    _tv().orc_rotation_to_quaternion(C.c_void_p(R.ctypes.data), C.c_void_p(q.ctypes.data))
    return q


def relative_pose(cam1, pts1, cam2, pts2, config, E, H, inlier_matches) -> OrcRelPose:
    """The part of TwoViewGeometry::EstimateWithRelativePose after EstimateCalibrated (two_view_geometry.cc:239-289)."""
    pts1, pts2, E, H = _p(pts1), _p(pts2), _p(E), _p(H)
    m = np.ascontiguousarray(inlier_matches, dtype=np.uint32).reshape(-1, 2)
    out = OrcRelPose()
    f = _tv().orc_relative_pose
    f.restype = None
    f(C.byref(cam1), C.c_void_p(pts1.ctypes.data), C.byref(cam2), C.c_void_p(pts2.ctypes.data), int(config),
      C.c_void_p(E.ctypes.data), C.c_void_p(H.ctypes.data), C.c_void_p(m.ctypes.data), len(m), C.byref(out))
    return out

# ------------------------------------------------- reference-owned BA (vendored PBA, oracle/_ref)
PBA_REF_PATH = HERE / "_ref" / "libpba_ref.so"


def pba_ref_available() -> bool:
    return PBA_REF_PATH.exists()


def pba_ref_solve(prob: dict, n_threads=8, max_iter=50, lib_path=None):
    """Runs the reference's vendored PBA (CPU double) on a tests/ba_scene problem with unshared
    SIMPLE_RADIAL cameras.  Gauge: images with const pose become constant cameras (PBA cannot
    fix partial extrinsics).  Returns dict(initial_mse, final_mse, lm_iterations, seconds) and
    updates xyz in prob (float32-quantised, as PBA's interface is)."""
    from tests.ba_scene import R_from_quat
    L = C.CDLL(str(lib_path or PBA_REF_PATH))
    L.pba_ref_run.argtypes = [C.c_int] + [C.c_void_p] * 5 + [C.c_int, C.c_void_p, C.c_long, C.c_void_p, C.c_void_p,
                                                             C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5
    n = len(prob["qvec"])
    assert (prob["img_cam"] == np.arange(n)).all() and (prob["cam_model"] == 2).all()
    f = np.ascontiguousarray(prob["cam_params"][:, 0]); k = np.ascontiguousarray(prob["cam_params"][:, 3])
    R = np.ascontiguousarray(np.stack([R_from_quat(q) for q in prob["qvec"]]).reshape(n, 9))
    t = np.ascontiguousarray(prob["tvec"])
    cc = np.ascontiguousarray(prob["pose_const"])
    xy = np.ascontiguousarray(prob["obs_xy"] - prob["cam_params"][prob["obs_img"], 1:3])
    of, ok, oR, ot, st = np.zeros(n), np.zeros(n), np.zeros((n, 9)), np.zeros((n, 3)), np.zeros(4)
    import os, sys
    sys.stdout.flush()
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)                       # PBA prints its banner with std::cout
    try:
        _pba_call(L, n, f, k, R, t, cc, prob, xy, n_threads, max_iter, of, ok, oR, ot, st)
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)
    return {"initial_mse": st[0], "final_mse": st[1], "lm_iterations": int(st[2]), "seconds": st[3],
            "focal": of, "radial": ok, "R": oR.reshape(n, 3, 3), "t": ot}


def _pba_call(L, n, f, k, R, t, cc, prob, xy, n_threads, max_iter, of, ok, oR, ot, st):
    L.pba_ref_run(n, f.ctypes.data, k.ctypes.data, R.ctypes.data, t.ctypes.data, cc.ctypes.data, len(prob["xyz"]),
                  prob["xyz"].ctypes.data, len(prob["obs_img"]), xy.ctypes.data, prob["obs_pt"].ctypes.data,
                  prob["obs_img"].ctypes.data, n_threads, max_iter, of.ctypes.data, ok.ctypes.data, oR.ctypes.data,
                  ot.ctypes.data, st.ctypes.data)


# ======================================================================== retrieval (vocabulary tree)
class RetrievalOracle:
    """oracle/retrieval_oracle.cc: VisualIndex Add / Prepare / Query with exact word search."""

    def __init__(self, words, proj, thresholds, has_embedding):
        L = lib()
        vp = C.c_void_p
        L.orc_retrieval_create.restype = vp
        L.orc_retrieval_create.argtypes = [C.c_int, vp, vp, vp, vp]
        L.orc_retrieval_destroy.argtypes = [vp]
        L.orc_retrieval_word_ids.argtypes = [vp, C.c_int, vp, C.c_int, vp]
        L.orc_retrieval_signatures.argtypes = [vp, C.c_int, vp, vp, vp]
        L.orc_retrieval_add.argtypes = [vp, C.c_int, C.c_int, vp]
        L.orc_retrieval_prepare.argtypes = [vp]
        L.orc_retrieval_query.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, vp, vp]
        L.orc_retrieval_lut.argtypes = [vp, vp]
        self._w = np.ascontiguousarray(words, np.uint8).reshape(-1, 128)
        self._p = np.ascontiguousarray(proj, np.float32).reshape(64, 128)
        self._t = np.ascontiguousarray(thresholds, np.float32).reshape(len(self._w), 64)
        self._h = np.ascontiguousarray(has_embedding, np.uint8)
        self._x = vp(L.orc_retrieval_create(len(self._w), self._w.ctypes.data, self._p.ctypes.data, self._t.ctypes.data, self._h.ctypes.data))
        self._n_images = 0

    def __del__(self):
        try:
            lib().orc_retrieval_destroy(self._x)
        except Exception:
            pass

    def word_ids(self, desc, k):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 128)
        out = np.zeros((len(d), k), np.int32)
        lib().orc_retrieval_word_ids(self._x, len(d), d.ctypes.data, k, out.ctypes.data)
        return out

    def signatures(self, desc, word):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 128)
        w = np.ascontiguousarray(word, np.int32)
        out = np.zeros(len(d), np.uint64)
        lib().orc_retrieval_signatures(self._x, len(d), d.ctypes.data, w.ctypes.data, out.ctypes.data)
        return out

    def lut(self):
        out = np.zeros(65, np.float32)
        lib().orc_retrieval_lut(self._x, out.ctypes.data)
        return out

    def Add(self, image_id, desc):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 128)
        lib().orc_retrieval_add(self._x, int(image_id), len(d), d.ctypes.data)
        self._n_images += 1

    def Prepare(self):
        lib().orc_retrieval_prepare(self._x)

    def Query(self, desc, num_neighbors=5, max_num_images=-1):
        d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 128)
        ids = np.zeros(max(self._n_images, 1), np.int32)
        sc = np.zeros(max(self._n_images, 1), np.float32)
        m = lib().orc_retrieval_query(self._x, len(d), d.ctypes.data, num_neighbors, max_num_images, ids.ctypes.data, sc.ctypes.data)
        return ids[:m].copy(), sc[:m].copy()


# ------------------------------------------------------------------------ the reference's vendored FLANN (oracle/_ref)
FLANN_REF_PATH = Path(__file__).resolve().parent / "_ref" / "libflann_ref.so"


def flann_ref_available() -> bool:
    return FLANN_REF_PATH.exists()


def _flann():
    L = C.CDLL(str(FLANN_REF_PATH))
    L.flann_ref_knn_linear.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.flann_ref_knn_autotuned.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p]
    return L


def flann_ref_knn_linear(words, desc, k):
    """The k nearest words by the reference's own FLANN in EXACT mode (flann::LinearIndex over flann::L2<uint8>).
    -> (ids int32 [n, k], squared distances float32 [n, k])"""
    w = np.ascontiguousarray(words, np.uint8).reshape(-1, 128)
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 128)
    ids = np.zeros((len(d), k), np.int32)
    dist = np.zeros((len(d), k), np.float32)
    _flann().flann_ref_knn_linear(w.ctypes.data, len(w), d.ctypes.data, len(d), k, ids.ctypes.data, dist.ctypes.data)
    return ids, dist


def flann_ref_knn_autotuned(words, desc, k, num_checks=256, target_precision=0.95, cores=1):
    """What VisualIndex::Build + FindWordIds do (visual_index.h:517-521, 701-744; BuildOptions::target_precision 0.95,
    num_checks 256): approximate, tuned by timing experiments on this host, randomised trees.  -> ids int32 [n, k]"""
    w = np.ascontiguousarray(words, np.uint8).reshape(-1, 128)
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 128)
    ids = np.zeros((len(d), k), np.int32)
    _flann().flann_ref_knn_autotuned(w.ctypes.data, len(w), d.ctypes.data, len(d), k, num_checks, C.c_float(target_precision), cores, ids.ctypes.data)
    return ids


def reference_similarity_ransac_test(use_lo: bool):
    """ransac_test.cc:89-133 / loransac_test.cc:57-107 replayed literally on the oracle's RANSAC loop (same data recipe and PRNG
    stream: std::mt19937(0), RandomReal outliers, then the sampler), with the reference's 3-D similarity estimator restated.
    -> dict(success, num_trials, num_inliers, mask bool [1000], matrix_diff)"""
    L = _tv()
    nt, ni, md = C.c_int64(0), C.c_int32(0), C.c_double(0)
    mask = np.zeros(1000, np.uint8)
    ok = L.orc_reference_similarity_ransac_test(int(use_lo), C.byref(nt), C.byref(ni), mask.ctypes.data_as(C.c_void_p), C.byref(md))
    return {"success": bool(ok), "num_trials": nt.value, "num_inliers": ni.value, "mask": mask.astype(bool), "matrix_diff": md.value}
