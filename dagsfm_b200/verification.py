"""Host-side mirror of the reference's two-view verification interface.

Reference names kept: TwoViewGeometry (src/estimators/two_view_geometry.h:52-306) with
its Options / ConfigurationType, Camera (src/base/camera.h), and the batched seam that
replaces the TwoViewGeometryVerifier threads (src/feature/matching.cc:571-608).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from ._lib import check, lib


class Camera(C.Structure):
    """b2_camera.  model: the ids of camera_models.h:117-129 (0 SIMPLE_PINHOLE ... 10 THIN_PRISM_FISHEYE)."""
    _fields_ = [("model", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("has_prior_focal_length", C.c_int32), ("params", C.c_double * 12)]

    @staticmethod
    def make(model=2, width=1000, height=1000, params=(1200.0, 500.0, 500.0, 0.0), prior_focal=True):
        c = Camera()
        c.model, c.width, c.height, c.has_prior_focal_length = model, width, height, int(prior_focal)
        for i, v in enumerate(params):
            c.params[i] = v
        return c


class TwoViewOptions(C.Structure):
    """b2_two_view_options == TwoViewGeometry::Options + RANSACOptions (matcher defaults)."""
    _fields_ = [("min_num_inliers", C.c_int32), ("detect_watermark", C.c_int32),
                ("min_E_F_inlier_ratio", C.c_double), ("max_H_inlier_ratio", C.c_double),
                ("watermark_min_inlier_ratio", C.c_double), ("watermark_border_size", C.c_double),
                ("max_error", C.c_double), ("min_inlier_ratio", C.c_double), ("confidence", C.c_double),
                ("min_num_trials", C.c_int64), ("max_num_trials", C.c_int64)]

    @staticmethod
    def default():
        o = TwoViewOptions()
        lib().b2_two_view_default_options(C.byref(o))
        return o


class TwoViewResult(C.Structure):
    _fields_ = [("config", C.c_int32), ("n_inliers", C.c_int32),
                ("E_num_inliers", C.c_int32), ("F_num_inliers", C.c_int32), ("H_num_inliers", C.c_int32),
                ("E_num_trials", C.c_int32), ("F_num_trials", C.c_int32), ("H_num_trials", C.c_int32),
                ("E", C.c_double * 9), ("F", C.c_double * 9), ("H", C.c_double * 9)]


RESULT_DTYPE = np.dtype([("config", "<i4"), ("n_inliers", "<i4"), ("E_num_inliers", "<i4"),
                         ("F_num_inliers", "<i4"), ("H_num_inliers", "<i4"), ("E_num_trials", "<i4"),
                         ("F_num_trials", "<i4"), ("H_num_trials", "<i4"),
                         ("E", "<f8", (9,)), ("F", "<f8", (9,)), ("H", "<f8", (9,))])
assert RESULT_DTYPE.itemsize == C.sizeof(TwoViewResult)

POSE_DTYPE = np.dtype([("qvec", "<f8", (4,)), ("tvec", "<f8", (3,)), ("tri_angle", "<f8"), ("config", "<i4"),
                       ("n_points3D", "<i4")])   # b2_relative_pose

# TwoViewGeometry::ConfigurationType
UNDEFINED, DEGENERATE, CALIBRATED, UNCALIBRATED, PLANAR, PANORAMIC, PLANAR_OR_PANORAMIC, WATERMARK, MULTIPLE = range(9)

_bound = False


def _L():
    global _bound
    L = lib()
    if not _bound:
        vp, i32, i64, P = C.c_void_p, C.c_int32, C.c_int64, C.POINTER
        L.b2_two_view_default_options.argtypes = [P(TwoViewOptions)]
        L.b2_two_view_default_options.restype = None
        L.b2_verify_create.argtypes = [C.c_int, P(vp)]
        L.b2_verify_destroy.argtypes = [vp]
        L.b2_verify_set_images.argtypes = [vp, i32, vp, P(vp), P(i32)]
        L.b2_verify_pairs.argtypes = [vp, i64, vp, vp, vp, P(TwoViewOptions), vp, vp, vp]
        L.b2_verify_pairs_device.argtypes = [vp, i64, vp, vp, vp, P(TwoViewOptions), vp, vp, vp]
        L.b2_verify_pairs_multiple.argtypes = [vp, i64, vp, vp, vp, P(TwoViewOptions), i32, vp, vp, vp]
        L.b2_verify_relative_pose.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        L.b2_verify_relative_pose_device.argtypes = [vp, i64, vp, vp, vp, vp, vp]
        L.b2_score_models.argtypes = [vp, i32, i32, vp, vp, i32, vp, C.c_double, vp, vp, vp]
        L.b2_verify_debug_sample_stream.argtypes = [vp, C.c_uint32, i32, i32, i32, vp]
        L.b2_verify_debug_solve.argtypes = [vp, i32, i32, vp, vp, vp, P(i32)]
        L.b2_verify_last_timing.argtypes = [vp, P(C.c_double)]
        L.b2_verify_debug_normalized.argtypes = [vp, i32, vp]
        _bound = True
    return L


class TwoViewGeometryVerifier:
    """One verifier per GPU (the reference runs num_threads CPU verifier threads)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(_L().b2_verify_create(device, C.byref(self._h)))

    def close(self):
        if self._h:
            _L().b2_verify_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_images(self, cameras: list, keypoints_xy: list) -> None:
        n = len(cameras)
        cams = (Camera * max(n, 1))(*cameras)
        self._xy = [np.ascontiguousarray(k, dtype=np.float64).reshape(-1, 2) for k in keypoints_xy]
        ptrs = (C.c_void_p * max(n, 1))(*[k.ctypes.data for k in self._xy])
        cnt = (C.c_int32 * max(n, 1))(*[len(k) for k in self._xy])
        check(_L().b2_verify_set_images(self._h, n, C.cast(cams, C.c_void_p), ptrs, cnt))

    def verify_pairs(self, pairs, match_offsets, matches, options: TwoViewOptions | None = None, seeds=None):
        """Returns (results structured array [n_pairs], inlier_matches uint32 [total,2]); the
        inliers of pair p are inlier_matches[match_offsets[p] : match_offsets[p] + n_inliers[p]]."""
        options = options or TwoViewOptions.default()
        pr = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        off = np.ascontiguousarray(match_offsets, dtype=np.int64)
        mt = np.ascontiguousarray(matches, dtype=np.uint32).reshape(-1, 2)
        n = len(pr)
        assert len(off) == n + 1 and off[-1] == len(mt)
        sd = np.ascontiguousarray(seeds if seeds is not None else np.arange(n), dtype=np.uint32)
        res = np.zeros(n, dtype=RESULT_DTYPE)
        inl = np.zeros((max(len(mt), 1), 2), dtype=np.uint32)
        check(_L().b2_verify_pairs(self._h, n, pr.ctypes.data, off.ctypes.data, mt.ctypes.data, C.byref(options),
                                   sd.ctypes.data, res.ctypes.data, inl.ctypes.data))
        return res, inl[: len(mt)]

    def verify_pairs_multiple(self, pairs, match_offsets, matches, options: TwoViewOptions | None = None, seeds=None,
                              multiple_ignore_watermark: bool = True):
        """TwoViewGeometry::EstimateMultiple for every pair (options.multiple_models of the reference's
        verifier, matching.cc:595-598).  Same layout as verify_pairs; config 8 = MULTIPLE."""
        options = options or TwoViewOptions.default()
        pr = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        off = np.ascontiguousarray(match_offsets, dtype=np.int64)
        mt = np.ascontiguousarray(matches, dtype=np.uint32).reshape(-1, 2)
        n = len(pr)
        assert len(off) == n + 1 and off[-1] == len(mt)
        sd = np.ascontiguousarray(seeds if seeds is not None else np.arange(n), dtype=np.uint32)
        res = np.zeros(n, dtype=RESULT_DTYPE)
        inl = np.zeros((max(len(mt), 1), 2), dtype=np.uint32)
        check(_L().b2_verify_pairs_multiple(self._h, n, pr.ctypes.data, off.ctypes.data, mt.ctypes.data,
                                            C.byref(options), 1 if multiple_ignore_watermark else 0, sd.ctypes.data,
                                            res.ctypes.data, inl.ctypes.data))
        return res, inl[: len(mt)]

    def relative_pose(self, pairs, match_offsets, results, inlier_matches):
        """TwoViewGeometry::EstimateWithRelativePose's post-processing (two_view_geometry.cc:239-289) of the output
        of verify_pairs: structured array [n_pairs] with qvec (w,x,y,z), tvec, tri_angle, config (PLANAR_OR_PANORAMIC
        resolved), n_points3D.  Pairs with a camera lacking a prior focal length keep qvec = tvec = 0 (TwoViewGeometry()), exactly
        the pairs for which TwoViewGeometry::Estimate takes the uncalibrated path."""
        pr = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        off = np.ascontiguousarray(match_offsets, dtype=np.int64)
        res = np.ascontiguousarray(results, dtype=RESULT_DTYPE)
        inl = np.ascontiguousarray(inlier_matches, dtype=np.uint32).reshape(-1, 2)
        n = len(pr)
        assert len(off) == n + 1 and len(res) == n and off[-1] <= max(len(inl), 0) + (0 if len(inl) else off[-1])
        poses = np.zeros(n, dtype=POSE_DTYPE)
        check(_L().b2_verify_relative_pose(self._h, n, pr.ctypes.data, off.ctypes.data, res.ctypes.data,
                                           inl.ctypes.data if len(inl) else None, poses.ctypes.data))
        return poses

    def relative_pose_device(self, n_pairs, pairs_ptr, off_ptr, results_ptr, inl_ptr, poses_ptr):
        check(_L().b2_verify_relative_pose_device(self._h, n_pairs, C.c_void_p(pairs_ptr), C.c_void_p(off_ptr),
                                                  C.c_void_p(results_ptr), C.c_void_p(inl_ptr), C.c_void_p(poses_ptr)))

    def verify_pairs_device(self, n_pairs, pairs_ptr, off_ptr, matches_ptr, options, seeds_ptr, results_ptr, inl_ptr):
        check(_L().b2_verify_pairs_device(self._h, n_pairs, C.c_void_p(pairs_ptr), C.c_void_p(off_ptr),
                                          C.c_void_p(matches_ptr), C.byref(options), C.c_void_p(seeds_ptr),
                                          C.c_void_p(results_ptr), C.c_void_p(inl_ptr)))

    def score_models(self, est_type, xy1, xy2, models, max_residual):
        """Estimator::Residuals + InlierSupportMeasurer::Evaluate -> (counts, sums, masks)."""
        a = np.ascontiguousarray(xy1, dtype=np.float64).reshape(-1, 2)
        b = np.ascontiguousarray(xy2, dtype=np.float64).reshape(-1, 2)
        m = np.ascontiguousarray(models, dtype=np.float64).reshape(-1, 9)
        counts = np.zeros(len(m), dtype=np.int32)
        sums = np.zeros(len(m))
        masks = np.zeros((len(m), max(len(a), 1)), dtype=np.uint8)
        check(_L().b2_score_models(self._h, est_type, len(a), a.ctypes.data, b.ctypes.data, len(m), m.ctypes.data,
                                   float(max_residual), counts.ctypes.data, sums.ctypes.data, masks.ctypes.data))
        return counts, sums, masks[:, : len(a)].astype(bool)

    def debug_sample_stream(self, seed, total, k, n_trials):
        out = np.zeros((n_trials, k), dtype=np.int32)
        check(_L().b2_verify_debug_sample_stream(self._h, seed, total, k, n_trials, out.ctypes.data))
        return out

    def debug_solve(self, est_type, xy1, xy2):
        a = np.ascontiguousarray(xy1, dtype=np.float64).reshape(-1, 2)
        b = np.ascontiguousarray(xy2, dtype=np.float64).reshape(-1, 2)
        out = np.zeros((10, 3, 3))
        nm = C.c_int32(0)
        check(_L().b2_verify_debug_solve(self._h, est_type, len(a), a.ctypes.data, b.ctypes.data, out.ctypes.data,
                                         C.byref(nm)))
        return out[: nm.value]

    def debug_normalized(self, image: int):
        """Camera::ImageToWorld of the image's keypoints as the verifier holds them (test hook)."""
        out = np.zeros((max(len(self._xy[image]), 1), 2))
        check(_L().b2_verify_debug_normalized(self._h, image, out.ctypes.data))
        return out[: len(self._xy[image])]

    def last_kernel_seconds(self) -> float:
        t = C.c_double(0)
        check(_L().b2_verify_last_timing(self._h, C.byref(t)))
        return t.value
