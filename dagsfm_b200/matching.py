"""Host-side mirror of the reference's GPU matcher interface.

Reference names kept (src/feature/sift.h:116-165,229-239; lib/SiftGPU/SiftGPU.h:276-373):
  SiftMatchingOptions, SiftMatchGPU.{SetDescriptors, GetSiftMatch}, MatchSiftFeaturesGPU
plus the batched seam `SiftMatchGPU.set_images / match_pairs` that replaces the
per-pair loop of SiftGPUFeatureMatcher::Run (src/feature/matching.cc:376-427).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import MatchOptions, check, lib


@dataclass
class SiftMatchingOptions:
    """Subset of reference SiftMatchingOptions that drives the matcher (sift.h:116-165)."""
    max_ratio: float = 0.8
    max_distance: float = 0.7
    cross_check: bool = True
    max_num_matches: int = 32768
    max_error: float = 4.0          # guided matching only (sift.h:143)

    def check(self) -> bool:  # SiftMatchingOptions::Check, sift.cc:236-250
        return self.max_ratio > 0 and self.max_distance > 0 and self.max_num_matches > 0 and self.max_error > 0

    def to_c(self) -> MatchOptions:
        return MatchOptions(np.float32(self.max_ratio), np.float32(self.max_distance),
                            1 if self.cross_check else 0, int(self.max_num_matches))


# b2_guided_geometry
GUIDED_GEOMETRY_DTYPE = np.dtype([("config", "<i4"), ("reserved", "<i4"), ("F", "<f8", (9,)), ("H", "<f8", (9,))])


def _as_desc(d) -> np.ndarray:
    a = np.ascontiguousarray(d, dtype=np.uint8)
    if a.ndim != 2 or (a.shape[0] > 0 and a.shape[1] != 128):
        if a.size == 0:
            return a.reshape(0, 128)
        raise ValueError("descriptors must be N x 128 uint8 (FeatureDescriptors)")
    return a


class SiftMatchGPU:
    """One matcher per GPU, single-threaded (as SiftMatchGPU in the reference)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().b2_match_create(device, C.byref(self._h)))
        self.gpu_index = device
        self._n_images = 0

    def close(self):
        if self._h:
            lib().b2_match_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- two-slot seam ---------------------------------------------------
    def SetDescriptors(self, index: int, descriptors) -> None:
        """SiftMatchGPU::SetDescriptors(index, num, ptr); None keeps the previous upload."""
        if descriptors is None:
            check(lib().b2_match_set_descriptors(self._h, index, 0, None))
            return
        d = _as_desc(descriptors)
        check(lib().b2_match_set_descriptors(self._h, index, d.shape[0], d.ctypes.data_as(C.c_void_p)))

    def GetSiftMatch(self, options: SiftMatchingOptions) -> np.ndarray:
        """Returns FeatureMatches as uint32 [n,2] (ascending idx1)."""
        opt = options.to_c()
        out = np.empty((options.max_num_matches, 2), dtype=np.uint32)
        n = C.c_int32(0)
        check(lib().b2_match_run(self._h, C.byref(opt), out.ctypes.data_as(C.c_void_p), C.byref(n)))
        return out[: n.value].copy()

    # ---- batched seam ----------------------------------------------------
    def set_images(self, descriptors: list) -> None:
        ds = [_as_desc(d) for d in descriptors]
        n = len(ds)
        ptrs = (C.c_void_p * max(n, 1))(*[d.ctypes.data for d in ds])
        cnt = (C.c_int32 * max(n, 1))(*[d.shape[0] for d in ds])
        check(lib().b2_match_set_images(self._h, n, ptrs, cnt))
        self._n_images = n
        self._max_n = max([d.shape[0] for d in ds], default=0)

    def set_images_device(self, desc_dev_ptr: int, row_offset: np.ndarray, n_desc: np.ndarray) -> None:
        ro = np.ascontiguousarray(row_offset, dtype=np.int64)
        nd = np.ascontiguousarray(n_desc, dtype=np.int32)
        check(lib().b2_match_set_images_device(
            self._h, len(nd), C.c_void_p(desc_dev_ptr),
            ro.ctypes.data_as(C.POINTER(C.c_int64)), nd.ctypes.data_as(C.POINTER(C.c_int32))))
        self._n_images = len(nd)
        self._max_n = int(nd.max()) if len(nd) else 0

    def match_pairs(self, pairs, options: SiftMatchingOptions, capacity: int | None = None):
        """pairs: [n,2] image indices.  Returns (offsets int64[n+1], matches uint32[total,2])."""
        pr = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        n = pr.shape[0]
        if capacity is None:
            capacity = n * min(self._max_n, options.max_num_matches)
        opt = options.to_c()
        offsets = np.zeros(n + 1, dtype=np.int64)
        matches = np.empty((max(capacity, 1), 2), dtype=np.uint32)
        total = C.c_int64(0)
        check(lib().b2_match_pairs(self._h, n, pr.ctypes.data_as(C.c_void_p), C.byref(opt),
                                   offsets.ctypes.data_as(C.c_void_p),
                                   matches.ctypes.data_as(C.c_void_p), capacity, C.byref(total)))
        return offsets, matches[: total.value]

    # ---- guided matching (MatchGuidedSiftFeaturesGPU, sift.cc:987-1066) ----
    def set_keypoints(self, keypoints_xy: list) -> None:
        """(x, y) of every keypoint of the images of the last set_images call."""
        ks = [np.ascontiguousarray(k, dtype=np.float32).reshape(-1, 2) for k in keypoints_xy]
        n = len(ks)
        ptrs = (C.c_void_p * max(n, 1))(*[k.ctypes.data for k in ks])
        cnt = (C.c_int32 * max(n, 1))(*[k.shape[0] for k in ks])
        check(lib().b2_match_set_keypoints(self._h, n, ptrs, cnt))

    def match_guided_pairs(self, pairs, geometries, options: SiftMatchingOptions, capacity: int | None = None):
        """geometries: per pair (config, F 3x3 or None, H 3x3 or None), TwoViewGeometry fields.
        Returns (offsets int64[n+1], matches uint32[total,2]); a pair whose config has no guided
        filter gets no matches (the reference leaves inlier_matches untouched there)."""
        pr = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        n = pr.shape[0]
        assert len(geometries) == n
        geo = np.zeros(n, dtype=GUIDED_GEOMETRY_DTYPE)
        for k, (cfg, F, H) in enumerate(geometries):
            geo["config"][k] = int(cfg)
            geo["F"][k] = np.zeros(9) if F is None else np.asarray(F, dtype=np.float64).reshape(9)
            geo["H"][k] = np.zeros(9) if H is None else np.asarray(H, dtype=np.float64).reshape(9)
        if capacity is None:
            capacity = n * min(self._max_n, options.max_num_matches)
        opt = options.to_c()
        offsets = np.zeros(n + 1, dtype=np.int64)
        matches = np.empty((max(capacity, 1), 2), dtype=np.uint32)
        total = C.c_int64(0)
        check(lib().b2_match_guided_pairs(self._h, n, pr.ctypes.data_as(C.c_void_p), geo.ctypes.data_as(C.c_void_p),
                                          float(options.max_error), C.byref(opt), offsets.ctypes.data_as(C.c_void_p),
                                          matches.ctypes.data_as(C.c_void_p), capacity, C.byref(total)))
        return offsets, matches[: total.value]

    def match_pairs_device(self, n_pairs: int, pairs_dev_ptr: int, options: SiftMatchingOptions,
                           offsets_dev_ptr: int, matches_dev_ptr: int, capacity: int) -> int:
        opt = options.to_c()
        total = C.c_int64(0)
        check(lib().b2_match_pairs_device(self._h, n_pairs, C.c_void_p(pairs_dev_ptr), C.byref(opt),
                                          C.c_void_p(offsets_dev_ptr), C.c_void_p(matches_dev_ptr),
                                          capacity, C.byref(total)))
        return total.value

    def match_guided_pairs_device(self, n_pairs: int, pairs_dev_ptr: int, results_dev_ptr: int, min_num_inliers: int,
                                  options: SiftMatchingOptions, offsets_dev_ptr: int, matches_dev_ptr: int, capacity: int) -> int:
        """The guided stage chained onto verify_pairs_device: geometries are read from the verifier's device results."""
        opt = options.to_c()
        total = C.c_int64(0)
        L = lib()
        L.b2_match_guided_pairs_device.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32, C.c_double,
                                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        check(L.b2_match_guided_pairs_device(self._h, n_pairs, C.c_void_p(pairs_dev_ptr), C.c_void_p(results_dev_ptr),
                                             int(min_num_inliers), float(options.max_error), C.byref(opt),
                                             C.c_void_p(offsets_dev_ptr), C.c_void_p(matches_dev_ptr), capacity, C.byref(total)))
        return total.value

    def last_timing(self) -> dict:
        tc, al = C.c_double(0), C.c_double(0)
        nl, nc = C.c_int64(0), C.c_int64(0)
        check(lib().b2_match_last_timing(self._h, C.byref(tc), C.byref(al), C.byref(nl), C.byref(nc)))
        return {"tc_kernel_s": tc.value, "all_kernels_s": al.value, "tc_launches": nl.value,
                "fixup_candidates": nc.value}


def match_guided_sift_features_gpu(options: SiftMatchingOptions, keypoints1, keypoints2, descriptors1, descriptors2,
                                   sift_match_gpu: SiftMatchGPU, config: int, F=None, H=None):
    """MatchGuidedSiftFeaturesGPU (sift.cc:987-1066) for one pair: returns the new inlier_matches
    (uint32 [n,2]) or None when `config` has no guided filter (inlier_matches stay as they were)."""
    assert options.check()
    if int(config) not in (2, 3, 4, 5, 6):
        return None
    d1 = _as_desc(descriptors1)[: options.max_num_matches]
    d2 = _as_desc(descriptors2)[: options.max_num_matches]
    k1 = np.asarray(keypoints1, dtype=np.float32).reshape(-1, 2)[: options.max_num_matches]
    k2 = np.asarray(keypoints2, dtype=np.float32).reshape(-1, 2)[: options.max_num_matches]
    sift_match_gpu.set_images([d1, d2])
    sift_match_gpu.set_keypoints([k1, k2])
    _, m = sift_match_gpu.match_guided_pairs([(0, 1)], [(config, F, H)], options)
    return m.copy()


def match_sift_features_gpu(options: SiftMatchingOptions, descriptors1, descriptors2,
                            sift_match_gpu: SiftMatchGPU) -> np.ndarray:
    """MatchSiftFeaturesGPU (sift.cc:941-985): None reuses the previous upload; features
    beyond max_num_matches are clamped at upload time (SiftMatchCU.cpp:108)."""
    assert options.check()
    if descriptors1 is not None:
        sift_match_gpu.SetDescriptors(0, _as_desc(descriptors1)[: options.max_num_matches])
    if descriptors2 is not None:
        sift_match_gpu.SetDescriptors(1, _as_desc(descriptors2)[: options.max_num_matches])
    return sift_match_gpu.GetSiftMatch(options)
