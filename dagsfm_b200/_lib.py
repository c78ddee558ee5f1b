"""ctypes binding of the C ABI declared in include/dagsfm_b200.h.

The library is the product: if it is missing this module raises -- there is no
Python / torch / CPU fallback behind any call.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "libdagsfm_b200.so"

B2_OK, B2_ERR_INVALID, B2_ERR_CUDA, B2_ERR_NO_DEVICE, B2_ERR_CAPACITY, B2_ERR_INTERNAL = range(6)


class B2Error(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"dagsfm_b200 error {code}: {msg}")
        self.code = code


class MatchOptions(C.Structure):
    """b2_match_options == SiftMatchingOptions subset (reference src/feature/sift.h:116-165)."""
    _fields_ = [("max_ratio", C.c_float), ("max_distance", C.c_float),
                ("cross_check", C.c_int32), ("max_num_matches", C.c_int32)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m dagsfm_b200.build` "
            "(the CUDA library is the product; there is no fallback)")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, u64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64
    P = C.POINTER
    L.b2_last_error.restype = C.c_char_p
    L.b2_version.restype = C.c_char_p
    L.b2_kernel_launch_count.restype = u64
    L.b2_match_default_options.argtypes = [P(MatchOptions)]
    L.b2_match_default_options.restype = None
    L.b2_match_create.argtypes = [C.c_int, P(vp)]
    L.b2_match_destroy.argtypes = [vp]
    L.b2_match_set_images.argtypes = [vp, i32, P(vp), P(i32)]
    L.b2_match_set_images_device.argtypes = [vp, i32, vp, P(i64), P(i32)]
    L.b2_match_pairs.argtypes = [vp, i64, vp, P(MatchOptions), vp, vp, i64, P(i64)]
    L.b2_match_pairs_device.argtypes = [vp, i64, vp, P(MatchOptions), vp, vp, i64, P(i64)]
    L.b2_match_set_descriptors.argtypes = [vp, C.c_int, i32, vp]
    L.b2_match_run.argtypes = [vp, P(MatchOptions), vp, P(i32)]
    L.b2_match_last_timing.argtypes = [vp, P(C.c_double), P(C.c_double), P(i64), P(i64)]
    L.b2_match_set_keypoints.argtypes = [vp, i32, P(vp), P(i32)]
    L.b2_match_guided_pairs.argtypes = [vp, i64, vp, vp, C.c_double, P(MatchOptions), vp, vp, i64, P(i64)]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != B2_OK:
        raise B2Error(rc, lib().b2_last_error().decode())
