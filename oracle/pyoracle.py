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


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            build()
        L = C.CDLL(str(LIB_PATH))
        vp = C.c_void_p
        L.orc_match_sift.argtypes = [vp, C.c_int, vp, C.c_int, C.c_float, C.c_float, C.c_int, vp, C.c_int]
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
