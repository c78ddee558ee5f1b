#!/usr/bin/env python
"""Roofline denominators MEASURED_PEAKS.json does not hold, measured on the box (one JSON line):
FP64 / FP32 FMA throughput of the CUDA cores (b2_measure_*_peak, dagsfm_b200/csrc/peaks.cu) and the dense int8 GEMM
rate of the library (torch._int_mm, cuBLASLt) measured the way MEASURED_PEAKS.json measures bf16 (8192^3, best of 10)."""
import ctypes as C
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def measure(device=0):
    import torch
    from dagsfm_b200 import lib
    L = lib()
    out = {}
    for name in ("fp64", "ffma"):
        v = C.c_double(0)
        fn = getattr(L, f"b2_measure_{name}_peak")
        fn.argtypes = [C.c_int, C.POINTER(C.c_double)]
        rc = fn(device, C.byref(v))
        out[f"{name}_tflops"] = v.value if rc == 0 else None
    try:
        dev = torch.device("cuda", device)
        a = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device=dev)
        b = torch.randint(-128, 127, (8192, 8192), dtype=torch.int8, device=dev)
        best = 0.0
        for _ in range(3):
            torch._int_mm(a, b)
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch._int_mm(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 2 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        out["i8_tops_library_gemm"] = best
    except Exception as e:   # a reporting extra
        out["i8_tops_library_gemm"] = None
        out["i8_error"] = repr(e)
    return out


if __name__ == "__main__":
    print(json.dumps(measure()))
