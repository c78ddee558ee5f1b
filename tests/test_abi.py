"""CPU-side checks of the drop-in boundary: the library builds for sm_100a, loads,
and exports every symbol include/dagsfm_b200.h declares; without a GPU the entry
points fail loudly (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    txt = (ROOT / "include" / "dagsfm_b200.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from dagsfm_b200 import build as b
    b.build()
    L = C.CDLL(str(b.LIB))
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(L, s), f"{s} declared in the header but not exported"


def test_version_and_launch_counter():
    from dagsfm_b200 import lib
    assert b"sm_100a" in lib().b2_version()
    assert lib().b2_kernel_launch_count() >= 0


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dagsfm_b200 import B2Error, SiftMatchGPU
    with pytest.raises(B2Error) as e:
        SiftMatchGPU(0)
    assert e.value.code in (2, 3)


def test_sass_is_blackwell_native():
    """The matcher kernel must contain tcgen05 MMA (UTCIMMA), TMEM loads (LDTM) and TMA
    (UTMALDG) in its SASS -- evidence table of B200_PROFILING.md."""
    import shutil
    import subprocess
    from dagsfm_b200 import build as b
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.run(["cuobjdump", "-sass", str(b.LIB)], capture_output=True, text=True).stdout
    for mnem in ("UTCIMMA", "LDTM", "UTMALDG", "VIMNMX3"):
        assert mnem in sass, mnem
