"""include/dagsfm_b200/two_view_shim.hpp -- the reference's TwoViewGeometry class over the C ABI.
* the reference's own two_view_geometry_test.cc (TestDefault, TestInvert) replayed on the adaptor (host only);
* Estimate / EstimateUncalibrated / EstimateMultiple / Invert on a synthetic pair, with the adaptor linked against
  the CUDA-emulator build of the library's sources (a CPU test of the adaptor and of the C-ABI call sequence);
* the same program against the product library on a GPU."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "cpp" / "two_view_shim_test.cc"


def _build(lib: Path, exe: Path) -> Path:
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(SRC), "-o", str(exe), str(lib),
           f"-Wl,-rpath,{lib.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _product_exe() -> Path:
    from dagsfm_b200 import build as b
    b.build()
    return _build(b.LIB, ROOT / "tests" / "cpp" / "_two_view_shim_test")


def test_reference_default_and_invert_tests_replayed_on_the_adaptor():
    r = subprocess.run([str(_product_exe())], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host tests ok" in r.stdout


def test_estimate_through_the_adaptor_on_the_emulated_library():
    from tests.cuda_emu.build_emu import BA_SOURCES, VERIFY_SOURCES, build
    lib = build("verify", VERIFY_SOURCES)
    exe = _build(lib, ROOT / "tests" / "cuda_emu" / "_build" / "two_view_shim_test_emu")
    r = subprocess.run([str(exe), "estimate"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "two-view shim ok" in r.stdout


@pytest.mark.gpu
def test_estimate_through_the_adaptor_on_gpu():
    r = subprocess.run([str(_product_exe()), "estimate"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "two-view shim ok" in r.stdout
