"""include/dagsfm_b200/bundle_adjustment_shim.hpp -- the reference's BundleAdjustmentOptions / BundleAdjustmentConfig /
BundleAdjuster over the C ABI.  The reference's bundle_adjustment_test.cc cases (reduced residual / parameter counts,
which blocks move) are replayed through the adaptor: host-only parts always, Solve against the CUDA-emulator build of
the library's sources on the CPU and against the product library on a GPU."""
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "cpp" / "ba_shim_test.cc"


def _build(lib: Path, exe: Path) -> Path:
    r = subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(SRC), "-o", str(exe), str(lib),
                        f"-Wl,-rpath,{lib.parent}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def _product_exe() -> Path:
    from dagsfm_b200 import build as b
    b.build()
    return _build(b.LIB, ROOT / "tests" / "cpp" / "_ba_shim_test")


def test_config_container_and_num_residuals_replayed():
    r = subprocess.run([str(_product_exe()), "config"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "config tests ok" in r.stdout, r.stdout + r.stderr


def test_reference_structure_tests_through_the_adaptor_on_the_emulated_library():
    from tests.cuda_emu.build_emu import BA_SOURCES, VERIFY_SOURCES, build
    lib = build("ba", BA_SOURCES)
    exe = _build(lib, ROOT / "tests" / "cuda_emu" / "_build" / "ba_shim_test_emu")
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=1800)
    assert r.returncode == 0 and "ba shim ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_reference_structure_tests_through_the_adaptor_on_gpu():
    r = subprocess.run([str(_product_exe())], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ba shim ok" in r.stdout, r.stdout + r.stderr
