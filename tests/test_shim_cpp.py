"""The C++ adaptors with the reference's signatures (include/dagsfm_b200/colmap_shim.hpp)
compile against the C ABI and link with the library (CPU); the replayed reference test runs
on the GPU."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
EXE = ROOT / "tests" / "cpp" / "_shim_test"


def build():
    from dagsfm_b200 import build as b
    b.build()
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests/cpp/shim_test.cc"),
           "-o", str(EXE), str(b.LIB), f"-Wl,-rpath,{b.LIB.parent}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return EXE


def test_shim_compiles_and_links():
    assert build().exists()


@pytest.mark.gpu
def test_shim_replays_reference_test_on_gpu():
    exe = EXE if EXE.exists() else build()
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "shim ok" in r.stdout
