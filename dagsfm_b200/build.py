"""In-tree build of the sm_100a shared library (nvcc cross-compiles without a GPU).

    python -m dagsfm_b200.build          # build if stale
    python -m dagsfm_b200.build --force [-v]

Each .cu is compiled to an object (in parallel) and the objects are linked into
dagsfm_b200/libdagsfm_b200.so.  The two-view verifier and bundle-adjustment
translation units are built with --fmad=false: their FP64 arithmetic must be the
same IEEE operations as the reference's scalar C++ (no FMA contraction).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "_obj"
LIB = PKG / "libdagsfm_b200.so"

COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
          "-Xcompiler", "-fPIC"]
NO_FMA_PREFIXES = ("verify_", "ba_", "match_guided", "retrieval")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*")) + [PKG.parent / "include" / "dagsfm_b200.h"]
    return any(p.stat().st_mtime > t for p in deps)


def _compile(nvcc: str, src: Path, verbose: bool) -> tuple[Path, str]:
    obj = OBJ / (src.stem + ".o")
    hdr_t = max([p.stat().st_mtime for p in list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) +
                 [PKG.parent / "include" / "dagsfm_b200.h"]])
    if obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, hdr_t) and not verbose:
        return obj, ""
    cmd = [nvcc, *COMMON, "-c", "-o", str(obj), str(src)]
    if src.name.startswith(NO_FMA_PREFIXES):
        cmd.insert(1, "--fmad=false")
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}{r.stderr}")
    return obj, r.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    OBJ.mkdir(exist_ok=True)
    if force:
        for o in OBJ.glob("*.o"):
            o.unlink()
    with ThreadPoolExecutor(max_workers=8) as ex:
        res = list(ex.map(lambda s: _compile(nvcc, s, verbose), sources()))
    if verbose:
        for _, log in res:
            sys.stderr.write(log)
    cuda_lib = str(Path(nvcc).resolve().parent.parent / "lib64")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "static",
           "-o", str(LIB), *[str(o) for o, _ in res],
           # dense Cholesky of the reduced camera system (bundle adjustment) only
           f"-L{cuda_lib}", "-lcusolver", "-Xlinker", f"-rpath={cuda_lib}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
    return LIB


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(p)
