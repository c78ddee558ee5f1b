"""Compiles CUDA sources of the library for the HOST against tests/cuda_emu/cuda_emu.h.

The only source transformation is syntactic: `kernel<<<grid, block, smem, stream>>>(args);` becomes
`cuda_emu::launch(grid, block, smem, [&] { kernel(args); });` and `extern __shared__ T name[];` becomes a
pointer into the launch's dynamic shared memory.  Everything else is handled by macros / inline
functions of cuda_emu.h.  TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import re
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
CSRC = ROOT / "dagsfm_b200" / "csrc"
import os

# B2_EMU_BUILD_DIR: a second build tree, so that a run in another mode (tests/test_emu_modes.py) does not replace the
# libraries the parent process has loaded
OUT = Path(os.environ["B2_EMU_BUILD_DIR"]) if os.environ.get("B2_EMU_BUILD_DIR") else HERE / "_build"


# the translation units of each emulated library (the product's own sources)
BA_SOURCES = ["common.cu", "match_post.cu", "ba_kernels.cu", "ba_fused.cu", "ba_chol.cu", "ba_iterative.cu", "ba_metrics.cu", "ba_api.cu"]
VERIFY_SOURCES = ["common.cu", "verify_kernel.cu", "verify_pose.cu", "verify_api.cu"]
RETRIEVAL_SOURCES = ["common.cu", "match_post.cu", "retrieval.cu"]


def _matching(s: str, i: int, open_c: str, close_c: str) -> int:
    depth = 0
    for k in range(i, len(s)):
        if s[k] == open_c:
            depth += 1
        elif s[k] == close_c:
            depth -= 1
            if depth == 0:
                return k
    raise ValueError("unbalanced")


def _split_top(s: str) -> list[str]:
    parts, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def rewrite(src: str) -> str:
    out, pos = "", 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            break
        # kernel name: identifier / scope / template arguments immediately before <<<
        j = i
        while j > 0 and src[j - 1].isspace():
            j -= 1
        if src[j - 1] == ">":                       # template arguments
            depth, k = 0, j - 1
            while True:
                if src[k] == ">":
                    depth += 1
                elif src[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                k -= 1
            j = k
        while j > 0 and (src[j - 1].isalnum() or src[j - 1] in "_:"):
            j -= 1
        name = src[j:i].strip()
        e = src.find(">>>", i)
        cfg = _split_top(src[i + 3:e])
        while len(cfg) < 3:
            cfg.append("0")
        a0 = src.find("(", e)
        a1 = _matching(src, a0, "(", ")")
        semi = src.find(";", a1)
        args = src[a0 + 1:a1]
        out += src[pos:j]
        out += f"cuda_emu::launch({cfg[0]}, {cfg[1]}, {cfg[2]}, [&] {{ {name}({args}); }});"
        pos = semi + 1
    out += src[pos:]
    out = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"\1* \2 = (\1*)cuda_emu::dyn_smem();", out)
    return out


def build(name: str, sources: list[str], extra: list[str] | None = None) -> Path:
    import os
    # every "device" allocation of the emulated libraries ends at an inaccessible page: an out-of-bounds access
    # fails at the access (with a native backtrace) instead of depending on the heap layout
    os.environ.setdefault("B2_EMU_GUARD", "1")
    # ... and is inaccessible to host code outside kernels / cudaMemcpy / cuSOLVER calls: a host dereference of a device
    # pointer (which the device would punish) faults here too
    os.environ.setdefault("B2_EMU_DEVMEM", "1")
    OUT.mkdir(exist_ok=True)
    gen = []
    for s in sources:
        p = CSRC / s
        g = OUT / (p.stem + "_emu.cc")
        g.write_text(f'#line 1 "{p}"\n' + rewrite(p.read_text()))
        gen.append(str(g))
    # test hooks that stand in for device-side code (the gloo all-reduce of the multi-rank tests works on the emulated
    # device buffer directly) open the same window a kernel gets when B2_EMU_DEVMEM=1 protects device memory from host code
    exp = OUT / f"{name}_exports_emu.cc"
    exp.write_text('#include <cuda_runtime.h>\nextern "C" void cuda_emu_device_window(int on) { cuda_emu::device_access(on != 0); }\n')
    gen.append(str(exp))
    lib = OUT / f"libemu_{name}.so"
    cxx = "/usr/bin/g++" if Path("/usr/bin/g++").exists() else "g++"
    import os
    san = ["-fsanitize=address", "-fno-omit-frame-pointer"] if os.environ.get("B2_EMU_ASAN") else []   # debugging aid
    if os.environ.get("B2_EMU_UBSAN"):   # misaligned vector loads / stores (a device fault) trap instead of passing silently
        san += ["-fsanitize=alignment", "-fsanitize-undefined-trap-on-error"]
    cmd = [cxx, "-std=c++17", "-O1", "-g", "-ffp-contract=off", "-U_FORTIFY_SOURCE", "-fno-gnu-unique", "-shared", "-fPIC", "-w"] + san + [
           "-I", str(HERE / "include"), "-I", str(CSRC), "-I", str(ROOT / "include"), "-o", str(lib)] + gen + (extra or [])
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr[-6000:])
    return lib


if __name__ == "__main__":
    import sys
    print(build(sys.argv[1], sys.argv[2:]))
