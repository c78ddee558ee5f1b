#!/usr/bin/env python
"""Generates tests/golden/e5_poly_golden.npz from the REFERENCE's own generated
5-point polynomial code (run in the build container only; /root/reference does
not exist on the GPU box):

  src/estimators/essential_matrix_poly.h    a[200] (10x20, column-major) from e[36]
  src/estimators/essential_matrix_coeffs.h  coeffs(0..10) from b[39] (13x3 column-major)

The headers are plain arithmetic; they are parsed as text, the C expressions are
evaluated with Python floats on seeded random inputs, and only the numeric
inputs/outputs are stored (no reference source is copied into the repo).
"""
import re
import sys
from pathlib import Path

import numpy as np

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent / "e5_poly_golden.npz"


def statements(path):
    txt = path.read_text()
    txt = re.sub(r"//[^\n]*", "", txt)
    body = txt[txt.index("{") + 1: txt.rindex("}")]
    return [s.strip().replace("\n", " ") for s in body.split(";") if "=" in s]


def eval_poly(e):
    e = list(map(float, e))
    e2 = [v * v for v in e]
    e3 = [v * v * v for v in e]
    a = [0.0] * 200
    for s in statements(REF / "src/estimators/essential_matrix_poly.h"):
        m = re.search(r"\ba\[(\d+)\]\s*=\s*(.*)$", s)
        if m:
            a[int(m.group(1))] = eval(m.group(2), {"e": e, "e2": e2, "e3": e3})
    return np.array(a)


def eval_coeffs(b):
    b = list(map(float, b))
    c = [0.0] * 11
    for s in statements(REF / "src/estimators/essential_matrix_coeffs.h"):
        m = re.search(r"\bcoeffs\((\d+)\)\s*=\s*(.*)$", s)
        if m:
            c[int(m.group(1))] = eval(m.group(2), {"b": b})
    return np.array(c)


def main():
    rng = np.random.default_rng(20260923)
    n = 8
    e_in = rng.normal(size=(n, 36))     # E.data(): 9x4 column-major -> e[9*k + i] = basis k, entry i
    b_in = rng.normal(size=(n, 39))     # B.data(): 13x3 column-major
    b_in[:, [0, 4, 8]] = rng.normal(size=(n, 3))
    a_out = np.stack([eval_poly(e) for e in e_in])
    assert (a_out != 0).all(), "a statement of the generated header was not parsed"
    c_out = np.stack([eval_coeffs(b) for b in b_in])
    np.savez(OUT, e_in=e_in, a_out=a_out, b_in=b_in, c_out=c_out)
    print("wrote", OUT, a_out.shape, c_out.shape)


if __name__ == "__main__":
    main()
