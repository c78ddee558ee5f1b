#!/usr/bin/env python
"""Per-source-line stall samples of one kernel from an `ncu --set full --import-source on` capture.

  python profiles/source_hot.py gpurun_out/x.ncu-rep [top_n] [kernel-substring]
Lists the source lines by warp-stall samples (the line-level view of ncu's Source page), with executed instruction
counts and the dominant stall reasons, so a kernel's time can be attributed to its code without the GUI."""
import csv
import subprocess
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
    hdr = rows[hdr_i]
    col = {h: i for i, h in enumerate(hdr) if h not in ("Source",)}
    si, ii = col["# Samples"], col["Instructions Executed"]
    stall_cols = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    fname = ""
    lines = []
    for r in rows[hdr_i + 1:]:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if len(r) < len(hdr) or not r[0].isdigit():
            continue
        try:
            smp = int(r[si] or 0)
        except ValueError:
            continue
        st = sorted(((int(r[i] or 0), h) for h, i in stall_cols), reverse=True)[:3]
        lines.append((smp, int(r[ii] or 0), fname, int(r[0]), r[1].strip()[:90], st))
    tot = sum(l[0] for l in lines) or 1
    print(f"# {path}: {tot} stall samples over {len(lines)} source lines")
    for smp, ins, f, ln, src, st in sorted(lines, reverse=True)[:top]:
        stx = " ".join(f"{h[6:]}={v}" for v, h in st if v)
        print(f"{100.0 * smp / tot:5.1f}% {smp:7d} smp {ins:10d} inst  {f}:{ln:<4d} {src:90s} | {stx}")


if __name__ == "__main__":
    main()
