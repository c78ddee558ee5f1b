"""Static SASS summary of kernels that have no ncu capture yet (written after the round's GPU budget was spent):
per kernel the instruction count by class and the resource usage, from cuobjdump of the in-tree objects.
NOT a measurement -- it documents what was compiled (FP64 math vs memory vs atomics, registers, local memory)."""
import re
import subprocess
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OBJ = ROOT / "dagsfm_b200" / "_obj"
CLASSES = [("fp64", r"^(DFMA|DMUL|DADD|DSETP|MUFU\.RCP64H|MUFU\.RSQ64H|DMNMX)"), ("ld_global", r"^LDG"), ("st_global", r"^STG"),
           ("atomics", r"^(ATOM|ATOMG|RED)"), ("ld_st_local", r"^(LDL|STL)"), ("ld_st_shared", r"^(LDS|STS)"),
           ("shuffle_vote", r"^(SHFL|VOTE|MATCH)"), ("barrier", r"^(BAR|WARPSYNC)"), ("branch", r"^(BRA|BSSY|BSYNC|CALL|RET|EXIT)"),
           ("int_addr", r"^(IMAD|IADD3|LEA|SHF|LOP3|ISETP|VIADD|MOV|SEL)")]


def demangle(names):
    r = subprocess.run(["c++filt"] + names, capture_output=True, text=True)
    return r.stdout.strip().split("\n")


def main(objs, pattern):
    for obj in objs:
        sass = subprocess.run(["cuobjdump", "-sass", str(OBJ / obj)], capture_output=True, text=True).stdout
        res = subprocess.run(["cuobjdump", "-res-usage", str(OBJ / obj)], capture_output=True, text=True).stdout
        usage = {}
        cur = None
        for line in res.splitlines():
            m = re.search(r"Function (\S+):", line)
            if m:
                cur = m.group(1)
            elif cur and "REG:" in line:
                usage[cur] = " ".join(t for t in line.split() if t.split(":")[0] in ("REG", "STACK", "SHARED"))
        funcs, cur = {}, None
        for line in sass.splitlines():
            m = re.search(r"Function : (\S+)", line)
            if m:
                cur = m.group(1)
                funcs[cur] = Counter()
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
            if cur and m:
                op = m.group(1)
                funcs[cur]["total"] += 1
                for name, rx in CLASSES:
                    if re.match(rx, op):
                        funcs[cur][name] += 1
                        break
        names = [n for n in funcs if re.search(pattern, n)]
        for n, d in zip(names, demangle(names)):
            c = funcs[n]
            short = re.sub(r"\(.*", "", d)
            print(f"{obj}: {short}\n    {usage.get(n, '')} | instructions {c['total']}: " +
                  ", ".join(f"{k} {c[k]}" for k, _ in CLASSES if c[k]))


if __name__ == "__main__":
    main(sys.argv[2:], sys.argv[1])
