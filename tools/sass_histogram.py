"""SASS opcode histogram of the in-tree CUDA library (evidence of what the binary does or does not use: bulk asynchronous
copies UBLKCP, mbarrier SYNCS, vector loads LDG.E.128, warp reductions REDUX, shared atomics ATOMS, local memory LDL/STL ...).
usage: python tools/sass_histogram.py [out.md]"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ahsoka_b200", "lib", "libahsoka_b200.so")
WATCH = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "LDG.E.128", "LDG.E.64", "LDG.E", "STG.E.128", "LDS.128", "LDS.64", "ATOMS", "ATOMG", "RED",
         "REDUX", "SHFL", "VOTE", "POPC", "BAR", "LDL", "STL", "IMAD", "VIMNMX", "UMOV"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True).stdout
    per_kernel, cur = collections.OrderedDict(), None
    pat = re.compile(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)")
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0].replace("ahs::", "").replace("void ", "")
            per_kernel[cur] = collections.Counter()
            continue
        m = pat.search(line)
        if m and cur:
            per_kernel[cur][m.group(1)] += 1
    total = collections.Counter()
    for c in per_kernel.values():
        total.update(c)

    def watch(c):
        r = {}
        for w in WATCH:
            if w == "LDG.E":
                r[w] = sum(v for k, v in c.items() if k.startswith("LDG.E") and ".128" not in k and ".64" not in k)
            else:
                r[w] = sum(v for k, v in c.items() if k.startswith(w))
        return r
    lines = ["# SASS opcode histogram of `ahsoka_b200/lib/libahsoka_b200.so` (sm_100a)", "",
             f"{len(per_kernel)} kernels, {sum(total.values())} SASS instructions.  `cuobjdump -sass`, counted by `tools/sass_histogram.py`.", "",
             "| opcode (prefix) | whole library |", "|---|---|"]
    tw = watch(total)
    for w in WATCH:
        lines.append(f"| `{w}` | {tw[w]} |")
    lines += ["", "Per kernel (instructions; bulk copies `UBLKCP`; mbarrier `SYNCS`; 128-bit global loads; shared atomics; warp reductions `REDUX`; local memory `LDL`+`STL`):", "",
              "| kernel | SASS instructions | UBLKCP | SYNCS | LDG.E.128 | ATOMS | REDUX | LDL+STL |", "|---|---|---|---|---|---|---|---|"]
    for k, c in per_kernel.items():
        w = watch(c)
        lines.append(f"| `{k}` | {sum(c.values())} | {w['UBLKCP']} | {w['SYNCS']} | {w['LDG.E.128']} | {w['ATOMS']} | {w['REDUX']} | {w['LDL'] + w['STL']} |")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text)
    print(text[:3000])


if __name__ == "__main__":
    main()
