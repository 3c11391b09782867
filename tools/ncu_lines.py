"""Map an `ncu --set full --import-source on` report to CUDA source lines (executed instructions and
stall samples per line) by aligning ncu's SASS page with `nvdisasm -g` line info of the in-tree library.
usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [top_n]"""
import collections
import csv
import glob
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    td = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "ahsoka_b200", "lib", "libahsoka_b200.so")], cwd=td,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    sass = subprocess.run(["nvdisasm", "-g", "-c"] + glob.glob(os.path.join(td, "*.cubin")), stdout=subprocess.PIPE, text=True).stdout.split("\n")
    start = [i for i, l in enumerate(sass) if l.startswith("\t.section\t.text.") and re.search(kern, l)]
    if not start:
        start = [i for i, l in enumerate(sass) if ".text." in l and re.search(kern, l)]
    pat_line = re.compile(r'//## File "([^"]+)", line (\d+)')
    pat_ins = re.compile(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);')
    cur, seq = None, []
    for l in sass[start[0] + 1:]:
        if (l.startswith("\t.section") or l.startswith(".section")) and seq:
            break
        m = pat_line.search(l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = pat_ins.match(l)
        if m:
            seq.append(cur)
    if rep.endswith(".csv"):      # `ncu -i x.ncu-rep --page source --csv --kernel-id :::N` exported on the GPU box
        out = open(rep).read()
    else:
        ncu_kern = os.environ.get("NCU_KERNEL", kern)
        out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + ncu_kern], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.split("\n")))
    hi = [i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r]
    H = rows[hi[0]]
    iex, ist = H.index("Instructions Executed"), H.index("# Samples")
    data = rows[hi[0] + 1:(hi[1] if len(hi) > 1 else len(rows))]
    data = [r for r in data if len(r) > iex]
    agg, samp = collections.Counter(), collections.Counter()
    for k in range(min(len(data), len(seq))):
        agg[seq[k]] += int(data[k][iex] or 0)
        samp[seq[k]] += int(data[k][ist] or 0)
    tot, ts = sum(agg.values()), max(1, sum(samp.values()))
    print(f"kernel {kern}: {len(seq)} SASS instructions, {tot} executed warp-instructions (first captured launch)")
    src = {}
    for k, c in agg.most_common(top):
        f, ln = k if k else ("?", 0)
        if f not in src:
            p = os.path.join(ROOT, "ahsoka_b200", "csrc", f)
            src[f] = open(p).read().split("\n") if os.path.exists(p) else []
        line = src[f][ln - 1].strip() if 0 < ln <= len(src[f]) else ""
        print(f"{100 * c / tot:5.1f}% inst {100 * samp[k] / ts:5.1f}% samp  {f}:{ln}  {line[:120]}")


if __name__ == "__main__":
    main()
