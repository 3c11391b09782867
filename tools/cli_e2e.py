"""Stage times of the drop-in CLI (`ahsoka_b200/bin/Ahsoka_b200 phase -g G -a A -o P`) on a GFA/GAF rendering of a
BASELINE workload sample: reference host translation units for GFA parsing / bubble detection, this repo's
GAF reader, allele-path enumeration, flattening, CUDA phasing and emission.
    python tools/cli_e2e.py --workload cfg2 --scale 0.1 [--host reference] [--exe ahsoka_b200/bin/Ahsoka_b200]
Prints one JSON line: {"stages_ms": {...}, "wall_s": ..., "lines": ..., "chains": ...}.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ahsoka_b200 import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--scale", type=float, default=0.1)
    ap.add_argument("--host", default="native", choices=["native", "reference"])
    ap.add_argument("--exe", default=os.path.join(ROOT, "ahsoka_b200", "bin", "Ahsoka_b200"))
    ap.add_argument("--threads", type=int, default=1, help="-t of the CLI (the reader uses all cores when 1)")
    ap.add_argument("--repeat", type=int, default=2, help="runs; the last one is reported (first warms the page cache / CUDA context)")
    a = ap.parse_args()
    a.exe = os.path.abspath(a.exe)
    with tempfile.TemporaryDirectory() as td:
        b = synth.generate(synth.config(a.workload, a.scale), os.path.join(td, "s"))
        env = dict(os.environ, AHSOKA_TIMING="1", AHSOKA_PLOIDY=str(int(b.ploidy)))
        if os.environ.get("CLI_E2E_VERBOSE"):
            print("affinity", len(os.sched_getaffinity(0)), file=sys.stderr)
        if a.host == "reference":
            env["AHSOKA_HOST"] = "reference"
        out = None
        counts = {"chains": int(b.n_chains)}
        del b                                   # the generator's copy of the batch is not needed while the CLI runs
        for _ in range(a.repeat):
            t0 = time.time()
            r = subprocess.run([a.exe, "phase", "-g", "s.gfa", "-a", "s.gaf", "-o", "out", "-t", str(a.threads)], cwd=td, env=env,
                               stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True)
            wall = time.time() - t0
            if r.returncode != 0:
                sys.exit("CLI failed: " + r.stderr[-2000:])
            stages = {}
            for l in r.stderr.split("\n"):
                if l.startswith("timing:"):
                    f = l.split()
                    stages[" ".join(f[1:-1])] = round(float(f[-1]), 1)
            trace = [l for l in r.stderr.split("\n") if l.startswith("[ahs trace]")]
            out = {"workload": a.workload, "scale": a.scale, "host": a.host, "chains": counts["chains"], "gaf_lines": sum(1 for _ in open(os.path.join(td, "s.gaf"))),
                   "gaf_mb": round(os.path.getsize(os.path.join(td, "s.gaf")) / 1e6, 1), "gfa_mb": round(os.path.getsize(os.path.join(td, "s.gfa")) / 1e6, 1),
                   "stages_ms": stages, "wall_s": round(wall, 2), "host_cores": os.cpu_count()}
            if trace:
                out["ahs_trace"] = trace[-1]
            if os.environ.get("CLI_E2E_VERBOSE"):
                print(json.dumps(out), file=sys.stderr)
        print(json.dumps(out))


if __name__ == "__main__":
    main()
