"""Regenerates the golden fixtures: small seeded batches + the CPU oracle's result on them, and
(when the reference-verbatim binary oracle/_ref/Ahsoka_ref exists) the -result.txt that binary
writes for the GFA/GAF form of the same instance.  Run from the repo root:
    python tests/golden/make_golden.py
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ahsoka_b200 import synth  # noqa: E402
from tests.oracle_binding import oracle_phase  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [
    ("dip_a", synth.params(2, 6, 1, 12, depth=20.0, seed=101), True),
    ("dip_multiline", synth.params(2, 4, 1, 14, depth=20.0, dup_lines=300, seed=102), True),
    ("dip_tiny", synth.params(2, 10, 1, 3, min_len=1, depth=12.0, seed=103), True),
    ("trip", synth.params(3, 3, 1, 12, depth=30.0, seed=104), False),
    ("tetra", synth.params(4, 2, 1, 10, depth=40.0, seed=105), False),
    ("penta", synth.params(5, 2, 1, 9, depth=50.0, seed=106), False),        # rule R3c: canonical tuples over the first p + 2 clusters
    ("hexa", synth.params(6, 2, 1, 8, depth=60.0, max_alleles=9, seed=107), False),
    ("dip_200_reads", synth.params(2, 1, 0, 120, depth=40.0, seed=108), False),   # one chain above 160 final reads: the HBM-resident scoring / cluster-editing route
]
ONLY = set(sys.argv[1:])          # names to (re)generate; none = all


def main():
    index = {"cases": []}
    old = {}
    if ONLY and os.path.exists(os.path.join(HERE, "index.json")):
        old = {c["name"]: c for c in json.load(open(os.path.join(HERE, "index.json")))["cases"]}
    ref = os.path.join(ROOT, "oracle", "_ref", "Ahsoka_ref")
    for name, prm, cli in CASES:
        if ONLY and name not in ONLY and name in old:
            index["cases"].append(old[name])
            continue
        with tempfile.TemporaryDirectory() as td:
            b = synth.generate(prm, os.path.join(td, name))
            b.save(os.path.join(HERE, name + ".batch.npz"))
            r = oracle_phase(b)
            np.savez_compressed(os.path.join(HERE, name + ".result.npz"), **{k: getattr(r, k) for k in r.ARRAYS})
            entry = {"name": name, "batch": name + ".batch.npz", "result": name + ".result.npz"}
            if cli and os.path.exists(ref):
                for ext in (".gfa", ".gaf"):
                    open(os.path.join(HERE, name + ext), "w").write(open(os.path.join(td, name + ext)).read())
                subprocess.run([ref, "phase", "-g", name + ".gfa", "-a", name + ".gaf", "-o", "out"], cwd=td, check=True,
                               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
                open(os.path.join(HERE, name + ".ref-result.txt"), "w").write(open(os.path.join(td, "out-result.txt")).read())
                entry.update({"gfa": name + ".gfa", "gaf": name + ".gaf", "ref_result": name + ".ref-result.txt"})
            index["cases"].append(entry)
    json.dump(index, open(os.path.join(HERE, "index.json"), "w"), indent=1)
    print("wrote", len(index["cases"]), "cases")


if __name__ == "__main__":
    main()
