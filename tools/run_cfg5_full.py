"""One pass over a cfg5-shaped batch WITH full-length chains (10,000 bubbles, ~35,000 final reads at 80x, ploidy 6): too big
for the CPU oracle, so the result is checked against the generator's truth (haplotype agreement up to relabelling) and the
chain-level invariants.  usage: python tools/run_cfg5_full.py [scale]"""
import itertools
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ahsoka_b200 import api, synth


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.02
    prm = synth.config("cfg5", scale)
    b = synth.generate(prm)
    p = int(b.ploidy)
    t0 = time.perf_counter(); r = api.phase_batch(b); dt = time.perf_counter() - t0
    nb = np.diff(b.bubble_off); nr = np.diff(r.read_off)
    out = {"chains": b.n_chains, "ploidy": p, "largest_chain_bubbles": int(nb.max()), "largest_chain_final_reads": int(nr.max()),
           "status_counts": {int(k): int(v) for k, v in zip(*np.unique(r.status, return_counts=True))}, "cells": int(r.n_cells),
           "pairs": int(r.n_pairs), "wall_s": dt, "timings_ms": {k: v for k, v in r.timings.items() if k.startswith("ms_")}}
    # the longest chain against the truth: per position, the multiset of emitted alleles vs the multiset of true alleles
    c = int(np.argmax(nr))
    q0, q1 = int(r.pos_off[c]), int(r.pos_off[c + 1])
    truth = b.truth["hap_allele"].reshape(-1, p)[int(b.bubble_off[c]) + r.pos[q0:q1]]
    hap = r.hap_allele[q0 * p:q1 * p].reshape(-1, p)
    same = int((np.sort(truth, axis=1) == np.sort(hap, axis=1)).all(axis=1).sum())
    out["longest_chain"] = {"positions": q1 - q0, "clusters": int(r.n_clusters[c]), "positions_with_the_true_allele_multiset": same,
                            "dp_cost": float(r.dp_cost[c])}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
