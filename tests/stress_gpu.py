"""Randomised parity stress: many seeded batches of varied shape (chain length, depth, ploidy, alleles, error rate)
through the C ABI against the CPU oracle, each phased twice (determinism).  usage: python tests/stress_gpu.py [n_cases] [seed]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ahsoka_b200 import api, synth  # noqa: E402
from tests.oracle_binding import oracle_phase  # noqa: E402


def main():
    n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 12345)
    bad_cases = 0
    t0 = time.time()
    for it in range(n_cases):
        ploidy = int(rng.choice([2, 2, 2, 2, 3, 4]))
        mean_len = int(rng.choice([3, 8, 20, 40, 70, 120]))
        depth = float(rng.choice([3.0, 10.0, 20.0, 30.0, 45.0])) * (ploidy / 2 if ploidy > 2 else 1)
        n_chains = int(rng.integers(4, 60))
        prm = synth.params(ploidy, n_chains, int(rng.choice([0, 1])), mean_len, min_len=int(rng.choice([1, 2])), depth=depth,
                           mean_span=float(rng.choice([6.0, 16.0, 30.0])), err=float(rng.choice([0.0, 0.05, 0.15])),
                           miss=float(rng.choice([0.0, 0.02, 0.2])), max_alleles=int(rng.choice([2, 2, 3, 6, 12])) if ploidy == 2 else None,
                           dup_lines=int(rng.choice([0, 0, 50])), seed=int(rng.integers(1, 2 ** 31)))
        b = synth.generate(prm)
        want = oracle_phase(b, os.cpu_count() or 1)
        g1 = api.phase_batch(b)
        g2 = api.phase_batch(b)
        d1, d2 = g1.diff(want), g2.diff(want)
        tag = "ok" if not d1 and not d2 else f"DIFF run1={d1} run2={d2}"
        if d1 or d2:
            bad_cases += 1
        nmax = int(np.diff(want.read_off).max()) if want.read_off.size > 1 else 0
        print(f"case {it:3d}: p={ploidy} chains={n_chains:3d} len~{mean_len:3d} depth={depth:5.1f} cells={want.n_cells:7d} max reads/chain={nmax:4d} {tag}", flush=True)
    print(f"{n_cases - bad_cases}/{n_cases} cases bit-identical to the oracle in both runs ({time.time() - t0:.0f} s)")
    return 1 if bad_cases else 0


if __name__ == "__main__":
    sys.exit(main())
