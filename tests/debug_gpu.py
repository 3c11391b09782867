"""Diagnostic: run several seeded cases on the GPU and report where the result leaves the oracle."""
import os
import sys
import time
import traceback

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ahsoka_b200 import api, synth  # noqa: E402
from tests.oracle_binding import oracle_phase  # noqa: E402

CASES = {
    "dip_small": synth.params(2, 40, 1, 24, depth=30.0, seed=1),
    "dip_tiny": synth.params(2, 60, 1, 3, min_len=1, depth=12.0, seed=11),
    "dip_lowdepth": synth.params(2, 50, 1, 10, depth=2.0, seed=5),
    "multiline": synth.params(2, 30, 1, 30, depth=25.0, dup_lines=300, seed=21),
    "trip": synth.params(3, 12, 1, 30, depth=30.0, seed=31),
    "tetra": synth.params(4, 12, 1, 30, depth=40.0, seed=41),
    "fourbit": synth.params(2, 20, 1, 20, depth=30.0, max_alleles=6, seed=51),
    "dip_mid": synth.params(2, 30, 1, 80, depth=30.0, seed=71),
    "dip_big": synth.params(2, 12, 1, 140, depth=30.0, seed=72),
    "long_lowdepth": synth.params(2, 8, 1, 400, depth=4.0, seed=73),
    "cfg2_small": synth.config("cfg2", 0.01),
    "cfg1": synth.config("cfg1"),
    "cfg2_quarter": synth.config("cfg2", 0.25),
}


def first_diff(a, b):
    if a.shape != b.shape:
        return f"shape {a.shape} vs {b.shape}"
    idx = np.nonzero(a != b)[0]
    return f"{len(idx)} of {len(a)} differ; first at {idx[0]}: got {a[idx[0]]} want {b[idx[0]]}" if len(idx) else "equal"


def main():
    names = sys.argv[1:] or [k for k in CASES if k != "cfg2_quarter"]
    ok = True
    for name in names:
        batch = synth.generate(CASES[name])
        t0 = time.time()
        want = oracle_phase(batch)
        t1 = time.time()
        try:
            got = api.phase_batch(batch)
        except Exception:
            traceback.print_exc()
            ok = False
            continue
        t2 = time.time()
        bad = got.diff(want)
        print(f"== {name}: chains {batch.n_chains} ok {want.n_chains_ok} cells {want.n_cells} pairs {want.n_pairs} "
              f"oracle {t1 - t0:.2f}s gpu-call {t2 - t1:.3f}s  {'MATCH' if not bad else 'DIFF ' + str(bad)}")
        print("   timings", {k: round(v, 3) for k, v in got.timings.items()})
        for k in bad:
            a, b = getattr(got, k), getattr(want, k)
            if isinstance(a, np.ndarray):
                print(f"   {k}: {first_diff(a, b)}")
            else:
                print(f"   {k}: got {a} want {b}")
        if bad:
            ok = False
            # locate the first chain where something differs
            for c in range(batch.n_chains):
                if got.status[c] != want.status[c]:
                    print(f"   chain {c}: status got {got.status[c]} want {want.status[c]}")
                    break
            for k in ("read_off", "pos_off", "n_clusters", "dp_cost", "maxpos"):
                a, b = getattr(got, k), getattr(want, k)
                if a.shape == b.shape and (a != b).any():
                    idx = np.nonzero(a != b)[0]
                    print(f"   {k}: first differing chain {idx[0]}, last {idx[-1]}, n={len(idx)}")
    print("ALL MATCH" if ok else "MISMATCHES")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
