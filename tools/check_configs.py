import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ahsoka_b200 import api, synth
from tests.oracle_binding import oracle_phase
for name, scale in (("cfg4", 0.004), ("cfg3", 0.01), ("cfg1", 1.0)):
    b = synth.generate(synth.config(name, scale))
    want = oracle_phase(b, os.cpu_count() or 1)
    got = api.phase_batch(b)
    print(name, scale, "chains", b.n_chains, "cells", want.n_cells, "diff:", got.diff(want) or "none", flush=True)
