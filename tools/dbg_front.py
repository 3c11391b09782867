import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
from ahsoka_b200 import api, synth
from ahsoka_b200.api import Batch
from tests.oracle_binding import oracle_phase
anode = [1, 2, 4, 1, 3, 4,   4, 5, 7, 7, 4,   7, 8, 10, 7, 9, 10]
anode_off = [0, 3, 6, 9, 11, 14, 17]
allele_off = [0, 2, 4, 6]
reads = [[1, 2, 4, 5, 7, 8, 10], [1, 3, 4, 7, 9, 10], [1, 2, 4], [7, 8, 10, 4, 5], [1, 3, 4, 5, 7, 9, 10]]
enode = sum(reads, [])
enode_off = np.cumsum([0] + [len(x) for x in reads])
b = Batch(2, [0], [0, 3], allele_off, anode_off, anode, [2, 1, 0], [0, len(reads)], [0, len(reads)], enode_off, enode, list(range(len(reads))), [0.99] * len(reads))
want = oracle_phase(b)
got = api.phase_batch(b)
print("diff", got.diff(want))
for k in got.ARRAYS:
    print(k, getattr(got, k), getattr(want, k))
