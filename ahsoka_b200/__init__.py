"""ahsoka_b200 — B200-native (sm_100a CUDA) implementation of Ahsoka's per-chain phasing hot path.

The product is the C-ABI shared library ``ahsoka_b200/lib/libahsoka_b200.so`` declared in
``include/ahsoka_b200.h`` (drop-in for reference ``src/alignmentstoreadset.cpp:55``).  This
Python package is only the thin host mirror used by tests and ``bench.py``: ctypes bindings
(:mod:`ahsoka_b200.api`), the synthetic workload generator (:mod:`ahsoka_b200.synth`) and the
chain sharder for one-process-per-GPU runs (:mod:`ahsoka_b200.shard`).
"""
from .api import Batch, PhaseResult, load_library, phase_batch, LIB_PATH  # noqa: F401
