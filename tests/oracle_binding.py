"""ctypes binding of oracle/_ref/liboracle.so (the CPU restatement).  TEST INFRASTRUCTURE:
imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference."""
import ctypes as C
import os

import numpy as np

from ahsoka_b200.api import ROOT, Batch, BatchIn, BatchOut, PhaseResult, result_from_struct

ORACLE_LIB = os.path.join(ROOT, "oracle", "_ref", "liboracle.so")
_lib = None


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(ORACLE_LIB)
        lib.ahs_oracle_phase_batch.argtypes = [C.POINTER(BatchIn), C.POINTER(BatchOut), C.c_int]
        lib.ahs_oracle_phase_batch.restype = C.c_int
        lib.ahs_oracle_free_out.argtypes = [C.POINTER(BatchOut)]
        i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        lib.ahs_oracle_score.argtypes = [C.c_int, i64p, i32p, i32p, C.c_int, C.c_int64] + [i32p] * 7
        lib.ahs_oracle_score.restype = C.c_int64
        lib.ahs_oracle_cluster.argtypes = [C.c_int, C.c_int64, i32p, i32p, i32p, C.c_int, i32p]
        lib.ahs_oracle_cluster.restype = C.c_int
        lib.ahs_oracle_log_tables.argtypes = [i64p, i64p]
        lib.ahs_oracle_std_sort.argtypes = [i32p, i32p, C.c_int32, C.c_int]
        lib.ahs_oracle_std_sort.restype = None
        lib.ahs_oracle_antiqsort.argtypes = [C.c_int32, i32p]
        lib.ahs_oracle_antiqsort.restype = None
        _lib = lib
    return _lib


def oracle_phase(batch: Batch, n_threads: int = 0) -> PhaseResult:
    lib = load()
    s, o = batch.c_struct(), BatchOut()
    rc = lib.ahs_oracle_phase_batch(C.byref(s), C.byref(o), n_threads)
    if rc != 0:
        raise RuntimeError(f"oracle failed ({rc})")
    try:
        return result_from_struct(o)
    finally:
        lib.ahs_oracle_free_out(C.byref(o))


def oracle_score(rows, ploidy):
    """rows: list of (positions, alleles).  Returns dict with pair arrays and per-read rates."""
    lib = load()
    n = len(rows)
    off = np.zeros(n + 1, dtype=np.int64)
    for i, (p, _) in enumerate(rows):
        off[i + 1] = off[i] + len(p)
    pos = np.concatenate([np.asarray(p, dtype=np.int32) for p, _ in rows]) if n else np.zeros(0, np.int32)
    al = np.concatenate([np.asarray(a, dtype=np.int32) for _, a in rows]) if n else np.zeros(0, np.int32)
    cap = n * (n - 1) // 2 + 1
    bufs = [np.zeros(cap, dtype=np.int32) for _ in range(5)]
    es, ed = np.zeros(max(n, 1), dtype=np.int32), np.zeros(max(n, 1), dtype=np.int32)
    i32p, i64p = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    m = lib.ahs_oracle_score(n, off.ctypes.data_as(i64p), pos.ctypes.data_as(i32p), al.ctypes.data_as(i32p), ploidy, cap,
                             *[b.ctypes.data_as(i32p) for b in bufs], es.ctypes.data_as(i32p), ed.ctypes.data_as(i32p))
    return {"i": bufs[0][:m], "j": bufs[1][:m], "n": bufs[2][:m], "k": bufs[3][:m], "w": bufs[4][:m], "es": es[:n], "ed": ed[:n]}


def oracle_cluster(n, pi, pj, pw, paranoid=False):
    lib = load()
    pi, pj, pw = (np.ascontiguousarray(x, dtype=np.int32) for x in (pi, pj, pw))
    label = np.full(max(n, 1), -1, dtype=np.int32)
    i32p = C.POINTER(C.c_int32)
    k = lib.ahs_oracle_cluster(n, len(pi), pi.ctypes.data_as(i32p), pj.ctypes.data_as(i32p), pw.ctypes.data_as(i32p), int(paranoid),
                               label.ctypes.data_as(i32p))
    return k, label[:n]


def oracle_std_sort(keys, values, descending=False):
    """libstdc++ std::sort on (key, value) pairs, comparator on the key only.  Returns sorted copies."""
    lib = load()
    k = np.ascontiguousarray(keys, dtype=np.int32).copy(); v = np.ascontiguousarray(values, dtype=np.int32).copy()
    i32p = C.POINTER(C.c_int32)
    lib.ahs_oracle_std_sort(k.ctypes.data_as(i32p), v.ctypes.data_as(i32p), len(k), int(descending))
    return k, v


def oracle_antiqsort(n):
    """Keys that drive libstdc++'s introsort into its heap-sort fall-back (McIlroy's adversary)."""
    lib = load()
    k = np.zeros(max(n, 1), dtype=np.int32)
    lib.ahs_oracle_antiqsort(n, k.ctypes.data_as(C.POINTER(C.c_int32)))
    return k[:n]
