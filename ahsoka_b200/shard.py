"""Chain sharding for the one-process-per-GPU launch.

Bubble chains are independent (reference src/alignmentstoreadset.cpp:75), so N ranks phase N
disjoint sets of chains with no data-path collective; the host gathers the per-chain results in
input order.

`contiguous_partition` is the library's own rule (ahs_phase_batch_multi, through `ahs_plan_shares`): one contiguous
share of the chains per part — they arrive largest first (polyassembly.cpp:135-140) — with the largest summed
`ahs_chain_cost` as small as contiguous cuts allow; the long chains of a part stay together and cluster side by side.
`lpt_partition` (chains by decreasing cost, each onto the least loaded part) balances arbitrary orders better and is
kept for host tools that re-pack their parts anyway."""
from __future__ import annotations

import numpy as np

from .api import Batch, PhaseResult, load_library


def chain_costs(batch: Batch) -> np.ndarray:
    lib = load_library()
    nb = np.diff(batch.bubble_off)
    ne = np.diff(batch.entry_off)
    nen = batch.enode_off[batch.entry_off[1:]] - batch.enode_off[batch.entry_off[:-1]]
    return np.array([lib.ahs_chain_cost(int(b), int(e), int(n), int(batch.ploidy)) for b, e, n in zip(nb, ne, nen)], dtype=np.float64)


def plan_shares(cost: np.ndarray, n_parts: int) -> np.ndarray:
    """cuts[0..n_parts] of the library's share rule for the given chain costs (host-only call)."""
    import ctypes as C
    lib = load_library()
    cost = np.ascontiguousarray(cost, dtype=np.float64)
    cuts = np.zeros(n_parts + 1, dtype=np.int64)
    rc = lib.ahs_plan_shares(cost.ctypes.data_as(C.POINTER(C.c_double)), int(cost.shape[0]), int(n_parts),
                             cuts.ctypes.data_as(C.POINTER(C.c_int64)))
    if rc != 0:
        raise ValueError(lib.ahs_last_error().decode())
    return cuts


def contiguous_partition(batch: Batch, n_parts: int) -> list[np.ndarray]:
    """Chain indices of each part: part g = chains [cuts[g], cuts[g+1]) of `plan_shares` (a part may be empty)."""
    cuts = plan_shares(chain_costs(batch), n_parts)
    return [np.arange(cuts[g], cuts[g + 1], dtype=np.int64) for g in range(n_parts)]


def lpt_partition(batch: Batch, n_parts: int) -> list[np.ndarray]:
    """Chain indices of each part, ascending within a part (every chain in exactly one part)."""
    cost = chain_costs(batch)
    order = np.argsort(-cost, kind="stable")
    load = np.zeros(n_parts)
    parts: list[list[int]] = [[] for _ in range(n_parts)]
    for c in order:
        g = int(np.argmin(load))
        load[g] += cost[c]
        parts[g].append(int(c))
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def gather(parts: list[np.ndarray], results: list[PhaseResult], n_chains: int) -> PhaseResult:
    """Host gather of per-part results back into input chain order."""
    where = {}
    for g, idx in enumerate(parts):
        for i, c in enumerate(idx):
            where[int(c)] = (g, i)
    p = results[0].ploidy
    status, n_clusters, dp_cost, maxpos = [], [], [], []
    read_id, read_mapq, read_cluster, cell_pos, cell_allele, pos, path, hap = [], [], [], [], [], [], [], []
    read_off, cell_off, pos_off = [0], [0], [0]
    for c in range(n_chains):
        g, i = where[c]
        r = results[g]
        status.append(r.status[i]); n_clusters.append(r.n_clusters[i]); dp_cost.append(r.dp_cost[i]); maxpos.append(r.maxpos[i])
        r0, r1 = int(r.read_off[i]), int(r.read_off[i + 1])
        read_id.append(r.read_id[r0:r1]); read_mapq.append(r.read_mapq[r0:r1]); read_cluster.append(r.read_cluster[r0:r1])
        c0, c1 = int(r.cell_off[r0]), int(r.cell_off[r1])
        cell_pos.append(r.cell_pos[c0:c1]); cell_allele.append(r.cell_allele[c0:c1])
        base = cell_off[-1] - c0
        cell_off.extend((r.cell_off[r0 + 1:r1 + 1] + base).tolist())
        read_off.append(read_off[-1] + (r1 - r0))
        q0, q1 = int(r.pos_off[i]), int(r.pos_off[i + 1])
        pos.append(r.pos[q0:q1]); path.append(r.path[q0 * p:q1 * p]); hap.append(r.hap_allele[q0 * p:q1 * p])
        pos_off.append(pos_off[-1] + (q1 - q0))

    def cat(xs, dt):
        return np.concatenate(xs).astype(dt) if xs else np.zeros(0, dtype=dt)
    return PhaseResult(p, np.array(status, np.int32), np.array(read_off, np.int64), cat(read_id, np.int32), cat(read_mapq, np.int32),
                       cat(read_cluster, np.int32), np.array(cell_off, np.int64), cat(cell_pos, np.int32), cat(cell_allele, np.uint8),
                       np.array(n_clusters, np.int32), np.array(pos_off, np.int64), cat(pos, np.int32), cat(path, np.int32),
                       cat(hap, np.uint8), np.array(dp_cost, np.float64), np.array(maxpos, np.int32),
                       sum(r.n_cells for r in results), sum(r.n_pairs for r in results), sum(r.n_chains_ok for r in results), {})
