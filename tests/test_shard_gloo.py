"""N>1 path on CPU: two gloo ranks shard the chains of one batch (the library's share rule, no data-path collective), each phases
its part, the host gathers in input order.  The phasing engine here is the CPU oracle (test infrastructure):
what is under test is the sharding / gather logic and the torch.distributed plumbing bench.py uses."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ahsoka_b200 import shard, synth
from tests.oracle_binding import oracle_phase


def _params():
    return synth.params(2, 24, 1, 20, depth=25.0, seed=0x51AD)


def test_lpt_partition_is_a_balanced_partition():
    b = synth.generate(synth.params(2, 200, 2, 60, 2, 400, 1.2, 2, depth=20.0, seed=7))
    for n in (1, 2, 4, 8):
        parts = shard.lpt_partition(b, n)
        allc = np.sort(np.concatenate(parts))
        assert np.array_equal(allc, np.arange(b.n_chains))
        cost = shard.chain_costs(b)
        loads = np.array([cost[p].sum() for p in parts])
        # LPT bound: makespan <= mean + largest item
        assert loads.max() <= loads.mean() + cost.max() + 1e-6


def _optimal_makespan(cost, n_parts):
    # smallest possible largest share over all contiguous partitions into <= n_parts shares (dynamic programme)
    pre = np.concatenate([[0.0], np.cumsum(cost)])
    n = len(cost)
    best = [pre[i] for i in range(n + 1)]                  # one share
    for _ in range(1, n_parts):
        nxt = list(best)
        for i in range(1, n + 1):
            nxt[i] = min(max(best[j], pre[i] - pre[j]) for j in range(i + 1))
        best = nxt
    return best[n]


def test_share_rule_is_the_optimal_contiguous_partition():
    rng = np.random.default_rng(20261019)
    for trial in range(60):
        n = int(rng.integers(1, 40)); g = int(rng.integers(1, 9))
        cost = rng.pareto(1.2, n) + 0.01 if trial % 2 else rng.uniform(0.0, 1.0, n)
        if trial % 5 == 0:
            cost = np.sort(cost)[::-1].copy()              # largest first, as the chains arrive
        cuts = shard.plan_shares(cost, g)
        assert cuts[0] == 0 and cuts[-1] == n and np.all(np.diff(cuts) >= 0)
        loads = np.array([cost[cuts[k]:cuts[k + 1]].sum() for k in range(g)])
        assert loads.max() <= _optimal_makespan(cost, g) * (1 + 1e-5) + 1e-12
    assert list(shard.plan_shares(np.zeros(0), 3)) == [0, 0, 0, 0]
    with pytest.raises(ValueError):
        shard.plan_shares(np.array([1.0, -1.0]), 2)


def test_contiguous_partition_covers_every_chain_once():
    b = synth.generate(synth.params(2, 200, 2, 60, 2, 400, 1.2, 2, depth=20.0, seed=7))
    cost = shard.chain_costs(b)
    for n in (1, 2, 3, 8):
        parts = shard.contiguous_partition(b, n)
        assert np.array_equal(np.concatenate(parts), np.arange(b.n_chains))
        loads = np.array([cost[p].sum() for p in parts])
        assert loads.max() <= max(cost.max(), cost.sum() / n) + cost.max() + 1e-6


def _worker(rank, world, initfile, outdir):
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    b = synth.generate(_params())
    parts = shard.contiguous_partition(b, world)
    mine = oracle_phase(b.select(parts[rank]))
    # the only collectives of the N>1 path: scalar reductions for the report (bench.py)
    s = torch.tensor([mine.n_cells, mine.n_chains_ok], dtype=torch.int64)
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        full = shard.gather(parts, gathered, b.n_chains)
        want = oracle_phase(b)
        bad = full.diff(want)
        with open(os.path.join(outdir, "result.txt"), "w") as f:
            f.write(f"{bad}|{int(s[0])}|{want.n_cells}|{int(s[1])}|{want.n_chains_ok}|{float(t[0])}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sharding_matches_single_process():
    with tempfile.TemporaryDirectory() as td:
        initfile = os.path.join(td, "init")
        mp.spawn(_worker, args=(2, initfile, td), nprocs=2, join=True)
        bad, cells, want_cells, ok, want_ok, tmax = open(os.path.join(td, "result.txt")).read().split("|")
        assert bad == "[]"
        assert cells == want_cells and ok == want_ok
        assert float(tmax) == 2.0
