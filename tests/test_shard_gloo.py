"""N>1 path on CPU: two gloo ranks shard the chains of one batch (LPT, no data-path collective), each phases
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


def _worker(rank, world, initfile, outdir):
    dist.init_process_group("gloo", init_method=f"file://{initfile}", rank=rank, world_size=world)
    b = synth.generate(_params())
    parts = shard.lpt_partition(b, world)
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
