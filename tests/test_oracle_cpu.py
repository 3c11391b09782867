"""CPU tests of the oracle itself: hand-derived known answers for the three restated algorithms
(the reference has no golden vectors) and committed golden fixtures."""
import json
import os

import numpy as np
import pytest

from ahsoka_b200 import synth
from ahsoka_b200.api import Batch
from tests.oracle_binding import load, oracle_cluster, oracle_phase, oracle_score

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_log_tables_known_values():
    import ctypes as C
    ln = np.zeros(1025, dtype=np.int64); ln1 = np.zeros(1025, dtype=np.int64)
    p = C.POINTER(C.c_int64)
    load().ahs_oracle_log_tables(ln.ctypes.data_as(p), ln1.ctypes.data_as(p))
    assert ln[1024] == 0 and ln1[0] == 0
    assert ln[512] == round(np.log(0.5) * 2 ** 20) == -726817
    assert ln1[512] == -726817
    assert np.all(np.diff(ln[1:]) > 0) and np.all(np.diff(ln1[:1024]) < 0)


def test_score_overlap_and_disagreement_counts():
    # read0: pos 0..3 alleles 0000 ; read1: pos 2..5 alleles 0 1 . . ; read2: pos 10..11 (no overlap)
    rows = [([0, 1, 2, 3], [0, 0, 0, 0]), ([2, 3, 4, 5], [0, 1, 0, 0]), ([10, 11], [1, 1])]
    r = oracle_score(rows, 2)
    assert list(zip(r["i"], r["j"], r["n"], r["k"])) == [(0, 1, 2, 1)]
    # single partner: cut = max(1, 1//2) = 1, same = (k=1,n=2) -> rq = (1024+1)//2 = 512, diff set empty -> ed = es
    assert r["es"][0] == 512 and r["ed"][0] == 512 and r["es"][2] == 0


def test_score_weight_formula_hand_computed():
    # 4 reads on positions 0..7; reads 0,1 identical haplotype A, reads 2,3 haplotype B (all differ from A)
    A = [0] * 8; B = [1] * 8
    rows = [(list(range(8)), A), (list(range(8)), A), (list(range(8)), B), (list(range(8)), B)]
    r = oracle_score(rows, 2)
    # every read: partners rates {0/8, 8/8, 8/8}; cut = max(1, 3//2) = 1 -> es = 0, ed = rq(16,16) = 1024
    assert set(r["es"]) == {0} and set(r["ed"]) == {1024}
    # pair rates clamp: es -> 10, ed -> 972
    ln = lambda x: round(np.log(x / 1024) * 2 ** 20); ln1 = lambda x: round(np.log(1 - x / 1024) * 2 ** 20)
    same = (8 * (ln1(10) - ln1(972))) >> 10
    diff = (8 * (ln(10) - ln(972))) >> 10
    w = {(int(i), int(j)): int(v) for i, j, v in zip(r["i"], r["j"], r["w"])}
    assert w[(0, 1)] == same > 0 and w[(2, 3)] == same
    assert w[(0, 2)] == diff < 0 and w[(1, 3)] == diff


def test_cluster_edit_two_cliques():
    # two positive triangles joined by negative edges
    pi, pj, pw = [], [], []
    for a in range(6):
        for b in range(a + 1, 6):
            pi.append(a); pj.append(b); pw.append(1000 if (a < 3) == (b < 3) else -800)
    k, label = oracle_cluster(6, pi, pj, pw, paranoid=True)
    assert k == 2 and list(label) == [0, 0, 0, 1, 1, 1]


def test_cluster_edit_singletons_and_zero_weights():
    k, label = oracle_cluster(4, [0, 1], [1, 2], [0, -5], paranoid=True)
    assert k == 4 and list(label) == [0, 1, 2, 3]
    k, label = oracle_cluster(0, [], [], [])
    assert k == 0


def test_cluster_edit_tie_break_smallest_pair():
    # path 0-1-2 with equal positive weights and a weak negative 0-2:
    # icf(0,1) = 10, icf(1,2) = 10 (no common positive neighbour) -> tie -> (0,1) merges first, then 2 joins (10 - 3 > 0)
    k, label = oracle_cluster(3, [0, 1, 0], [1, 2, 2], [10, 10, -3], paranoid=True)
    assert k == 1
    # with a strong negative the merged weight to node 2 is 10 - 30 < 0 -> stays apart
    k, label = oracle_cluster(3, [0, 1, 0], [1, 2, 2], [10, 10, -30], paranoid=True)
    assert k == 2 and list(label) == [0, 0, 1]


@pytest.mark.parametrize("seed", range(6))
def test_cluster_edit_incremental_equals_definition(seed):
    # paranoid mode re-derives every induced cost from the definition after every step and aborts on a mismatch
    rng = np.random.default_rng(seed)
    n = 24
    pi, pj, pw = [], [], []
    for a in range(n):
        for b in range(a + 1, min(n, a + 8)):
            if rng.random() < 0.8:
                pi.append(a); pj.append(b); pw.append(int(rng.integers(-4000, 4000)))
    k, label = oracle_cluster(n, pi, pj, pw, paranoid=True)
    assert 1 <= k <= n and label.min() == 0 and label.max() == k - 1


def test_phase_recovers_clean_diploid_truth():
    # error-free reads: two clusters per chain, haplotypes complementary at every covered position
    b = synth.generate(synth.params(2, 8, 0, 30, depth=30.0, err=0.0, miss=0.0, seed=3))
    r = oracle_phase(b)
    assert (r.status == 0).all() and (r.n_clusters == 2).all()
    assert (r.dp_cost == np.round(r.dp_cost)).all()          # integer-valued costs
    hap = r.hap_allele.reshape(-1, 2)
    assert ((hap[:, 0].astype(int) + hap[:, 1]) == 1).mean() > 0.95
    truth = b.truth["hap_allele"].reshape(-1, 2)
    for c in range(b.n_chains):
        pos = r.pos[r.pos_off[c]:r.pos_off[c + 1]]
        t = truth[b.bubble_off[c] + pos]
        h = hap[r.pos_off[c]:r.pos_off[c + 1]]
        agree = max((h == t).all(axis=1).mean(), (h == t[:, ::-1]).all(axis=1).mean())
        assert agree >= 0.9, (c, agree)


def test_status_codes_trivial_and_empty():
    b = synth.generate(synth.params(2, 60, 1, 3, min_len=1, depth=12.0, seed=11))
    r = oracle_phase(b)
    nb = np.diff(b.bubble_off)
    assert ((r.status == 1) == (nb <= 1)).all()
    assert (r.status == 2).any() and (r.status == 0).any()
    for c in np.nonzero(r.status != 0)[0]:
        assert r.read_off[c] == r.read_off[c + 1] and r.pos_off[c] == r.pos_off[c + 1]


def test_filter_mapq_threshold_float32():
    # int(float32(id) * 100) >= 93 (SURVEY A#4): 0.93 -> 93 kept, 0.9299 -> 92 dropped
    b = synth.generate(synth.params(2, 1, 0, 12, depth=30.0, err=0.0, miss=0.0, seed=9))
    r0 = oracle_phase(b)
    kept = set(r0.read_id)
    rid = int(r0.read_id[0])
    ident = b.entry_identity.copy()
    ident[b.entry_read == rid] = np.float32(0.9299)
    b2 = Batch(b.ploidy, b.chain_id, b.bubble_off, b.allele_off, b.anode_off, b.anode, b.stage_a_order, b.read_off, b.entry_off,
               b.enode_off, b.enode, b.entry_read, ident)
    assert rid not in set(oracle_phase(b2).read_id)
    ident[b.entry_read == rid] = np.float32(0.93)
    b3 = Batch(b.ploidy, b.chain_id, b.bubble_off, b.allele_off, b.anode_off, b.anode, b.stage_a_order, b.read_off, b.entry_off,
               b.enode_off, b.enode, b.entry_read, ident)
    r3 = oracle_phase(b3)
    assert rid in set(r3.read_id) and r3.read_mapq[list(r3.read_id).index(rid)] == 93
    assert kept


def test_deletion_allele_matches_every_entry():
    # a 2-node allele path has no inner node: stage B matches it for EVERY entry of the chain (SURVEY A#9)
    # chain: bubbles 0,1,2 simple; bubble 1 gets alleles [s,i0,t], [t,s] (deletion)
    anode = [1, 2, 4, 1, 3, 4,   4, 5, 7, 7, 4,   7, 8, 10, 7, 9, 10]
    anode_off = [0, 3, 6, 9, 11, 14, 17]
    allele_off = [0, 2, 4, 6]
    reads = [[1, 2, 4, 5, 7, 8, 10], [1, 3, 4, 7, 9, 10], [1, 2, 4], [7, 8, 10, 4, 5], [1, 3, 4, 5, 7, 9, 10]]
    enode = sum(reads, []); enode_off = np.cumsum([0] + [len(x) for x in reads])
    b = Batch(2, [0], [0, 3], allele_off, anode_off, anode, [2, 1, 0], [0, len(reads)], [0, len(reads)], enode_off, enode,
              list(range(len(reads))), [0.99] * len(reads))
    r = oracle_phase(b)
    assert r.status[0] == 0
    cells = {(int(r.read_id[i]), int(p)): int(a) for i in range(len(r.read_id))
             for p, a in zip(r.cell_pos[r.cell_off[i]:r.cell_off[i + 1]], r.cell_allele[r.cell_off[i]:r.cell_off[i + 1]])}
    assert cells[(0, 1)] == 0            # real inner node 5 present -> allele 0 wins (first matching allele)
    assert cells[(1, 1)] == 1            # no inner node of allele 0 -> the deletion allele matches anyway
    assert cells[(2, 1)] == 1            # read that does not even reach bubble 1


def test_golden_fixtures():
    """Outputs of the oracle on committed inputs; regenerated by tests/golden/make_golden.py."""
    idx = json.load(open(os.path.join(GOLDEN, "index.json")))
    assert idx["cases"]
    for case in idx["cases"]:
        b = Batch.load(os.path.join(GOLDEN, case["batch"]))
        want = np.load(os.path.join(GOLDEN, case["result"]))
        got = oracle_phase(b)
        for k in got.ARRAYS:
            assert np.array_equal(getattr(got, k), want[k]), (case["name"], k)


def test_radix_selection_of_the_scoring_kernel_equals_rule_r1_by_sorting(tmp_path):
    # the product's own statements (ahsoka_b200/csrc/k_select.cuh, host-compilable) against a sort-based restatement
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "select_harness")
    subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(root, "tests", "native", "select_harness.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe, "60000"], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout


@pytest.mark.parametrize("ploidy", [5, 6])
def test_rule_r3c_definition_equals_sub_multiset_dp(ploidy):
    # canonical-tuple threading (ploidy > 4): the all-pairs definition against the sub-multiset DP the parity tests use at size
    import ctypes as C
    lib = load()
    lib.ahs_oracle_canonical_selfcheck.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.POINTER(C.c_double)]
    for seed in range(12):
        cost = C.c_double()
        assert lib.ahs_oracle_canonical_selfcheck(ploidy, 10, 9, seed, C.byref(cost)) == 0
        assert cost.value > 0
