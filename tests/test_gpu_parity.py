"""-m gpu: the CUDA path, through the C ABI, against the CPU oracle on the same seeded inputs.
Bar: bit-exact on every output array (integer / index work; dp_cost is integer valued)."""
import numpy as np
import pytest

from ahsoka_b200 import api, synth
from tests.oracle_binding import oracle_phase

pytestmark = pytest.mark.gpu


def _check(batch, **kw):
    got = api.phase_batch(batch, **kw)
    want = oracle_phase(batch)
    bad = got.diff(want)
    assert not bad, f"GPU differs from oracle in {bad}"
    return got


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_diploid_small_chains(seed):
    got = _check(synth.generate(synth.params(2, 40, 1, 24, depth=30.0, seed=seed)))
    assert got.n_chains_ok > 30


@pytest.mark.parametrize("mean_len,n_chains,seed", [(80, 30, 71), (140, 12, 72)])
def test_diploid_mid_and_big_chains(mean_len, n_chains, seed):
    # ~100 and ~180 final reads per chain: the large shared-memory classes of the fused kernel and,
    # above 166 reads, the HBM-resident scoring / cluster-editing path
    got = _check(synth.generate(synth.params(2, n_chains, 1, mean_len, depth=30.0, seed=seed)))
    assert got.n_chains_ok == n_chains


def test_cfg2_sample():
    got = _check(synth.generate(synth.config("cfg2", 0.01)))
    assert got.n_chains_ok > 400


def test_diploid_tiny_and_trivial_chains():
    # chains of 1..6 bubbles: <=1 bubble is header-only (status 1), very short chains often end empty (status 2)
    _check(synth.generate(synth.params(2, 60, 1, 3, min_len=1, depth=12.0, seed=11)))


def test_diploid_low_depth_empty_chains():
    got = _check(synth.generate(synth.params(2, 50, 1, 10, depth=2.0, seed=5)))
    assert (got.status == 2).any()


def test_cfg1_one_long_chain():
    got = _check(synth.generate(synth.config("cfg1")))
    assert got.n_chains_ok == 1 and got.n_cells > 15000


def test_multiline_reads():
    # same read name on two GAF lines with different identity (SURVEY A#3)
    _check(synth.generate(synth.params(2, 30, 1, 30, depth=25.0, dup_lines=300, seed=21)))


@pytest.mark.parametrize("ploidy,seed", [(3, 31), (4, 41)])
def test_polyploid(ploidy, seed):
    _check(synth.generate(synth.params(ploidy, 12, 1, 30, depth=10.0 * ploidy, seed=seed)))


def test_four_bit_codes_diploid_many_alleles():
    _check(synth.generate(synth.params(2, 20, 1, 20, depth=30.0, max_alleles=6, seed=51)))


def test_resident_mode_same_result():
    b = synth.generate(synth.params(2, 30, 1, 30, depth=30.0, seed=61))
    _check(b, resident_iters=2, warmup=1)


def test_zero_copy_views_and_release():
    b = synth.generate(synth.params(2, 20, 1, 20, depth=30.0, seed=81))
    want = oracle_phase(b)
    got = api.phase_batch(b, copy=False)
    assert not got.diff(want)
    # the output buffers are still owned by the library: a second call on the device must be refused ...
    with pytest.raises(RuntimeError):
        api.phase_batch(b)
    got.release()
    # ... and accepted again after the release
    assert not api.phase_batch(b).diff(want)


def _broken(b, **kw):
    import copy
    c = copy.copy(b)
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def test_malformed_batches_are_rejected_not_mis_phased():
    b = synth.generate(synth.params(2, 8, 1, 12, depth=20.0, seed=91))
    er = b.entry_read.copy(); er[3] = 10 ** 6
    with pytest.raises(RuntimeError, match="entry_read"):
        api.phase_batch(_broken(b, entry_read=er))
    ao = b.allele_off.copy(); ao[5], ao[6] = ao[6] + 1, ao[5]
    with pytest.raises(RuntimeError, match="monotone"):
        api.phase_batch(_broken(b, allele_off=ao))
    so = b.stage_a_order.copy(); so[1] = so[0]
    with pytest.raises(RuntimeError, match="permutation"):
        api.phase_batch(_broken(b, stage_a_order=so))
    eo = b.enode_off.copy(); eo[2] = eo[3] + 5
    with pytest.raises(RuntimeError, match="monotone"):
        api.phase_batch(_broken(b, enode_off=eo))
    # the library is still usable afterwards
    assert not api.phase_batch(b).diff(oracle_phase(b))
