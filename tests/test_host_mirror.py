"""Host-side mirror of the reference's PairHMM type (lorikeet_b200/engine.py): the argument checks, messages and the
haplotype-prefix bookkeeping of pair_hmm.rs:63-207,404-500 — everything that runs before / after the GPU call."""
import numpy as np
import pytest

from lorikeet_b200.engine import PairHMM


class _FakeCtx:
    """stands in for the device so the bookkeeping can be checked without one (returns a fixed log10 likelihood)"""

    def __init__(self, value=-1.5):
        self.value = value
        self.calls = 0

    def pairhmm_forward(self, *a):
        self.calls += 1
        return self.value


def test_initialize_keeps_list_order_and_sizes():
    haps = [b"ACGTACGT", b"ACGTTCGTA", b"ACG"]
    hmm = PairHMM.initialize(None, haps, {0: [b"ACGT", b"ACGTAC"], 1: [b"AC"]})
    assert hmm.initialized and hmm.max_read_length == 6 and hmm.max_haplotype_length == 9
    assert hmm.m_haplotype_data_array == haps
    assert [hmm.haplotype_to_haplotype_list_index_map[h] for h in haps] == [0, 1, 2]


def test_quick_initialize_with_zero_sizes_is_not_initialized():  # pair_hmm.rs:138-141
    hmm = PairHMM.quick_initialize(None, 0, 0)
    assert not hmm.initialized
    with pytest.raises(AssertionError, match="Must call initialize"):
        hmm.compute_read_likelihood_given_haplotype_log10(b"A", b"A", [30], [45], [45], [10])
    hmm.reinitialize(10, 10)
    assert hmm.initialized


def test_argument_asserts_carry_the_reference_messages():
    hmm = PairHMM.quick_initialize(_FakeCtx(), 10, 4)
    q = [30] * 4
    with pytest.raises(AssertionError, match="Haplotype bases is too long"):
        hmm.compute_read_likelihood_given_haplotype_log10(b"ACGTA", b"ACGT", q, q, q, q)
    with pytest.raises(AssertionError, match="read quals aren't the same size"):
        hmm.compute_read_likelihood_given_haplotype_log10(b"ACGT", b"ACGT", q[:3], q, q, q)
    with pytest.raises(AssertionError, match="insertion gcp aren't the same size"):
        hmm.compute_read_likelihood_given_haplotype_log10(b"ACGT", b"ACGT", q, q[:3], q, q)
    with pytest.raises(AssertionError, match="deletion gcp aren't the same size"):
        hmm.compute_read_likelihood_given_haplotype_log10(b"ACGT", b"ACGT", q, q, q[:3], q)
    with pytest.raises(AssertionError, match="overal GCP aren't the same size"):
        hmm.compute_read_likelihood_given_haplotype_log10(b"ACGT", b"ACGT", q, q, q, q[:3])
    hmm_pos = PairHMM.quick_initialize(_FakeCtx(0.25), 10, 4)
    with pytest.raises(AssertionError, match="cannot be greater than 0.0"):
        hmm_pos.compute_read_likelihood_given_haplotype_log10(b"ACGT", b"ACGT", q, q, q, q)


def test_hap_start_index_bookkeeping_follows_the_reference():  # pair_hmm.rs:442-496
    ctx = _FakeCtx()
    hmm = PairHMM.quick_initialize(ctx, 10, 8)
    q = [30] * 4
    h1, h2, h3 = b"ACGTACGT", b"ACGTAGGT", b"ACCTAGGT"
    assert hmm.compute_read_likelihood_given_haplotype_log10(h1, b"ACGT", q, q, q, q, True, h2) == -1.5
    assert hmm.hap_start_index == 5 and hmm.previous_haplotype_length == 8
    hmm.compute_read_likelihood_given_haplotype_log10(h2, b"ACGT", q, q, q, q, False, h3)
    assert hmm.hap_start_index == 0  # next difference (2) lies before the current start (5): reset
    hmm.compute_read_likelihood_given_haplotype_log10(h3, b"ACGT", q, q, q, q, False, b"ACC")
    assert hmm.hap_start_index == 0  # different length: no shared prefix is assumed
    hmm.compute_read_likelihood_given_haplotype_log10(h3, b"ACGT", q, q, q, q, False, None)
    assert hmm.hap_start_index == 0 and ctx.calls == 4
    assert PairHMM.find_first_position_where_haplotypes_differ(h1, h1) == 8
    hmm.set_do_exact_log10(True); hmm.set_logless(False)  # inert flags (pair_hmm.rs:172-184)
    assert hmm.do_exact_log10 and not hmm.logless
