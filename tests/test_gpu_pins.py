"""GPU halves (-m gpu) of tests/test_oracle_pins.py: the reference-held vectors through the device code.
  * dangling-tail alignments (tests/read_threading_graph_unit_tests.rs:667-685) through lk_dangling_end_sw_batch, plus random
    pairs against the oracle's LeadingInDel alignment + remove_trailing_deletions;
  * the two AlleleLikelihoods post-op properties (tests/allele_likelihoods_unit_tests.rs:399-446, :725-770) through the
    device epilogue on its own (lk_engine_normalize_and_filter_device)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import lorikeet_b200 as lk
import sw_cases as SC
from conftest import GOLDEN
from lorikeet_b200 import _lib
from oracle import lk_oracle as O
from test_oracle_pins import O_cigar_strip_trailing_deletion, _filter_case

pytestmark = pytest.mark.gpu
TAILS = json.load(open(os.path.join(GOLDEN, "dangling_tail_cases.json")))


@pytest.fixture(scope="module")
def ctx():
    c = lk.Context(0)
    yield c
    c.close()


def test_dangling_tail_known_answers_in_one_batch(ctx):
    reached = [c for c in TAILS if c["reached"]]
    ro, refs, ao, alts = SC.pack([(c["ref"], c["alt"]) for c in reached])
    got = ctx.dangling_end_sw_batch(ro, refs, ao, alts, reached[0]["params"])
    assert got == [c["cigar"] for c in reached]


def test_dangling_end_batch_matches_oracle_on_random_pairs(ctx):
    # dangling paths are short (tens of bases) suffixes / prefixes that diverge from the reference path
    rng = np.random.default_rng(12)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    pairs = []
    for _ in range(400):
        n = int(rng.integers(5, 120))
        ref = rng.choice(acgt, n)
        alt = ref.copy()[:int(rng.integers(2, n + 1))]
        for _ in range(int(rng.integers(0, 4))):
            if len(alt) < 3:
                break
            p = int(rng.integers(0, len(alt)))
            u = rng.random()
            if u < 0.5:
                alt[p] = rng.choice(acgt)
            elif u < 0.75:
                alt = np.delete(alt, slice(p, p + int(rng.integers(1, 4))))
            else:
                alt = np.insert(alt, p, rng.choice(acgt, int(rng.integers(1, 4))))
        if len(alt) == 0:
            alt = ref[:1].copy()
        pairs.append((bytes(ref), bytes(alt)))
    ro, refs, ao, alts = SC.pack(pairs)
    for params in (SC.STANDARD_NGS, SC.NEW_SW_PARAMETERS):
        got = ctx.dangling_end_sw_batch(ro, refs, ao, alts, params, max_ops=128)
        for (r, a), g in zip(pairs, got):
            off, cigar = O.sw_align(r, a, params, "LeadingInDel")
            assert g == O_cigar_strip_trailing_deletion(cigar), (r, a, params)
    # capacity: a CIGAR that fits max_ops only once its trailing deletion is gone still comes back
    ro, refs, ao, alts = SC.pack([("TAAAAAAAAAA", "TCAAA")])
    assert O.sw_align("TAAAAAAAAAA", "TCAAA", SC.STANDARD_NGS, "LeadingInDel")[1] in ("5M6D", "5M")
    assert ctx.dangling_end_sw_batch(ro, refs, ao, alts, SC.STANDARD_NGS, max_ops=1) == ["5M"]


def _postops(ctx, args, v_ar, read_len, qual, ref_idx=0):
    """[A x R] matrix -> device epilogue; returns (values [A x R], keep)"""
    A, R = v_ar.shape
    b = lk.PairHmmBatch([0, R], [0, A], np.arange(R + 1) * read_len, np.full(R * read_len, ord("A"), np.uint8),
                        np.full(R * read_len, qual, np.uint8), np.full(R * read_len, 45, np.uint8), np.full(R * read_len, 45, np.uint8),
                        np.arange(A + 1) * 10, np.full(A * 10, ord("A"), np.uint8))
    lkh = np.ascontiguousarray(v_ar.T).ravel()  # [r*A + a]
    values = np.empty(A * R)
    keep = np.full(R, 9, np.uint8)
    s = b.as_struct()
    ref = np.array([ref_idx], np.int32)
    st = ctx._L.lk_engine_normalize_and_filter_device(ctx._h, C.byref(args), C.byref(s), ref.ctypes.data_as(_lib.i32p),
                                                      lkh.ctypes.data_as(_lib.f64p), values.ctypes.data_as(_lib.f64p),
                                                      keep.ctypes.data_as(_lib.u8p))
    ctx._check(st)
    return values.reshape(A, R), keep


@pytest.mark.parametrize("n_alleles,n_reads", [(1, 1), (2, 7), (3, 10), (5, 33), (2, 64), (4, 100)])
def test_device_filter_keeps_every_other_read(ctx, n_alleles, n_reads):
    # threshold exactly -100 for every read, as the reference test's closure |_read| -100.0: dynamic disqualification with
    # 100-base Q40 reads and an expected error rate of 0.25 gives min(dynamic ~ -0.6, ceil(100 * 0.25) * -4 = -100) = -100;
    # no normalisation (mismapping rate -inf), as the reference test calls the filter alone
    args = _lib.EngineArgs(10, float("-inf"), 0, 18, 1, 1.0, 0.25, 1, 0, 0)
    L = O.lib()
    L.lko_log10_min_true_likelihood.restype = C.c_double
    L.lko_log10_dynamic_read_qual_threshold.restype = C.c_double
    q = np.full(100, 40, np.uint8)
    assert L.lko_log10_min_true_likelihood(100, C.c_double(0.25), 0) == -100.0
    assert L.lko_log10_dynamic_read_qual_threshold(q.ctypes.data_as(O.u8p), 100, C.c_double(1.0)) > -100.0
    orig = _filter_case(n_alleles, n_reads)
    v, keep = _postops(ctx, args, orig, 100, 40)
    n_kept = (n_reads + 1) // 2
    assert list(keep) == [1 - (r & 1) for r in range(n_reads)]
    for r in range(n_kept):
        assert np.array_equal(v[:, r], orig[:, 2 * r])
    assert np.all(np.isnan(v[:, n_kept:]))


@pytest.mark.parametrize("seed", range(4))
def test_device_normalize_caps_the_worst_likelihood(ctx, seed):
    rng = np.random.default_rng(seed)
    A, R = int(rng.integers(2, 6)), int(rng.integers(1, 70))
    orig = -np.abs(rng.standard_normal((A, R))) * 3
    # normalize_likelihoods(-0.001, true); thresholds far below every value so that nothing is filtered
    args = _lib.EngineArgs(10, -0.001, 0, 18, 0, 1.0, 0.02, 1, 0, 0)
    v, keep = _postops(ctx, args, orig, 100, 40)
    assert np.all(keep == 1)
    assert np.array_equal(v, np.maximum(orig, orig.max(axis=0, keepdims=True) - 0.001))
