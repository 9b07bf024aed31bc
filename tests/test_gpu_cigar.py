"""GPU test (-m gpu): CigarUtils::calculate_cigar as ONE batch (host shortcuts + GPU Smith-Waterman + host CIGAR surgery)
against the reference's 114 known answers (tests/cigar_utils_unit_tests.rs:36-110) and the host-only path."""
import json
import os

import numpy as np
import pytest

import sw_cases as SC
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
CASES = json.load(open(os.path.join(GOLDEN, "calculate_cigar_cases.json")))


@pytest.fixture(scope="module")
def ctx():
    import lorikeet_b200 as lk
    c = lk.Context(0)
    yield c
    c.close()


def test_calculate_cigar_known_answers_one_batch(ctx):
    ro, refs, ao, alts = SC.pack([(c["ref"], c["alt"]) for c in CASES])
    got = ctx.calculate_cigar_batch(ro, refs, ao, alts, SC.NEW_SW_PARAMETERS, "InDel")
    assert got == [c["cigar"] for c in CASES]


def test_calculate_cigar_matches_host_path_on_haplotype_like_pairs(ctx):
    from test_cigar_utils import calculate_cigar_host
    from lorikeet_b200 import synth
    ro, refs, ao, alts = synth.sw_pairs(300, pad=0)   # un-padded pairs: calculate_cigar adds the 10 x N itself
    got = ctx.calculate_cigar_batch(ro, refs, ao, alts, SC.NEW_SW_PARAMETERS, "InDel")
    for i in range(0, 300, 7):
        exp = calculate_cigar_host(refs[ro[i]:ro[i + 1]], alts[ao[i]:ao[i + 1]])
        assert got[i] == exp
    # the three long pairs of test_avx_mode with the strategies the reference test uses
    pairs = [(p["ref"], p["alt"]) for p in SC.LONG]
    ro, refs, ao, alts = SC.pack(pairs)
    for strategy in ("InDel", "SoftClip", "LeadingInDel"):
        got = ctx.calculate_cigar_batch(ro, refs, ao, alts, SC.NEW_SW_PARAMETERS, strategy, max_ops=1024)
        exp = [calculate_cigar_host(r, a, strategy=strategy) for r, a in pairs]
        assert got == exp


def test_large_batch_multithreaded_host_path_matches_small_batches(ctx):
    # above 2048 pairs the per-pair host work of lk_calculate_cigar_batch (shortcuts, N padding, trim / left-align) is
    # spread over the host cores and the Smith-Waterman call runs as a multi-chunk pipeline: every CIGAR must equal what the
    # same pairs give in small single-threaded batches, including pairs the equal-length shortcut answers without SW
    from lorikeet_b200 import synth
    n = 9000
    ro, refs, ao, alts = synth.sw_pairs(n, first_pair=40000, pad=0)
    refs = refs.copy()
    # every 50th pair: alternate := reference with one mismatch, so the <= 2-mismatch shortcut path is in the mix
    alts = alts.copy()
    pieces_r, pieces_a = [], []
    for i in range(n):
        r = refs[ro[i]:ro[i + 1]]
        a = alts[ao[i]:ao[i + 1]]
        if i % 50 == 0:
            a = r.copy()
            a[len(a) // 2] = ord("A") if a[len(a) // 2] != ord("A") else ord("C")
        pieces_r.append(r); pieces_a.append(a)
    ro2, refs2, ao2, alts2 = SC.pack(list(zip(pieces_r, pieces_a)))
    big = ctx.calculate_cigar_batch(ro2, refs2, ao2, alts2, SC.NEW_SW_PARAMETERS, "InDel")
    assert len(big) == n and sum(c is None for c in big) < n // 10
    for lo in range(0, n, 1500):
        hi = min(lo + 1500, n)
        r_off = ro2[lo:hi + 1] - ro2[lo]
        a_off = ao2[lo:hi + 1] - ao2[lo]
        small = ctx.calculate_cigar_batch(r_off, refs2[ro2[lo]:ro2[hi]], a_off, alts2[ao2[lo]:ao2[hi]], SC.NEW_SW_PARAMETERS, "InDel")
        assert small == big[lo:hi]
