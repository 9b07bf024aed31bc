"""CPU tests: the oracle's Smith-Waterman restatement against the reference's own known answers
(tests/smith_waterman_aligner_unit_tests.rs:229-400) and structural invariants."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import lk_oracle as O

KAT = json.load(open(os.path.join(GOLDEN, "sw_known_answers.json")))
LONG = json.load(open(os.path.join(GOLDEN, "sw_long_pairs.json")))

NEW_SW_PARAMETERS = [200, -150, -260, -11]
STANDARD_NGS = [25, -50, -110, -6]
ALIGNMENT_TO_BEST_HAPLOTYPE = [10, -15, -30, -5]
ORIGINAL_DEFAULT = [3, -1, -4, -3]


@pytest.mark.parametrize("case", KAT, ids=[c["source"].split(":")[-1] + c["strategy"] for c in KAT])
def test_known_answers(case):
    off, cig = O.sw_align(case["ref"], case["alt"], case["params"], case["strategy"])
    assert (off, cig) == (case["offset"], case["cigar"])


@pytest.mark.parametrize("case", KAT, ids=[c["source"].split(":")[-1] + c["strategy"] for c in KAT])
def test_known_answers_without_substring_shortcut(case):
    # the exact-substring shortcut (smith_waterman_aligner.rs:70-80) must agree with the matrix path
    off, cig = O.sw_align(case["ref"], case["alt"], case["params"], case["strategy"], use_shortcut=False)
    assert (off, cig) == (case["offset"], case["cigar"])


def cigar_lengths(cig):
    import re
    ref = alt = 0
    for n, op in re.findall(r"(\d+)([MIDS])", cig):
        n = int(n)
        if op in "M":
            ref += n
            alt += n
        elif op == "I" or op == "S":
            alt += n
        elif op == "D":
            ref += n
    return ref, alt


@pytest.mark.parametrize("idx", [0, 1, 2])
@pytest.mark.parametrize("params", [NEW_SW_PARAMETERS, STANDARD_NGS, ALIGNMENT_TO_BEST_HAPLOTYPE])
@pytest.mark.parametrize("strategy", ["InDel", "SoftClip", "LeadingInDel", "Ignore"])
def test_long_pairs_consistent(idx, params, strategy):
    # test_avx_mode (:403-1101) runs these pairs with 10xN padding; here: structural checks
    pad = "N" * 10
    ref = pad + LONG[idx]["ref"] + pad
    alt = pad + LONG[idx]["alt"] + pad
    off, cig = O.sw_align(ref, alt, params, strategy)
    r, a = cigar_lengths(cig)
    if strategy != "Ignore":  # Ignore folds the leading overhang p2 into the last element, whatever its type (:420-423)
        assert a == len(alt)
    if strategy == "InDel":
        assert r == len(ref) and off == 0


def test_empty_sequences_rejected():
    with pytest.raises(ValueError):
        O.sw_align("", "ACGT", ORIGINAL_DEFAULT, "SoftClip")
    with pytest.raises(ValueError):
        O.sw_align("ACGT", "", ORIGINAL_DEFAULT, "InDel")


def test_last_index_of():
    L = O.lib()
    f = lambda r, q: L.lko_last_index_of(O._p(O._u8(r), O.u8p), len(r), O._p(O._u8(q), O.u8p), len(q))
    assert f(b"ACGTACGT", b"ACG") == 4
    assert f(b"ACGT", b"ACGTA") == -1
    assert f(b"AAAA", b"AA") == 2
    assert f(b"ACGT", b"G") == 2


def test_batch_driver_matches_single():
    rng = np.random.default_rng(3)
    refs, alts = [], []
    for _ in range(40):
        r = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), int(rng.integers(20, 80)))
        a = r.copy()
        k = int(rng.integers(0, len(a) - 5))
        a = np.concatenate([a[:k], a[k + int(rng.integers(1, 5)):]])
        refs.append(r)
        alts.append(a)
    ro = np.zeros(41, np.int64); ro[1:] = np.cumsum([len(x) for x in refs])
    ao = np.zeros(41, np.int64); ao[1:] = np.cumsum([len(x) for x in alts])
    ops, n_ops, offs = O.sw_batch(ro, np.concatenate(refs), ao, np.concatenate(alts), NEW_SW_PARAMETERS, "SoftClip",
                                  max_ops=32, n_threads=3)
    for i in range(40):
        off, cig = O.sw_align(refs[i], alts[i], NEW_SW_PARAMETERS, "SoftClip")
        assert off == offs[i] and cig == O.cigar_to_string(ops[i, :n_ops[i]])
