"""CPU tests of the CUDA Smith-Waterman warp routine through the lock-step emulator (tests/emul), bit-exact
against the oracle (= the reference's scalar path).  The authoritative parity tests are the -m gpu ones."""
import numpy as np
import pytest

import emul_runner as E
import sw_cases as SC
from oracle import lk_oracle as O


def _compare(pairs, params, strategy, max_ops=256, force_k=0):
    ro, refs, ao, alts = SC.pack(pairs)
    ops, n_ops, offs = E.sw_batch(ro, refs, ao, alts, params, strategy, max_ops=max_ops, force_k=force_k)
    eops, en, eoffs = O.sw_batch(ro, refs, ao, alts, params, strategy, max_ops=max_ops)
    for i in range(len(pairs)):
        assert n_ops[i] == en[i], (i, strategy, params, O.cigar_to_string(ops[i, :max(n_ops[i], 0)]), O.cigar_to_string(eops[i, :en[i]]))
        assert offs[i] == eoffs[i], (i, strategy, params)
        assert np.array_equal(ops[i, :n_ops[i]], eops[i, :en[i]]), (i, strategy, params, O.cigar_to_string(ops[i, :n_ops[i]]), O.cigar_to_string(eops[i, :en[i]]))


def test_known_answers():
    for c in SC.KAT:
        ro, refs, ao, alts = SC.pack([(c["ref"], c["alt"])])
        ops, n_ops, offs = E.sw_batch(ro, refs, ao, alts, c["params"], c["strategy"])
        assert (int(offs[0]), O.cigar_to_string(ops[0, :n_ops[0]])) == (c["offset"], c["cigar"]), c


@pytest.mark.parametrize("strategy", SC.STRATEGIES)
def test_random_pairs_all_params(strategy):
    pairs = SC.random_pairs(60, seed=hash(strategy) % 1000)
    for params in SC.ALL_PARAMS:
        _compare(pairs, params, strategy)


@pytest.mark.parametrize("force_k", [2, 4, 12])
def test_multi_band_small_k(force_k):
    # forcing a small K makes references of 70..200 rows span several row bands
    pairs = SC.random_pairs(16, seed=3, min_len=70, max_len=200)
    for strategy in SC.STRATEGIES:
        _compare(pairs, SC.NEW_SW_PARAMETERS, strategy, force_k=force_k)


def test_long_pair_reference_fixture():
    # one of the test_avx_mode pairs (ref 275+20 x alt 1154+20), 3 strategies as in the reference test
    pair = SC.long_pairs_padded()[:1]
    for strategy in ("InDel", "SoftClip", "LeadingInDel"):
        _compare(pair, SC.NEW_SW_PARAMETERS, strategy, max_ops=512)


def test_capacity_and_empty():
    ro, refs, ao, alts = SC.pack([("ACGTACGTAAAA", "ACGAACGTTTTAAAA"), ("ACGT", "ACGT")])
    ops, n_ops, offs = E.sw_batch(ro, refs, ao, alts, SC.ORIGINAL_DEFAULT, "InDel", max_ops=1)
    assert n_ops[0] == -2 and n_ops[1] == 1
    ro = np.array([0, 0, 4]); ao = np.array([0, 3, 3])
    ops, n_ops, offs = E.sw_batch(ro, np.frombuffer(b"ACGT", np.uint8), ao, np.frombuffer(b"ACG", np.uint8), SC.ORIGINAL_DEFAULT, "SoftClip")
    assert list(n_ops) == [-1, -1]


def test_bytes_outside_acgtn_take_the_compare_path():
    # the substitution-weight table covers the exact bytes A C G T N; any other byte in an alternate column (lower case,
    # IUPAC codes, '-') must fall back to the reference's raw byte comparison for that column
    pairs = SC.random_pairs(40, seed=17, alphabet=b"ACGTNacgtRY-")
    for strategy in ("SoftClip", "InDel"):
        _compare(pairs, SC.NEW_SW_PARAMETERS, strategy)


@pytest.mark.parametrize("force_k", [16, 17])
def test_lean_steps_and_packed_planes(force_k):
    # references of 500..544 rows against alternates of 450..560 columns: the ungated middle steps (32 .. m-1) carry almost
    # all cells, K = 16 uses the two plane words only, K = 17 also the row-16 byte; every strategy (bottom-row summary from
    # scratch for SoftClip / Ignore, band-boundary feed when K = 16 needs a second band for > 512 rows)
    rng = np.random.default_rng(force_k)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    pairs = []
    for n in (500, 512, 513, 544):
        ref = rng.choice(acgt, n)
        alt = ref.copy()
        for _ in range(3):  # a few indels and SNPs
            p = int(rng.integers(20, len(alt) - 20))
            if rng.random() < 0.5:
                alt = np.delete(alt, slice(p, p + int(rng.integers(1, 12))))
            else:
                alt = np.insert(alt, p, rng.choice(acgt, int(rng.integers(1, 12))))
        alt[rng.integers(0, len(alt), 5)] = rng.choice(acgt, 5)
        pairs.append((bytes(ref), bytes(alt)))
    pairs.append((bytes(rng.choice(acgt, 520)), bytes(rng.choice(acgt, 470))))  # unrelated
    for strategy in SC.STRATEGIES:
        _compare(pairs, SC.NEW_SW_PARAMETERS, strategy, max_ops=512, force_k=force_k)
    _compare(pairs, SC.STANDARD_NGS, "LeadingInDel", max_ops=512, force_k=force_k)
