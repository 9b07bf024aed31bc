"""CPU tests of the packed two-pairs-per-warp Smith-Waterman routine (lorikeet_b200/csrc/sw16_core.cuh) through the
lock-step emulator, bit-exact against the oracle (= the reference's scalar path,
src/smith_waterman/smith_waterman_aligner.rs:124-442).  The authoritative parity tests are the -m gpu ones."""
import numpy as np
import pytest

import emul_runner as E
import sw_cases as SC
from oracle import lk_oracle as O

ACGT = np.frombuffer(b"ACGT", np.uint8)


def _mutate(rng, ref, n_edits, max_indel):
    alt = ref.copy()
    for _ in range(n_edits):
        if len(alt) < 4:
            break
        p = int(rng.integers(1, len(alt) - 1))
        t = rng.integers(0, 3)
        if t == 0:
            alt[p] = rng.choice(ACGT)
        elif t == 1:
            alt = np.delete(alt, slice(p, p + int(rng.integers(1, max_indel + 1))))
        else:
            alt = np.insert(alt, p, rng.choice(ACGT, int(rng.integers(1, max_indel + 1))))
    return alt


def equal_n_pairs(n_duos, seed, min_len, max_len, max_indel=6):
    """Duos of pairs whose references have the same length (what the packed path needs); alternates of different lengths:
    mutated copies, shifted windows (overhangs either side), unrelated sequences."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n_duos):
        n = int(rng.integers(min_len, max_len + 1))
        for _ in range(2):
            ref = rng.choice(ACGT, n)
            mode = int(rng.integers(0, 5))
            if mode == 0:
                alt = rng.choice(ACGT, int(rng.integers(max(1, n // 2), n + 20)))
            else:
                alt = _mutate(rng, ref, int(rng.integers(0, 5)), max_indel)
                if mode == 2 and len(alt) > 8:
                    alt = alt[int(rng.integers(1, len(alt) // 3)):]
                if mode == 3:
                    alt = np.concatenate([alt, rng.choice(ACGT, int(rng.integers(1, 9)))])
                if mode == 4 and len(alt) > 8:
                    alt = np.concatenate([rng.choice(ACGT, int(rng.integers(1, 9))), alt[:-int(rng.integers(1, len(alt) // 3))]])
            out.append((ref, alt))
    return out


def _compare(pairs, params, strategy, max_ops=256, force_k=0, expect_duo=None):
    ro, refs, ao, alts = SC.pack(pairs)
    ops, n_ops, offs, took = E.sw_duo_batch(ro, refs, ao, alts, params, strategy, max_ops=max_ops, force_k=force_k)
    eops, en, eoffs = O.sw_batch(ro, refs, ao, alts, params, strategy, max_ops=max_ops)
    for i in range(len(pairs)):
        ctx = (i, strategy, params, bool(took[i]), O.cigar_to_string(ops[i, :max(n_ops[i], 0)]), O.cigar_to_string(eops[i, :en[i]]))
        assert n_ops[i] == en[i], ctx
        assert offs[i] == eoffs[i], ctx
        assert np.array_equal(ops[i, :n_ops[i]], eops[i, :en[i]]), ctx
    if expect_duo is not None:
        assert bool(took.any()) == expect_duo
    return took


@pytest.mark.parametrize("strategy", SC.STRATEGIES)
@pytest.mark.parametrize("force_k", [8, 16])
def test_small_equal_length_references_all_params(strategy, force_k):
    # short references under K = 8 / 16: up to 31 of the 32 lanes hold virtual rows (SoftClip / Ignore) or stay idle (InDel)
    pairs = equal_n_pairs(24, seed=len(strategy), min_len=3, max_len=90)
    taken = 0
    for params in SC.ALL_PARAMS:
        taken += int(_compare(pairs, params, strategy, force_k=force_k).sum())
    assert taken > 0


def test_k8_class_sizes():
    # the class the host gives K = 8: references of 129..256 rows
    pairs = equal_n_pairs(6, seed=88, min_len=129, max_len=256, max_indel=15)
    for strategy in SC.STRATEGIES:
        _compare(pairs, SC.NEW_SW_PARAMETERS, strategy, max_ops=512, force_k=8)


def test_k12_class_sizes():
    # K = 12 (references of 257..384 rows): the second quad of 16-bit table rows is half used, six row pairs per plane byte
    pairs = equal_n_pairs(4, seed=12, min_len=257, max_len=384, max_indel=15)
    pairs += [(pairs[0][0][:300], pairs[1][1]), (pairs[2][0], pairs[3][1][:200])]  # and a duo of different lengths
    for strategy in SC.STRATEGIES:
        took = _compare(pairs, SC.NEW_SW_PARAMETERS, strategy, max_ops=512, force_k=12)
        assert took.any()
    small = equal_n_pairs(12, seed=13, min_len=3, max_len=120)
    for params in SC.ALL_PARAMS:
        _compare(small, params, "SoftClip", force_k=12)
        _compare(small, params, "LeadingInDel", force_k=12)


@pytest.mark.parametrize("force_k", [16, 17])
def test_haplotype_sized_duos(force_k):
    # references of 300..544 rows, alternates of different lengths within a duo: lean middle steps, the tail that runs to the
    # longer alternate, K = 17's row-16 byte, bottom alignment with few and many virtual rows
    rng = np.random.default_rng(force_k)
    pairs = []
    for n in ((300, 512) if force_k == 16 else (513, 544, 520)):
        for _ in range(2):
            ref = rng.choice(ACGT, n)
            pairs.append((ref, _mutate(rng, ref, 4, 12)))
    ref = rng.choice(ACGT, 520 if force_k == 17 else 480)
    pairs += [(ref, rng.choice(ACGT, 430)), (ref.copy(), _mutate(rng, ref, 2, 30))]  # unrelated next to related, same reference
    for strategy in SC.STRATEGIES:
        took = _compare(pairs, SC.NEW_SW_PARAMETERS, strategy, max_ops=512, force_k=force_k)
        assert took.all() or strategy in ("SoftClip", "Ignore")  # an exact-substring hit bypasses the matrix
    _compare(pairs, SC.STANDARD_NGS, "LeadingInDel", max_ops=512, force_k=force_k)
    _compare(pairs, SC.ALIGNMENT_TO_BEST_HAPLOTYPE, "SoftClip", max_ops=512, force_k=force_k)


@pytest.mark.parametrize("strategy", SC.STRATEGIES)
def test_references_of_different_lengths_share_a_warp(strategy):
    # each pair gets its own number of virtual rows (SoftClip / Ignore) or is followed by unread rows (InDel / LeadingInDel)
    rng = np.random.default_rng(len(strategy) + 40)
    pairs = []
    for _ in range(12):
        for n in (int(rng.integers(5, 200)), int(rng.integers(5, 200))):
            ref = rng.choice(ACGT, n)
            mode = int(rng.integers(0, 3))
            alt = _mutate(rng, ref, int(rng.integers(0, 4)), 6) if mode else rng.choice(ACGT, int(rng.integers(3, 150)))
            if mode == 2 and len(alt) > 10:
                alt = alt[int(rng.integers(1, 5)):-int(rng.integers(1, 5))]
            pairs.append((ref, alt))
    pairs += [(rng.choice(ACGT, 256), rng.choice(ACGT, 40)), (rng.choice(ACGT, 1), rng.choice(ACGT, 90))]  # the extremes of the class
    taken = 0
    for params in SC.ALL_PARAMS:
        taken += int(_compare(pairs, params, strategy, force_k=8).sum())
    assert taken > 0


def test_dispatch_falls_back_pair_by_pair():
    rng = np.random.default_rng(5)
    ref = rng.choice(ACGT, 60)
    a1, a2 = _mutate(rng, ref, 3, 5), _mutate(rng, ref, 3, 5)
    other = np.frombuffer(b"ACGTNRYacgt", np.uint8)
    pairs = [(ref, a1), (rng.choice(ACGT, 700), a2),                    # a reference beyond the class's one band
             (ref, a1), (ref, np.concatenate([a2[:20], other, a2[20:]])),  # bytes outside A C G T N in one alternate
             (ref, ref[5:40]), (ref, a2),                                # exact substring: finished by the prologue
             (ref, a1)]                                                  # odd one out
    for strategy in ("SoftClip", "InDel"):
        took = _compare(pairs, SC.NEW_SW_PARAMETERS, strategy, force_k=16)
        assert not took[:4].any() and not took[6]
    took = _compare(pairs, SC.NEW_SW_PARAMETERS, "InDel", force_k=16)
    assert took[4] and took[5]  # InDel has no substring shortcut


def test_reference_known_answers_in_duos():
    # the reference's own known answers (tests/smith_waterman_aligner_unit_tests.rs), each paired with itself
    for c in SC.KAT:
        if any(ch not in "ACGTN" for ch in c["alt"]) or not E_weights_ok(c["params"]):
            continue
        ro, refs, ao, alts = SC.pack([(c["ref"], c["alt"])] * 2)
        ops, n_ops, offs, took = E.sw_duo_batch(ro, refs, ao, alts, c["params"], c["strategy"], force_k=16)
        for i in range(2):
            assert (int(offs[i]), O.cigar_to_string(ops[i, :n_ops[i]])) == (c["offset"], c["cigar"]), (c, bool(took[i]))


def E_weights_ok(p):
    return all(abs(x) <= 1000 for x in p) and p[2] <= p[3] < 0


def test_capacity_overflow_in_a_duo():
    rng = np.random.default_rng(9)
    ref = rng.choice(ACGT, 80)
    pairs = [(ref, _mutate(rng, ref, 6, 4)), (ref, ref.copy())]
    ro, refs, ao, alts = SC.pack(pairs)
    ops, n_ops, offs, took = E.sw_duo_batch(ro, refs, ao, alts, SC.NEW_SW_PARAMETERS, "InDel", max_ops=1, force_k=16)
    eops, en, eoffs = O.sw_batch(ro, refs, ao, alts, SC.NEW_SW_PARAMETERS, "InDel", max_ops=1)
    assert took.all() and list(n_ops) == list(en)


def test_n_padded_haplotypes_as_calculate_cigar_makes_them():
    # CigarUtils::calculate_cigar wraps both sequences in ten N (reads/cigar_utils.rs:10,387-398): N columns match N rows only
    rng = np.random.default_rng(21)
    pad = np.frombuffer(b"N" * 10, np.uint8)
    pairs = []
    for _ in range(3):
        n = int(rng.integers(40, 120))
        for _ in range(2):
            ref = rng.choice(ACGT, n)
            alt = _mutate(rng, ref, 3, 8)
            alt[int(rng.integers(0, len(alt)))] = ord("N")  # and a stray N inside
            pairs.append((np.concatenate([pad, ref, pad]), np.concatenate([pad, alt, pad])))
    for strategy in SC.STRATEGIES:
        took = _compare(pairs, SC.NEW_SW_PARAMETERS, strategy, force_k=16)
        assert took.all()


def test_random_weights_up_to_the_eligibility_limits():
    # the packed arithmetic keeps every 16-bit half inside [0, 65535] only for bounded weights (sw16_weights_ok: |w| <= 1000,
    # open <= extend < 0): random parameter sets up to those limits, including the corners, against the oracle
    rng = np.random.default_rng(1000)
    sets = [[1000, -1000, -1000, -1], [1000, -1000, -1000, -1000], [1, -1, -1, -1], [1000, 1000, -1000, -500],
            [-1000, -1000, -2, -1], [0, 0, -1000, -999], [7, -1000, -13, -13], [1000, -1, -1000, -1]]
    for _ in range(10):
        wo = -int(rng.integers(1, 1001))
        sets.append([int(rng.integers(-50, 1001)), int(rng.integers(-1000, 51)), wo, -int(rng.integers(1, -wo + 1))])
    pairs = equal_n_pairs(10, seed=31, min_len=20, max_len=120, max_indel=10)
    for i, params in enumerate(sets):
        assert E_weights_ok(params), params
        strategy = SC.STRATEGIES[i % 4]
        took = _compare(pairs, params, strategy, force_k=8)
        assert took.any() or strategy in ("SoftClip", "Ignore")


def test_long_gaps_and_repeats():
    # long insertions / deletions (gap length recovered by walking the early-stored open flags) and low-complexity sequence
    # (many ties between opening and extending, and between the three moves)
    rng = np.random.default_rng(77)
    pairs = []
    for n in (150, 200):
        ref = rng.choice(ACGT, n)
        pairs.append((ref, np.concatenate([ref[:40], ref[100:]])))                      # 60-base deletion
        pairs.append((ref.copy(), np.concatenate([ref[:70], rng.choice(ACGT, 45), ref[70:]])))  # 45-base insertion
    rep = np.frombuffer(b"ACAC" * 40, np.uint8)
    pairs += [(rep, np.frombuffer(b"AC" * 71, np.uint8)), (rep.copy(), np.frombuffer(b"ACAC" * 20 + b"AAAA" + b"ACAC" * 18, np.uint8))]
    homo = np.frombuffer(b"A" * 140, np.uint8)
    pairs += [(homo, np.frombuffer(b"A" * 100, np.uint8)), (homo.copy(), np.frombuffer(b"A" * 60 + b"C" + b"A" * 90, np.uint8))]
    for strategy in SC.STRATEGIES:
        for params in (SC.NEW_SW_PARAMETERS, SC.ORIGINAL_DEFAULT, [5, -4, -6, -6]):
            _compare(pairs, params, strategy, max_ops=256, force_k=8)
