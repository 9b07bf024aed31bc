"""GPU parity tests (-m gpu): the CUDA Smith-Waterman path through the C ABI must reproduce the reference's
(offset, CIGAR) bit for bit (BASELINE.json north_star: "SW CIGARs bit-exact").  Mirrors
tests/smith_waterman_aligner_unit_tests.rs of the reference."""
import numpy as np
import pytest

import sw_cases as SC
from oracle import lk_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import lorikeet_b200 as lk
    c = lk.Context(0)
    yield c
    c.close()


def _compare(ctx, pairs, params, strategy, max_ops=256):
    ro, refs, ao, alts = SC.pack(pairs)
    ops, n_ops, offs = ctx.sw_batch(ro, refs, ao, alts, params, strategy, max_ops=max_ops)
    eops, en, eoffs = O.sw_batch(ro, refs, ao, alts, params, strategy, max_ops=max_ops)
    assert np.array_equal(n_ops, en)
    assert np.array_equal(offs, eoffs)
    for i in range(len(pairs)):
        assert np.array_equal(ops[i, :max(n_ops[i], 0)], eops[i, :max(en[i], 0)]), (i, strategy, params)


@pytest.mark.parametrize("case", SC.KAT, ids=[c["source"].split(":")[-1] + c["strategy"] for c in SC.KAT])
def test_known_answers_single_pair_entry_point(ctx, case):
    # the gkl::smithwaterman::align()-shaped wrapper: ASCII CIGAR + offset (smith_waterman_aligner.rs:56-61)
    off, cigar = ctx.sw_align(case["ref"], case["alt"], case["params"], case["strategy"])
    assert (off, cigar) == (case["offset"], case["cigar"])


@pytest.mark.parametrize("strategy", SC.STRATEGIES)
@pytest.mark.parametrize("params", SC.ALL_PARAMS, ids=["new", "ngs", "besthap", "orig"])
def test_random_pairs(ctx, strategy, params):
    _compare(ctx, SC.random_pairs(400, seed=len(strategy) * 7 + params[0], max_len=120), params, strategy)


@pytest.mark.parametrize("strategy", ["InDel", "SoftClip", "LeadingInDel", "Ignore"])
def test_avx_mode_long_pairs(ctx, strategy):
    # test_avx_mode (:403-1101): 3 long pairs x {NEW_SW, STANDARD_NGS, ALIGNMENT_TO_BEST_HAPLOTYPE}, 10xN padded
    for params in (SC.NEW_SW_PARAMETERS, SC.STANDARD_NGS, SC.ALIGNMENT_TO_BEST_HAPLOTYPE):
        _compare(ctx, SC.long_pairs_padded(), params, strategy, max_ops=1024)


def test_thread_safety_hammer(ctx):
    # the reference hammers align() from 1000 parallel tasks (:1001); here 8 host threads share one context
    import threading
    pairs = SC.long_pairs_padded()
    ro, refs, ao, alts = SC.pack(pairs)
    exp = O.sw_batch(ro, refs, ao, alts, SC.NEW_SW_PARAMETERS, "SoftClip", max_ops=1024)
    errs = []

    def work():
        try:
            for _ in range(5):
                got = ctx.sw_batch(ro, refs, ao, alts, SC.NEW_SW_PARAMETERS, "SoftClip", max_ops=1024)
                assert np.array_equal(got[1], exp[1]) and np.array_equal(got[2], exp[2])
                for i in range(len(pairs)):
                    assert np.array_equal(got[0][i, :got[1][i]], exp[0][i, :exp[1][i]])
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work) for _ in range(8)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs[:1]


def test_cfg4_slice_all_strategies(ctx):
    # cfg#4 generator (500 bp + 10xN pads, NEW_SW_PARAMETERS): SoftClip on 1500 pairs, the others on a 10 % slice
    from lorikeet_b200 import synth
    ro, refs, ao, alts = synth.sw_pairs(1500)
    for strategy, n in (("SoftClip", 1500), ("InDel", 150), ("LeadingInDel", 150), ("Ignore", 150)):
        r, a = ro[:n + 1], ao[:n + 1]
        ops, n_ops, offs = ctx.sw_batch(r, refs[:r[-1]], a, alts[:a[-1]], SC.NEW_SW_PARAMETERS, strategy, max_ops=64)
        eops, en, eoffs = O.sw_batch(r, refs[:r[-1]], a, alts[:a[-1]], SC.NEW_SW_PARAMETERS, strategy, max_ops=64)
        assert np.array_equal(n_ops, en) and np.array_equal(offs, eoffs)
        valid = np.arange(64)[None, :] < n_ops[:, None]  # entries past n_ops[p] are unspecified
        assert np.array_equal(ops[valid], eops[valid])
    ops, n_ops, offs = ctx.sw_batch(ro[:151], refs[:ro[150]], ao[:151], alts[:ao[150]], SC.STANDARD_NGS, "LeadingInDel")
    eops, en, eoffs = O.sw_batch(ro[:151], refs[:ro[150]], ao[:151], alts[:ao[150]], SC.STANDARD_NGS, "LeadingInDel")
    valid = np.arange(64)[None, :] < n_ops[:, None]
    assert np.array_equal(n_ops, en) and np.array_equal(ops[valid], eops[valid]) and np.array_equal(offs, eoffs)


def test_multi_band_long_reference(ctx):
    # references longer than 32*17 rows exercise the band boundary (H, F) hand-off
    rng = np.random.default_rng(4)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    pairs = []
    for n in (545, 700, 1300, 2000):
        ref = rng.choice(acgt, n)
        alt = np.concatenate([ref[30:n // 2], rng.choice(acgt, 17), ref[n // 2 + 40:n - 11]])
        alt[::97] = rng.choice(acgt, len(alt[::97]))
        pairs.append((ref, alt))
    for strategy in SC.STRATEGIES:
        _compare(ctx, pairs, SC.NEW_SW_PARAMETERS, strategy, max_ops=512)


def test_resident_repeatable_and_edge_cases(ctx):
    from lorikeet_b200 import synth
    ro, refs, ao, alts = synth.sw_pairs(300, first_pair=5000)
    res = ctx.sw_upload(ro, refs, ao, alts, SC.NEW_SW_PARAMETERS, "SoftClip", max_ops=64)
    ms = res.run()
    a = res.download()
    res.run()
    b = res.download()
    res.free()
    assert ms > 0 and all(np.array_equal(x, y) for x, y in zip(a, b))
    # empty sequences are per-pair errors (the reference asserts), a too-small max_ops is -2
    ro = np.array([0, 0, 4, 16]); ao = np.array([0, 3, 3, 18])
    ops, n_ops, offs = ctx.sw_batch(ro, np.frombuffer(b"ACGTACGTACGTAAAA", np.uint8), ao,
                                    np.frombuffer(b"ACGACGAACGTTTTAAAA", np.uint8), SC.ORIGINAL_DEFAULT, "InDel", max_ops=1)
    assert list(n_ops) == [-1, -1, -2]
    import lorikeet_b200 as lk
    with pytest.raises(lk.LorikeetError):
        ctx.sw_align(b"", b"ACGT", SC.ORIGINAL_DEFAULT, "SoftClip")


def test_host_pipeline_matches_resident(ctx):
    # lk_sw_batch_host cuts a large batch into contiguous pair ranges that flow through three workspaces / streams
    # (8.2e9 cells -> 3 chunks here); every (n_ops, offset, ops[:n_ops]) must equal the single resident launch set, with
    # short pairs of the other K classes mixed in so each chunk launches several kernels
    from lorikeet_b200 import synth
    rng = np.random.default_rng(77)
    ro, refs, ao, alts = synth.sw_pairs(30000, first_pair=20000)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    parts_r, parts_a = [], []
    for _ in range(600):
        n = int(rng.integers(1, 260))
        r = rng.choice(acgt, n)
        a = r.copy()
        if n > 8:
            a = np.delete(a, slice(3, 3 + int(rng.integers(0, 4))))
            a[int(rng.integers(0, a.size))] = ord("A")
        parts_r.append(r); parts_a.append(a)
    sro = np.concatenate([[0], np.cumsum([len(x) for x in parts_r])]).astype(np.int64)
    sao = np.concatenate([[0], np.cumsum([len(x) for x in parts_a])]).astype(np.int64)
    # short pairs first, then the long ones, then the short ones again
    all_ro = np.concatenate([sro, sro[-1] + ro[1:], sro[-1] + ro[-1] + sro[1:]])
    all_ao = np.concatenate([sao, sao[-1] + ao[1:], sao[-1] + ao[-1] + sao[1:]])
    all_refs = np.concatenate(parts_r + [refs] + parts_r)
    all_alts = np.concatenate(parts_a + [alts] + parts_a)
    ops, n_ops, offs = ctx.sw_batch(all_ro, all_refs, all_ao, all_alts, SC.NEW_SW_PARAMETERS, "SoftClip", max_ops=64)
    res = ctx.sw_upload(all_ro, all_refs, all_ao, all_alts, SC.NEW_SW_PARAMETERS, "SoftClip", max_ops=64)
    res.run()
    eops, en, eoffs = res.download()
    res.free()
    assert np.array_equal(n_ops, en) and np.array_equal(offs, eoffs) and (n_ops > 0).all()
    valid = np.arange(64)[None, :] < n_ops[:, None]
    assert np.array_equal(ops[valid], eops[valid])
    # the two copies of the short pairs sit in different chunks and must agree with each other
    assert np.array_equal(n_ops[:600], n_ops[-600:]) and np.array_equal(offs[:600], offs[-600:])


def test_bytes_outside_acgtn_and_clamped_form(ctx):
    # alternate columns holding bytes outside A C G T N use the raw byte compare instead of the weight table; weights large
    # enough that (n + m) * max|w| >= 1e8 select the kernel that keeps the reference's -1e8 clamp
    pairs = SC.random_pairs(400, seed=23, alphabet=b"ACGTNacgtRY-")
    for params, strategy in ((SC.NEW_SW_PARAMETERS, "SoftClip"), (SC.STANDARD_NGS, "InDel"), ([900000, -700000, -800000, -100000], "SoftClip"),
                             ([900000, -700000, -800000, -100000], "LeadingInDel")):
        _compare(ctx, pairs, params, strategy)


def test_packed_two_pairs_per_warp_kernel(ctx):
    """References of 129..544 rows go two to a warp (sw16_core.cuh, 16-bit difference arithmetic): same (offset, CIGAR) as the
    oracle for every strategy and parameter set, references and alternates of different lengths inside a duo, N pads as
    calculate_cigar adds them, and pairs that do not qualify (banded references, bytes outside ACGTN, exact-substring hits)
    mixed in."""
    rng = np.random.default_rng(77)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    pad = np.frombuffer(b"N" * 10, np.uint8)

    def mutate(ref, n_edits, max_indel):
        alt = ref.copy()
        for _ in range(n_edits):
            p = int(rng.integers(5, len(alt) - 5))
            t = rng.integers(0, 3)
            if t == 0:
                alt[p] = rng.choice(acgt)
            elif t == 1:
                alt = np.delete(alt, slice(p, p + int(rng.integers(1, max_indel + 1))))
            else:
                alt = np.insert(alt, p, rng.choice(acgt, int(rng.integers(1, max_indel + 1))))
        return alt

    pairs = []
    for n in (140, 200, 256, 260, 300, 411, 500, 512, 513, 520, 524, 544):
        for _ in range(5):
            ref = rng.choice(acgt, n - 20)
            mode = int(rng.integers(0, 4))
            alt = mutate(ref, int(rng.integers(1, 6)), 25)
            if mode == 1:
                alt = alt[int(rng.integers(1, 60)):]
            if mode == 2:
                alt = rng.choice(acgt, int(rng.integers(n // 2, n)))
            pairs.append((np.concatenate([pad, ref, pad]), np.concatenate([pad, alt, pad])))
    ref = rng.choice(acgt, 400)
    pairs += [(ref, ref[20:380].copy()), (ref, mutate(ref, 3, 9)),                       # substring hit next to a real one
              (ref, np.concatenate([mutate(ref, 2, 5), np.frombuffer(b"RY", np.uint8)])),  # bytes outside ACGTN
              (ref, mutate(ref, 2, 5)), (rng.choice(acgt, 700), rng.choice(acgt, 650))]    # and a banded reference
    for strategy in SC.STRATEGIES:
        _compare(ctx, pairs, SC.NEW_SW_PARAMETERS, strategy, max_ops=512)
    _compare(ctx, pairs, SC.STANDARD_NGS, "LeadingInDel", max_ops=512)
    _compare(ctx, pairs, SC.ALIGNMENT_TO_BEST_HAPLOTYPE, "SoftClip", max_ops=512)
    _compare(ctx, pairs, SC.ORIGINAL_DEFAULT, "InDel", max_ops=512)
    # the packed kernel is what ran: nearly all 60 pairs of the three classes under InDel (no substring shortcut)
    ro, refs, ao, alts = SC.pack(pairs)
    r = ctx.sw_upload(ro, refs, ao, alts, SC.NEW_SW_PARAMETERS, "InDel", max_ops=512)
    r.run()
    assert r.packed_pairs >= 40, r.packed_pairs
    # weights outside the packed kernel's range keep the one-pair kernel, with the same answers
    big = [2000, -1500, -2600, -110]
    _compare(ctx, pairs[:12], big, "SoftClip", max_ops=512)
    r2 = ctx.sw_upload(ro, refs, ao, alts, big, "InDel", max_ops=512)
    r2.run()
    assert r2.packed_pairs == 0
    r.free(); r2.free()


def test_mixed_lengths_at_scale(ctx):
    """2 400 pairs with reference lengths drawn from 100..600 and unrelated alternate lengths: the host order splits every class
    into equal-length duos (equal-length packed kernel), the odd ones (any-lengths packed kernel), banded references and small
    classes (one-pair kernels); every result against the oracle."""
    rng = np.random.default_rng(2024)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    lens = rng.choice(np.arange(100, 601, 7), 2400)   # 72 distinct lengths: most of them occur an odd number of times somewhere
    pairs = []
    for n in lens:
        ref = rng.choice(acgt, int(n))
        alt = ref.copy()
        for _ in range(int(rng.integers(0, 4))):
            p = int(rng.integers(5, len(alt) - 5))
            if rng.random() < 0.5:
                alt = np.delete(alt, slice(p, p + int(rng.integers(1, 20))))
            else:
                alt = np.insert(alt, p, rng.choice(acgt, int(rng.integers(1, 20))))
        if rng.random() < 0.3:
            alt = alt[int(rng.integers(0, 30)):len(alt) - int(rng.integers(0, 30))]
        pairs.append((ref, alt))
    for strategy, params in (("SoftClip", SC.NEW_SW_PARAMETERS), ("LeadingInDel", SC.STANDARD_NGS), ("InDel", SC.ALIGNMENT_TO_BEST_HAPLOTYPE)):
        _compare(ctx, pairs, params, strategy, max_ops=512)
    ro, refs, ao, alts = SC.pack(pairs)
    r = ctx.sw_upload(ro, refs, ao, alts, SC.NEW_SW_PARAMETERS, "InDel", max_ops=512)
    r.run()
    in_classes = int(((lens > 128) & (lens <= 544)).sum())
    assert r.packed_pairs >= in_classes - 40, (r.packed_pairs, in_classes)  # all but odd last ones and short tails that ride along unpacked
    r.free()
