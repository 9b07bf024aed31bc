"""BASELINE.json's configurations at FULL size (-m gpu).  The oracle cannot cover 4.6e11 cells, so these tests check
size-independent properties over the whole output and the oracle on random samples of it:
  cfg#2  10 000 regions x (64 reads x 150 bp, 16 haps x ~300 bp)
  cfg#3  long reads, 8 x 10 kbp vs 4 x 12 kbp per region (two full waves of the band kernel; one region against the oracle)
  cfg#4  100 000 Smith-Waterman pairs of 520 x 520 with CIGAR traceback."""
import numpy as np
import pytest

from oracle import lk_oracle as O

pytestmark = pytest.mark.gpu
REL_TOL = 1e-4  # north_star: log-likelihoods within 1e-4 relative


@pytest.fixture(scope="module")
def ctx():
    import lorikeet_b200 as lk
    c = lk.Context(0)
    yield c
    c.close()


def rel_err(a, b):
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-3)


def test_cfg2_full_size_properties_and_sampled_oracle(ctx):
    from lorikeet_b200 import synth
    n = 10000
    b = synth.pairhmm_regions(n)
    out, flags = ctx.pairhmm_batch(b, want_flags=True)
    assert out.shape == (n * 64 * 16,)
    # a likelihood is a probability: finite and <= 0 in log10 space (pair_hmm.rs:478-481 asserts the same)
    assert np.all(np.isfinite(out)) and out.max() <= 1e-9
    assert flags.mean() < 0.01  # the fp32 path carries the batch; only a sliver is redone in f64
    # every read was sampled from one haplotype of its region: its best likelihood must be far above its worst
    lk3 = out.reshape(n, 64, 16)
    assert np.mean(lk3.max(axis=2) - lk3.min(axis=2) > 1.0) > 0.9
    # sampled oracle: 16 random regions of the 10 000
    rng = np.random.default_rng(20)
    for g in rng.choice(n, 16, replace=False):
        exp = O.pairhmm_batch(b.region_slice(int(g), int(g) + 1))
        assert np.max(rel_err(lk3[g].ravel(), exp)) < REL_TOL, g
    # batch-composition invariance: a slice computed on its own is bit-identical to the same slice of the big batch
    alone = ctx.pairhmm_batch(b.region_slice(5000, 5064))
    assert np.array_equal(alone, out[5000 * 1024:5064 * 1024])
    # checksum of checksums: per-region sums are reproduced by a second, differently chunked pass (resident path)
    res = ctx.pairhmm_upload(b)
    res.run()
    again = res.download()
    res.free()
    assert np.array_equal(again.reshape(n, -1).sum(axis=1), lk3.reshape(n, -1).sum(axis=1))


def test_cfg3_long_reads_two_waves(ctx):
    from lorikeet_b200 import synth
    b = synth.pairhmm_regions(111, **synth.CFG3)
    out, flags = ctx.pairhmm_batch(b, want_flags=True)
    assert np.all(np.isfinite(out)) and out.max() <= 1e-9
    assert int((flags & 1).sum()) == 0  # nothing fell back to the f64 pass
    # one whole region (32 pairs of 10 kbp x 12 kbp) against the rescaled f64 oracle
    exp = O.pairhmm_batch(b.region_slice(7, 8), rescaled=True)
    assert np.max(rel_err(out[7 * 32:8 * 32], exp)) < REL_TOL


def test_cfg3_full_size_2000_regions(ctx):
    """BASELINE cfg#3 at its full size (2 000 regions x 8 reads x 10 kbp vs 4 haps x 12 kbp = 7.68e12 cells) through the
    host-buffer call; 8 random regions (256 pairs) against the rescaled f64 oracle, whole-output properties on the rest."""
    from lorikeet_b200 import synth, PairHmmBatch
    n = 2000
    b = synth.pairhmm_regions(n, **synth.CFG3)
    assert b.n_pairs == n * 32 and b.n_cells > 7.6e12
    out, flags = ctx.pairhmm_batch(b, want_flags=True)
    assert np.all(np.isfinite(out)) and out.max() <= 1e-9
    assert int((flags & 1).sum()) == 0  # nothing fell back to the f64 pass
    # the 111-region slice computed on its own is bit-identical to the same slice of the big batch
    alone = ctx.pairhmm_batch(b.region_slice(1000, 1020))
    assert np.array_equal(alone, out[1000 * 32:1020 * 32])
    # 10 kbp reads with 3 % indel noise at Q7-Q20 score around -2300 against their own region's haplotypes; a read against
    # an unrelated 12 kbp sequence would be below -15000
    per_read = out.reshape(n, 8, 4)
    assert per_read.max(axis=2).min() > -4000.0 and per_read.min() > -6000.0
    # oracle on 8 random regions; one oracle task per READ (a region re-packed as 8 single-read regions) so the host cores
    # share the 8 x 32 pairs of 1.2e8 cells each
    rng = np.random.default_rng(33)
    picks = sorted(int(g) for g in rng.choice(n, 8, replace=False))
    parts = []
    for g in picks:
        reg = b.region_slice(g, g + 1)
        for r in range(reg.n_reads):
            o0, o1 = int(reg.read_off[r]), int(reg.read_off[r + 1])
            parts.append(PairHmmBatch([0, 1], [0, reg.n_haps], [0, o1 - o0], reg.read_bases[o0:o1], reg.read_qual[o0:o1],
                                      reg.read_ins[o0:o1], reg.read_del[o0:o1], reg.hap_off, reg.hap_bases, gcp_constant=10))
    from lorikeet_b200.dispatch import concat_batches
    exp = O.pairhmm_batch(concat_batches(parts), rescaled=True).reshape(len(picks), 32)
    got = np.stack([out[g * 32:(g + 1) * 32] for g in picks])
    assert np.max(rel_err(got.ravel(), exp.ravel())) < REL_TOL


def test_cfg4_full_size_cigars(ctx):
    from lorikeet_b200 import synth
    n = 100000
    ro, refs, ao, alts = synth.sw_pairs(n)
    params = [200, -150, -260, -11]
    ops, n_ops, offs = ctx.sw_batch(ro, refs, ao, alts, params, "SoftClip", max_ops=64)
    assert np.all(n_ops > 0)
    op = ops & 0xF
    ln = (ops >> 4).astype(np.int64)
    valid = np.arange(64)[None, :] < n_ops[:, None]
    assert np.all(np.isin(op[valid], [0, 1, 2, 4]))  # M I D S only
    # a CIGAR consumes exactly the alternate; offset + reference-consuming ops stay inside the reference
    read_len = np.where(valid & np.isin(op, [0, 1, 4]), ln, 0).sum(axis=1)
    ref_len = np.where(valid & np.isin(op, [0, 2]), ln, 0).sum(axis=1)
    assert np.array_equal(read_len, np.diff(ao))
    assert np.all(offs >= 0) and np.all(offs + ref_len <= np.diff(ro))
    # no two adjacent elements share an operator (the builder merges them)
    same = (op[:, 1:] == op[:, :-1]) & valid[:, 1:]
    assert not same.any()
    # sampled oracle: 2 000 random pairs bit-exact
    rng = np.random.default_rng(4)
    pick = np.sort(rng.choice(n, 2000, replace=False))
    sro = np.concatenate([[0], np.cumsum(np.diff(ro)[pick])]); sao = np.concatenate([[0], np.cumsum(np.diff(ao)[pick])])
    srefs = np.concatenate([refs[ro[p]:ro[p + 1]] for p in pick]); salts = np.concatenate([alts[ao[p]:ao[p + 1]] for p in pick])
    eops, en, eoffs = O.sw_batch(sro, srefs, sao, salts, params, "SoftClip", max_ops=64)
    assert np.array_equal(n_ops[pick], en) and np.array_equal(offs[pick], eoffs)
    v = np.arange(64)[None, :] < en[:, None]
    assert np.array_equal(ops[pick][v], eops[v])
    # idempotence: a sequence aligned to itself is one match element at offset 0
    ops2, n2, off2 = ctx.sw_batch(ro[:1001], refs[:ro[1000]], ro[:1001], refs[:ro[1000]], params, "SoftClip", max_ops=8)
    assert np.all(n2 == 1) and np.all(off2 == 0) and np.array_equal(ops2[:, 0] >> 4, np.diff(ro[:1001]))
