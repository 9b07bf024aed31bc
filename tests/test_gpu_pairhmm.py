"""GPU parity tests (run with -m gpu on a B200): the CUDA PairHMM path, called through the C ABI,
against the CPU oracle on identical inputs.  Tolerances follow BASELINE.json's north_star:
log10 likelihoods within 1e-4 relative; the golden file within the reference's own 1e-5 absolute
(tests/vector_pair_hmm_unit_tests.rs:63)."""
import numpy as np
import pytest

import pairhmm_cases as PC
from conftest import load_pairhmm_golden
from oracle import lk_oracle as O

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4  # north_star: "within 1e-4 relative on the returned log-likelihood matrix"


@pytest.fixture(scope="module")
def ctx():
    import lorikeet_b200 as lk
    c = lk.Context(0)
    yield c
    c.close()


def _batch(d, **kw):
    from lorikeet_b200 import PairHmmBatch
    return PairHmmBatch.from_dict(d, **kw)


def rel_err(got, exp):
    return np.abs(got - exp) / np.maximum(np.abs(exp), 1e-3)


def test_native_library_loaded(ctx):
    assert "sm_100a" in ctx.version()


def test_golden_104_rows_batch(ctx):
    rows = load_pairhmm_golden()
    cases = [dict(hap=r["hap"], read=r["read"], qual=r["qual"], ins=r["ins"], dele=r["dele"], gcp=r["gcp"]) for r in rows]
    got, flags = ctx.pairhmm_batch(_batch(PC.pack_cases(cases)), want_flags=True)
    exp = np.array([r["expected"] for r in rows])
    assert np.max(np.abs(got - exp)) < 1e-5
    assert np.all(flags == 0)  # none of the golden rows needs the f64 pass


def test_golden_rows_single_pair_entry_point(ctx):
    # the gkl::pairhmm::forward()-shaped wrapper, exactly as tests/vector_pair_hmm_unit_tests.rs:51-63 calls it
    for r in load_pairhmm_golden()[::13]:
        got = ctx.pairhmm_forward(r["hap"], r["read"], r["qual"], r["ins"], r["dele"], r["gcp"])
        assert abs(got - r["expected"]) < 1e-5


def test_golden_rows_through_the_pairhmm_type(ctx):
    # the scalar arm's loop of PairHMM::compute_log10_likelihoods (pair_hmm.rs:290-333): for read { for allele { ... } } with
    # recache on the first haplotype and the next haplotype passed for the prefix bookkeeping
    from lorikeet_b200.engine import PairHMM
    rows = load_pairhmm_golden()[:24]
    hmm = PairHMM.initialize(ctx, [r["hap"] for r in rows], {0: [r["read"] for r in rows]})
    assert hmm.initialized and hmm.max_haplotype_length == max(len(r["hap"]) for r in rows)
    for r in rows[::6]:
        first = True
        for a, h in enumerate(rows[:6]):
            nxt = rows[a + 1]["hap"] if a + 1 < 6 else None
            got = hmm.compute_read_likelihood_given_haplotype_log10(h["hap"], r["read"], r["qual"], r["ins"], r["dele"], r["gcp"], first, nxt)
            first = False
            exp = O.pairhmm_forward(h["hap"], r["read"], r["qual"], r["ins"], r["dele"], r["gcp"])
            assert got <= 0.0 and abs(got - exp) <= REL_TOL * max(abs(exp), 1e-3)
            if h is r:
                assert abs(got - r["expected"]) < 1e-5


def test_golden_rows_forced_f64(ctx):
    rows = load_pairhmm_golden()
    cases = [dict(hap=r["hap"], read=r["read"], qual=r["qual"], ins=r["ins"], dele=r["dele"], gcp=r["gcp"]) for r in rows]
    d = PC.pack_cases(cases)
    ctx.set_force_f64(True)
    try:
        got, flags = ctx.pairhmm_batch(_batch(d), want_flags=True)
    finally:
        ctx.set_force_f64(False)
    exp = O.pairhmm_batch(d)
    assert np.all(flags == 1)
    assert np.max(np.abs(got - exp)) < 1e-10  # same operation order as the reference's f64 path


def test_basic_likelihood_family(ctx):
    # make_basic_likelihood_tests: 22 080 analytic cases, tristate off, 0.2 abs vs closed form (pair_hmm_unit_tests.rs:99,169)
    cases = PC.basic_likelihood_cases(extensive=True)
    d = PC.pack_cases(cases)
    ctx.set_tristate_correction(False)
    try:
        got = ctx.pairhmm_batch(_batch(d))
    finally:
        ctx.set_tristate_correction(True)
    closed_form = np.array([c["expected"] for c in cases])
    assert np.all(got <= 0.0)
    assert np.max(np.abs(got - closed_form)) <= 0.2
    exp = O.pairhmm_batch(d, use_tristate=False)
    assert np.max(rel_err(got, exp)) < REL_TOL


def test_other_reference_families(ctx):
    cases = PC.mismatch_sweep_cases() + PC.all_matching_cases() + PC.big_read_cases() + PC.haplotype_indexing_cases()[::7]
    d = PC.pack_cases(cases)
    ctx.set_tristate_correction(False)
    try:
        got = ctx.pairhmm_batch(_batch(d))
    finally:
        ctx.set_tristate_correction(True)
    exp = O.pairhmm_batch(d, use_tristate=False)
    assert np.all(np.isfinite(got)) and np.all(got <= 0.0)
    assert np.max(rel_err(got, exp)) < REL_TOL
    for g, c in zip(got, cases):
        if c["expected"] is not None:
            assert abs(g - c["expected"]) <= c["tol"], c["name"]


def test_synthetic_regions_cfg2(ctx):
    # cfg#2 shape at oracle-friendly size: 48 regions x (64 reads x 150 bp, 16 haps x ~300 bp)
    from lorikeet_b200 import synth
    b = synth.pairhmm_regions(48)
    got, flags = ctx.pairhmm_batch(b, want_flags=True)
    exp = O.pairhmm_batch(b)
    assert got.shape == exp.shape == (48 * 64 * 16,)
    assert np.max(rel_err(got, exp)) < REL_TOL
    assert flags.mean() < 0.2


def test_resident_matches_host_path_and_is_repeatable(ctx):
    from lorikeet_b200 import synth
    b = synth.pairhmm_regions(32, first_region=1000)
    host = ctx.pairhmm_batch(b)
    res = ctx.pairhmm_upload(b)
    ms1 = res.run()
    a = res.download()
    ms2 = res.run()
    c = res.download()
    res.free()
    assert ms1 > 0 and ms2 > 0
    assert np.array_equal(a, host) and np.array_equal(a, c)


def test_ragged_regions_mixed_lengths(ctx):
    rng = np.random.default_rng(17)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    rrs, rhs, roff, hoff = [0], [0], [0], [0]
    rb, rq, ri, rd, hb = [], [], [], [], []
    for g in range(40):
        base = rng.choice(acgt, int(rng.integers(60, 700)))
        nh = int(rng.integers(1, 9))
        for _ in range(nh):
            h = base.copy()
            h[rng.integers(0, len(h), 3)] = rng.choice(acgt, 3)
            hb.append(h); hoff.append(hoff[-1] + len(h))
        rhs.append(rhs[-1] + nh)
        nr = int(rng.integers(0, 30))
        for _ in range(nr):
            n = int(rng.integers(1, min(len(base), 400)))
            s = int(rng.integers(0, len(base) - n + 1))
            r = base[s:s + n].copy()
            if n > 20 and rng.random() < 0.3:
                r[rng.integers(0, n)] = ord("N")
            rb.append(r); rq.append(rng.integers(2, 42, n).astype(np.uint8))
            ri.append(rng.integers(20, 46, n).astype(np.uint8)); rd.append(rng.integers(20, 46, n).astype(np.uint8))
            roff.append(roff[-1] + n)
        rrs.append(rrs[-1] + nr)
    d = dict(region_read_start=np.array(rrs), region_hap_start=np.array(rhs), read_off=np.array(roff),
             hap_off=np.array(hoff), read_bases=np.concatenate(rb), read_qual=np.concatenate(rq),
             read_ins=np.concatenate(ri), read_del=np.concatenate(rd), read_gcp=np.full(roff[-1], 10, np.uint8),
             hap_bases=np.concatenate(hb))
    got = ctx.pairhmm_batch(_batch(d))
    exp = O.pairhmm_batch(d)
    assert np.max(rel_err(got, exp)) < REL_TOL


def test_variable_gcp_exotic_bases_and_rescue(ctx):
    rng = np.random.default_rng(11)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    hap = rng.choice(acgt, 190)
    read = hap[10:160].copy()
    n = len(read)
    gcp = rng.integers(5, 20, n).astype(np.uint8)
    c1 = dict(hap=hap, read=read, qual=np.full(n, 30, np.uint8), ins=np.full(n, 40, np.uint8),
              dele=np.full(n, 42, np.uint8), gcp=gcp)
    hap2 = hap.copy(); hap2[30] = ord("X"); read2 = read.copy(); read2[20] = ord("X")
    c2 = dict(c1, hap=hap2, read=read2)                               # exotic bytes -> byte-exact f64 pass
    c3 = dict(c1, read=rng.choice(acgt, n), qual=np.full(n, 40, np.uint8))  # unrelated read: fp32 underflow -> f64
    qz = np.full(n, 30, np.uint8); qz[5] = 0
    c4 = dict(c1, qual=qz)                                            # Q0 base: outside the hat form's domain
    d = PC.pack_cases([c1, c2, c3, c4])
    got, flags = ctx.pairhmm_batch(_batch(d), want_flags=True)
    exp = O.pairhmm_batch(d)
    assert list(flags) == [0, 1, 1, 1]
    assert np.max(rel_err(got, exp)) < REL_TOL
    assert np.max(np.abs(got[1:] - exp[1:])) < 1e-9


def test_long_reads_go_through_band_pass(ctx):
    from lorikeet_b200 import synth
    b = synth.pairhmm_regions(2, **dict(synth.CFG3, n_reads=2, n_haps=2, read_len=2500, hap_len=3000))
    got, flags = ctx.pairhmm_batch(b, want_flags=True)
    exp = O.pairhmm_batch(b, rescaled=True)
    assert np.all(np.isfinite(got))
    assert np.max(rel_err(got, exp)) < REL_TOL


def test_invalid_inputs_are_rejected(ctx):
    import lorikeet_b200 as lk
    d = PC.pack_cases([dict(hap=b"ACGT", read=b"ACG", qual=np.full(3, 30, np.uint8), ins=np.full(3, 40, np.uint8),
                            dele=np.full(3, 40, np.uint8), gcp=np.full(3, 10, np.uint8))])
    d["read_off"] = np.array([0, 0])  # empty read: the reference asserts on it (pair_hmm.rs:417-440)
    with pytest.raises(lk.LorikeetError):
        ctx.pairhmm_batch(_batch(d))
    empty = _batch(dict(region_read_start=np.array([0]), region_hap_start=np.array([0]), read_off=np.array([0]),
                        hap_off=np.array([0]), read_bases=b"", read_qual=b"", read_ins=b"", read_del=b"", hap_bases=b""))
    assert ctx.pairhmm_batch(empty).shape == (0,)


def test_chunked_host_pipeline_matches_resident(ctx):
    # > 2M pairs: lk_pairhmm_batch_host cuts the batch into chunks over two streams; the result must equal the
    # single-workspace resident path bit for bit, and the oracle on a sample of regions
    from lorikeet_b200 import synth
    b = synth.pairhmm_regions(2200, first_region=20000)
    host, hflags = ctx.pairhmm_batch(b, want_flags=True)
    res = ctx.pairhmm_upload(b)
    res.run()
    dev, dflags = res.download(want_flags=True)
    res.free()
    assert np.array_equal(host, dev) and np.array_equal(hflags, dflags)
    for g in (0, 1099, 2199):
        exp = O.pairhmm_batch(b.region_slice(g, g + 1))
        assert np.max(rel_err(host[g * 1024:(g + 1) * 1024], exp)) < REL_TOL
    # a second call reuses the context's workspaces
    again = ctx.pairhmm_batch(b)
    assert np.array_equal(again, host)
