"""GPU parity tests (-m gpu) of the engine-level entry point: compute_read_likelihoods
(pair_hmm_likelihood_calculation_engine.rs:195-242) = modify_read_qualities -> PairHMM -> normalize -> filter,
against the same composition of oracle functions.  Mirrors tests/pair_hmm_likelihood_calculation_engine_unit_tests.rs
and the post-op tests of tests/allele_likelihoods_unit_tests.rs."""
import ctypes as C

import numpy as np
import pytest

from oracle import lk_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import lorikeet_b200 as lk
    c = lk.Context(0)
    yield c
    c.close()


def _oracle_engine(batch, mapq, ref_idx, gcp, mismap, pcr, bq_thr, dynamic, scale, err_rate, symmetric, no_mapq_cap, rescaled=False):
    L = O.lib()
    p = lambda a, t: a.ctypes.data_as(t)
    q, ins, dele = batch.read_qual.copy(), batch.read_ins.copy(), batch.read_del.copy()
    for r in range(batch.n_reads):
        a, b = int(batch.read_off[r]), int(batch.read_off[r + 1])
        bb = np.ascontiguousarray(batch.read_bases[a:b]); qq = np.ascontiguousarray(q[a:b])
        ii = np.ascontiguousarray(ins[a:b]); dd = np.ascontiguousarray(dele[a:b])
        L.lko_apply_pcr_error_model(p(bb, O.u8p), b - a, p(ii, O.u8p), p(dd, O.u8p), pcr)
        L.lko_cap_minimum_read_qualities(p(qq, O.u8p), p(ii, O.u8p), p(dd, O.u8p), b - a, int(mapq[r]), bq_thr, no_mapq_cap)
        q[a:b], ins[a:b], dele[a:b] = qq, ii, dd
    d = dict(region_read_start=batch.region_read_start, region_hap_start=batch.region_hap_start, read_off=batch.read_off,
             hap_off=batch.hap_off, read_bases=batch.read_bases, read_qual=q, read_ins=ins, read_del=dele,
             read_gcp=np.full(batch.read_bases.shape, gcp, np.uint8), hap_bases=batch.hap_bases)
    lk = O.pairhmm_batch(d, rescaled=rescaled)
    values = np.empty_like(lk)
    keep = np.ones(batch.n_reads, np.uint8)
    base = 0
    for g in range(batch.n_regions):
        r0, r1 = int(batch.region_read_start[g]), int(batch.region_read_start[g + 1])
        R, A = r1 - r0, int(batch.region_hap_start[g + 1] - batch.region_hap_start[g])
        v = np.ascontiguousarray(lk[base:base + R * A].reshape(R, A).T)  # [a][r]
        L.lko_normalize_likelihoods(p(v, O.f64p), A, R, mismap, 1 if symmetric else 0, int(ref_idx[g]))
        thr = np.empty(R)
        for r in range(R):
            a, b = int(batch.read_off[r0 + r]), int(batch.read_off[r0 + r + 1])
            oq = np.ascontiguousarray(batch.read_qual[a:b])  # ORIGINAL quals / length (the "HMMQuals" quirk)
            stat_capped = L.lko_log10_min_true_likelihood(b - a, err_rate, 1)
            if dynamic:
                dyn = L.lko_log10_dynamic_read_qual_threshold(p(oq, O.u8p), b - a, scale)
                stat = L.lko_log10_min_true_likelihood(b - a, err_rate, 0)
                thr[r] = min(dyn, stat)
            else:
                thr[r] = stat_capped
        k = np.ones(R, np.uint8)
        L.lko_filter_poorly_modeled_evidence(p(v, O.f64p), A, R, p(thr, O.f64p), p(k, O.u8p))
        values[base:base + R * A] = v.ravel()
        keep[r0:r1] = k
        base += R * A
    return values, keep


@pytest.mark.parametrize("dynamic", [False, True])
@pytest.mark.parametrize("symmetric", [True, False])
def test_compute_read_likelihoods_matches_oracle_composition(ctx, dynamic, symmetric):
    from lorikeet_b200 import synth
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine, PCRErrorModel
    # raw reads: flat Q45 indel quals, no pre-ops applied (pcr_model 0 and threshold 0 in the generator)
    b = synth.pairhmm_regions(24, first_region=777, pcr_model=0, base_quality_score_threshold=0)
    rng = np.random.default_rng(1)
    # make some reads poorly modelled so the filter has something to remove
    for r in rng.choice(b.n_reads, 60, replace=False):
        a, e = int(b.read_off[r]), int(b.read_off[r + 1])
        b.read_bases[a:e] = rng.choice(np.frombuffer(b"ACGT", np.uint8), e - a)
    mapq = rng.integers(20, 61, b.n_reads).astype(np.uint8)
    ref_idx = rng.integers(0, 3, b.n_regions).astype(np.int32)
    eng = PairHMMLikelihoodCalculationEngine(ctx, 10, -4.5, PCRErrorModel.CONSERVATIVE, 18, dynamic, 1.0, 0.02, symmetric, False, False)
    values, keep = eng.compute_read_likelihoods(b, mapq, ref_idx)
    evalues, ekeep = _oracle_engine(b, mapq, ref_idx, 10, -4.5, 3, 18, dynamic, 1.0, 0.02, symmetric, 0)
    assert np.array_equal(keep, ekeep)
    assert 0 < (keep == 0).sum() < b.n_reads
    nan = np.isnan(evalues)
    assert np.array_equal(np.isnan(values), nan)
    rel = np.abs(values[~nan] - evalues[~nan]) / np.maximum(np.abs(evalues[~nan]), 1e-3)
    assert rel.max() < 1e-4


def test_reference_engine_unit_test(ctx):
    # tests/pair_hmm_likelihood_calculation_engine_unit_tests.rs:22-88: 1 read, 2 haplotypes (matching / 1 SNP),
    # gcp 93, mismap Q45 (-4.5), Conservative PCR, bq threshold 16, static disqualification: matching hap must score higher
    import lorikeet_b200 as lk
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine, PCRErrorModel
    read = b"ACGTACGTAC"
    hap1, hap2 = b"ACGTACGTACT", b"ACGTACCTACT"
    n = len(read)
    b = lk.PairHmmBatch([0, 1], [0, 2], [0, n], read, np.full(n, 30, np.uint8), np.full(n, 45, np.uint8), np.full(n, 45, np.uint8),
                        [0, len(hap1), len(hap1) + len(hap2)], hap1 + hap2)
    eng = PairHMMLikelihoodCalculationEngine(ctx, 93, -4.5, PCRErrorModel.CONSERVATIVE, 16, False, 1.0, 0.02, True, False, False)
    values, keep = eng.compute_read_likelihoods(b, np.array([60], np.uint8), np.array([0], np.int32))
    assert keep[0] == 1 and values[0] > values[1]


@pytest.mark.parametrize("pcr,no_mapq_cap,bq_thr", [(0, False, 18), (1, False, 6), (2, True, 18), (3, True, 30)])
def test_pre_ops_variants_match_oracle(ctx, pcr, no_mapq_cap, bq_thr):
    # every PCRErrorModel (None / Hostile / Aggressive / Conservative), with and without the MAPQ cap, several thresholds:
    # the device pre-ops (engine_modify_reads_kernel) against the oracle's apply_pcr_error_model + cap_minimum_read_qualities
    from lorikeet_b200 import synth
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine
    b = synth.pairhmm_regions(10, first_region=4242 + pcr, pcr_model=0, base_quality_score_threshold=0)
    rng = np.random.default_rng(pcr)
    # plant homopolymers / short tandem repeats so the PCR model has repeats of many lengths to cap
    for r in rng.choice(b.n_reads, 80, replace=False):
        a, e = int(b.read_off[r]), int(b.read_off[r + 1])
        unit = rng.choice(np.frombuffer(b"ACGT", np.uint8), int(rng.integers(1, 5)))
        k = int(rng.integers(a, e - 40))
        b.read_bases[k:k + 36] = np.resize(unit, 36)
    mapq = rng.integers(5, 61, b.n_reads).astype(np.uint8)
    ref_idx = np.zeros(b.n_regions, np.int32)
    eng = PairHMMLikelihoodCalculationEngine(ctx, 10, -4.5, pcr, bq_thr, False, 1.0, 0.02, True, no_mapq_cap, False)
    values, keep = eng.compute_read_likelihoods(b, mapq, ref_idx)
    evalues, ekeep = _oracle_engine(b, mapq, ref_idx, 10, -4.5, pcr, bq_thr, False, 1.0, 0.02, True, 1 if no_mapq_cap else 0)
    assert np.array_equal(keep, ekeep)
    nan = np.isnan(evalues)
    assert np.array_equal(np.isnan(values), nan)
    rel = np.abs(values[~nan] - evalues[~nan]) / np.maximum(np.abs(evalues[~nan]), 1e-3)
    assert rel.max() < 1e-4


def test_region_driver_is_chunk_invariant(ctx):
    # 2600 regions = 2.66 M pairs -> the host pipeline cuts the batch into 2 chunks (two workspaces, two streams); the
    # result must equal the concatenation of independent calls on thirds of the batch, bit for bit
    from lorikeet_b200 import synth
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine, PCRErrorModel
    n = 2600
    b = synth.pairhmm_regions(n, first_region=90000, pcr_model=0, base_quality_score_threshold=0)
    rng = np.random.default_rng(3)
    mapq = rng.integers(20, 61, b.n_reads).astype(np.uint8)
    ref_idx = rng.integers(0, 2, b.n_regions).astype(np.int32)
    eng = PairHMMLikelihoodCalculationEngine(ctx, 10, -4.5, PCRErrorModel.CONSERVATIVE, 18, True, 1.0, 0.02, True, False, False)
    values, keep = eng.compute_read_likelihoods(b, mapq, ref_idx)
    parts_v, parts_k = [], []
    for g0, g1 in ((0, 900), (900, 1800), (1800, n)):
        sb = b.region_slice(g0, g1)
        r0, r1 = int(b.region_read_start[g0]), int(b.region_read_start[g1])
        v, k = eng.compute_read_likelihoods(sb, mapq[r0:r1], ref_idx[g0:g1])
        parts_v.append(v); parts_k.append(k)
    assert np.array_equal(keep, np.concatenate(parts_k))
    pv = np.concatenate(parts_v)
    assert np.array_equal(np.isnan(values), np.isnan(pv))
    assert np.array_equal(values[~np.isnan(values)], pv[~np.isnan(pv)])


def test_region_driver_on_long_reads(ctx):
    # reads longer than the modify kernel's shared-memory slab (1024 bases) and than the fast kernel's 32*K rows: the
    # pre-ops scan global memory and the likelihoods come from the fp32 band kernel; oracle = rescaled f64 restatement
    from lorikeet_b200 import synth
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine, PCRErrorModel
    b = synth.pairhmm_regions(3, **dict(synth.CFG3, read_len=1500, hap_len=1800, n_reads=4, n_haps=3, pcr_model=0,
                                         base_quality_score_threshold=0))
    rng = np.random.default_rng(9)
    for r in range(b.n_reads):  # plant a long homopolymer and a dinucleotide repeat in every read
        a = int(b.read_off[r])
        b.read_bases[a + 1100:a + 1140] = ord("A")
        b.read_bases[a + 300:a + 340] = np.resize(np.frombuffer(b"CA", np.uint8), 40)
    mapq = np.full(b.n_reads, 50, np.uint8)
    ref_idx = np.zeros(b.n_regions, np.int32)
    # dynamic disqualification: with the static -8 every noisy 1.5 kbp read would be filtered and nothing would be compared
    eng = PairHMMLikelihoodCalculationEngine(ctx, 10, -4.5, PCRErrorModel.CONSERVATIVE, 18, True, 4.0, 0.3, True, False, False)
    values, keep = eng.compute_read_likelihoods(b, mapq, ref_idx)
    # oracle composition with the rescaled forward (the un-rescaled f64 underflows on long noisy reads)
    evalues, ekeep = _oracle_engine(b, mapq, ref_idx, 10, -4.5, 3, 18, True, 4.0, 0.3, True, 0, rescaled=True)
    assert np.array_equal(keep, ekeep)
    assert keep.sum() > 0
    nan = np.isnan(evalues)
    assert np.array_equal(np.isnan(values), nan)
    rel = np.abs(values[~nan] - evalues[~nan]) / np.maximum(np.abs(evalues[~nan]), 1e-3)
    assert rel.max() < 1e-4


def test_device_pre_ops_are_byte_identical_to_the_host_utility(ctx):
    # engine_modify_reads_kernel (shared-memory path up to 1024 bases, global-memory scan above) against
    # lk_engine_modify_read_qualities on reads built to be hard: homopolymers, tandem repeats of every unit length 1..22,
    # repeats touching either end, reads of length 1..3, lengths around the 1024-base path switch
    from lorikeet_b200 import _lib
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine
    L = _lib.load()
    rng = np.random.default_rng(123)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    reads = []
    for length in [1, 2, 3, 5, 31, 32, 33, 64, 100, 150, 151, 255, 256, 257, 300, 1023, 1024, 1025, 1500] + list(rng.integers(20, 260, 300)):
        length = int(length)
        r = rng.choice(acgt, length)
        for _ in range(int(rng.integers(0, 5))):
            u = int(rng.integers(1, 23))
            reps = int(rng.integers(2, 30))
            unit = rng.choice(acgt, u)
            k = int(rng.integers(-u, max(1, length)))
            seg = np.resize(unit, u * reps)
            a, e = max(k, 0), min(k + len(seg), length)
            if e > a:
                r[a:e] = seg[a - k:e - k]
        reads.append(r)
    off = np.concatenate([[0], np.cumsum([len(r) for r in reads])]).astype(np.int64)
    bases = np.ascontiguousarray(np.concatenate(reads))
    n = len(reads)
    for pcr, no_cap, thr in ((3, 0, 18), (1, 1, 6), (2, 0, 30), (0, 0, 18)):
        q = rng.integers(2, 42, bases.size).astype(np.uint8)
        ins = rng.integers(3, 50, bases.size).astype(np.uint8)
        dele = rng.integers(3, 50, bases.size).astype(np.uint8)
        mapq = rng.integers(0, 61, n).astype(np.uint8)
        eng = PairHMMLikelihoodCalculationEngine(ctx, 10, -4.5, pcr, thr, False, 1.0, 0.02, True, bool(no_cap), False)
        hq, hi, hd = q.copy(), ins.copy(), dele.copy()
        p = lambda a, t: a.ctypes.data_as(t)
        assert L.lk_engine_modify_read_qualities(C.byref(eng.args), n, p(off, _lib.i64p), p(bases, _lib.u8p), p(hq, _lib.u8p),
                                                 p(hi, _lib.u8p), p(hd, _lib.u8p), p(mapq, _lib.u8p), 4) == 0
        dq, di, dd = q.copy(), ins.copy(), dele.copy()
        ctx._check(L.lk_engine_modify_read_qualities_device(ctx._h, C.byref(eng.args), n, p(off, _lib.i64p), p(bases, _lib.u8p),
                                                            p(dq, _lib.u8p), p(di, _lib.u8p), p(dd, _lib.u8p), p(mapq, _lib.u8p)))
        assert np.array_equal(dq, hq)
        bad = np.nonzero((di != hi) | (dd != hd))[0]
        assert bad.size == 0, (pcr, bad[:5], di[bad[:5]], hi[bad[:5]])
