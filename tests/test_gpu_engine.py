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


def _oracle_engine(batch, mapq, ref_idx, gcp, mismap, pcr, bq_thr, dynamic, scale, err_rate, symmetric, no_mapq_cap):
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
    lk = O.pairhmm_batch(d)
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
