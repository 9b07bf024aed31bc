"""GPU tests (-m gpu) of several samples per region: the reference computes one [A x R_s] matrix per sample over the
region's shared haplotypes and normalises / filters / realigns each on its own
(pair_hmm_likelihood_calculation_engine.rs:201-239,321-340; assembly_based_caller_utils.rs:208-246).  A multi-sample batch
must equal (a) the oracle composition applied per sample and (b) bit for bit, the current single-sample path on the
FLATTENED equivalent (every sample as a region of its own with a copy of the haplotypes)."""
import numpy as np
import pytest

import lorikeet_b200 as lk
from oracle import lk_oracle as O
from test_gpu_engine import _oracle_engine

pytestmark = pytest.mark.gpu


def _same(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


def _sampled_batch(n_regions=12, seed=3):
    """cfg2-style raw regions, reads split into 3 samples of uneven size (one of them sometimes empty)."""
    from lorikeet_b200 import synth
    b = synth.pairhmm_regions(n_regions, first_region=4242, pcr_model=0, base_quality_score_threshold=0)
    rng = np.random.default_rng(seed)
    for r in rng.choice(b.n_reads, 40, replace=False):  # poorly modelled reads, so the filter removes some
        a, e = int(b.read_off[r]), int(b.read_off[r + 1])
        b.read_bases[a:e] = rng.choice(np.frombuffer(b"ACGT", np.uint8), e - a)
    rss, srs = [0], [0]
    for g in range(n_regions):
        r0, r1 = int(b.region_read_start[g]), int(b.region_read_start[g + 1])
        cuts = sorted(int(x) for x in rng.integers(r0, r1 + 1, 2))
        if g == 5:
            cuts = [r0, cuts[1]]  # an empty first sample
        for c in cuts + [r1]:
            srs.append(c)
        rss.append(rss[-1] + 3)
    b.set_samples(rss, srs)
    return b


def _flatten(b):
    """Every sample as its own region with a copy of its region's haplotypes."""
    from lorikeet_b200.dispatch import concat_batches
    parts, region_of = [], []
    for g in range(b.n_regions):
        reg = b.region_slice(g, g + 1)
        for s in range(int(reg.region_sample_start[1])):
            q0, q1 = int(reg.sample_read_start[s]), int(reg.sample_read_start[s + 1])
            o0, o1 = int(reg.read_off[q0]), int(reg.read_off[q1])
            parts.append(lk.PairHmmBatch([0, q1 - q0], [0, reg.n_haps], reg.read_off[q0:q1 + 1] - o0, reg.read_bases[o0:o1],
                                         reg.read_qual[o0:o1], reg.read_ins[o0:o1], reg.read_del[o0:o1], reg.hap_off, reg.hap_bases,
                                         gcp_constant=10))
            region_of.append(g)
    return concat_batches(parts), np.array(region_of)


@pytest.mark.parametrize("dynamic", [False, True])
def test_engine_multi_sample_equals_flattened_and_oracle(dynamic):
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine, PCRErrorModel
    b = _sampled_batch()
    flat, region_of = _flatten(b)
    assert flat.n_reads == b.n_reads and flat.n_pairs == b.n_pairs
    rng = np.random.default_rng(9)
    mapq = rng.integers(20, 61, b.n_reads).astype(np.uint8)
    ref_idx = rng.integers(0, 3, b.n_regions).astype(np.int32)
    with lk.Context(0) as ctx:
        eng = PairHMMLikelihoodCalculationEngine(ctx, 10, -4.5, PCRErrorModel.CONSERVATIVE, 18, dynamic, 1.0, 0.02, False, False, False)
        v, keep = eng.compute_read_likelihoods(b, mapq, ref_idx)
        fv, fkeep = eng.compute_read_likelihoods(flat, mapq, ref_idx[region_of])
        # the same batch WITHOUT sample arrays compacts per region instead: the two differ as soon as a read is filtered
        b1 = lk.PairHmmBatch(b.region_read_start, b.region_hap_start, b.read_off, b.read_bases, b.read_qual, b.read_ins, b.read_del,
                             b.hap_off, b.hap_bases, gcp_constant=10)
        v1, keep1 = eng.compute_read_likelihoods(b1, mapq, ref_idx)
    assert _same(v, fv) and np.array_equal(keep, fkeep)
    assert np.array_equal(keep, keep1) and 0 < (keep == 0).sum() < b.n_reads and not _same(v, v1)
    ev, ekeep = _oracle_engine(flat, mapq, ref_idx[region_of], 10, -4.5, 3, 18, dynamic, 1.0, 0.02, False, 0)
    assert np.array_equal(keep, ekeep) and np.array_equal(np.isnan(v), np.isnan(ev))
    nan = np.isnan(ev)
    assert (np.abs(v[~nan] - ev[~nan]) / np.maximum(np.abs(ev[~nan]), 1e-3)).max() < 1e-4


def test_engine_rejects_samples_that_do_not_tile_the_reads():
    from lorikeet_b200 import _lib
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine
    b = _sampled_batch(4)
    b.sample_read_start[2] += 1 if b.sample_read_start[2] < b.sample_read_start[3] else -1
    b.sample_read_start[3] = b.region_read_start[1] - 1  # the last sample of region 0 stops short of the region's reads
    with lk.Context(0) as ctx:
        with pytest.raises(lk.LorikeetError) as ei:
            PairHMMLikelihoodCalculationEngine(ctx).compute_read_likelihoods(b, np.full(b.n_reads, 60, np.uint8))
        assert ei.value.status == _lib.LK_ERR_INVALID


def test_regions_without_haplotypes_filter_every_read_and_single_allele_regions_work():
    """keep[] is defined (0) for reads of a region with no haplotype, also when the whole batch holds no pair;
    a region whose only allele is the reference (A == 1, asymmetric normalisation, ref index 0) reads nothing out of bounds."""
    from lorikeet_b200 import synth
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine, PCRErrorModel
    src = synth.pairhmm_regions(3, first_region=99, pcr_model=0, base_quality_score_threshold=0)
    r = src.region_read_start
    # region 0: its reads, NO haplotypes; region 1: only the first haplotype; region 2: untouched
    h1, h2 = int(src.region_hap_start[1]), int(src.region_hap_start[2])
    keep_haps = [h1] + list(range(h2, int(src.region_hap_start[3])))
    hap_off = [0]
    hb = []
    for h in keep_haps:
        seq = src.hap_bases[int(src.hap_off[h]):int(src.hap_off[h + 1])]
        hb.append(seq); hap_off.append(hap_off[-1] + len(seq))
    b = lk.PairHmmBatch(r, [0, 0, 1, len(keep_haps)], src.read_off, src.read_bases, src.read_qual, src.read_ins, src.read_del,
                        hap_off, np.concatenate(hb), gcp_constant=10)
    mapq = np.full(b.n_reads, 60, np.uint8)
    with lk.Context(0) as ctx:
        eng = PairHMMLikelihoodCalculationEngine(ctx, 10, -4.5, PCRErrorModel.CONSERVATIVE, 18, False, 1.0, 0.02, False, False, False)
        keep = np.full(b.n_reads, 7, np.uint8)
        v, keep = eng.compute_read_likelihoods(b, mapq, np.zeros(3, np.int32), keep=keep)
        assert np.all(keep[:int(r[1])] == 0) and set(np.unique(keep[int(r[1]):])) <= {0, 1} and keep[int(r[1]):].sum() > 0
        one = b.region_slice(1, 2)
        alone_v, alone_keep = eng.compute_read_likelihoods(one, mapq[:one.n_reads], np.zeros(1, np.int32))
        n1 = one.n_pairs
        assert _same(alone_v, v[:n1]) and np.array_equal(alone_keep, keep[int(r[1]):int(r[2])])
        # a batch made of the haplotype-less region only: no pair at all, nothing is launched, keep is still written
        empty = b.region_slice(0, 1)
        k0 = np.full(empty.n_reads, 7, np.uint8)
        _, k0 = eng.compute_read_likelihoods(empty, mapq[:empty.n_reads], keep=k0)
        assert np.all(k0 == 0)


def _realign_inputs(seed=11):
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", np.uint8)
    rrs, rhs, vals, haps, hcig, hstart, reads, rcig, rstart, rss, srs, nev = [0], [0], [], [], [], [], [], [], [], [0], [0], []
    for g in range(10):
        ref = rng.choice(acgt, 200)
        k, d = int(rng.integers(40, 150)), int(rng.integers(1, 6))
        region_haps = [ref, np.concatenate([ref[:k], ref[k + d:]]), np.concatenate([ref[:k], rng.choice(acgt, d), ref[k:]])]
        region_cigs = ["200M", "%dM%dD%dM" % (k, d, 200 - k - d), "%dM%dI%dM" % (k, d, 200 - k)]
        sizes = [int(x) for x in rng.integers(0, 6, 3)]
        blocks = []
        for Rs in sizes:
            V = -rng.random((3, Rs)) * 5
            for r in range(Rs):
                a = int(np.argmax(V[:, r]))
                if r % 3:
                    V[a, r] += 1.0                      # informative
                else:
                    V[(a + 1) % 3, r] = V[a, r] - 0.05  # a near tie: the priorities decide
                h = region_haps[a]
                s = int(rng.integers(0, len(h) - 70))
                rd = h[s:s + 70].copy()
                rd[int(rng.integers(0, 70))] = rng.choice(acgt)
                reads.append(rd); rcig.append("70M")
            blocks.append(V.ravel())
            srs.append(len(reads)); nev.append(max(0, Rs - (1 if Rs > 2 else 0)))  # the last read of a larger sample is no longer evidence
        rss.append(rss[-1] + 3)
        haps += region_haps; hcig += region_cigs; hstart += [0, 0, 0]
        vals.append(np.concatenate(blocks)); rstart.append(1000 * (g + 1))
        rrs.append(len(reads)); rhs.append(len(haps))
    return dict(rrs=rrs, rhs=rhs, vals=np.concatenate(vals), haps=haps, hcig=hcig, hstart=hstart, reads=reads, rcig=rcig, rstart=rstart,
                rss=rss, srs=srs, nev=nev)


def test_realign_multi_sample_equals_flattened_and_device_values_equal_host_values():
    import torch
    d = _realign_inputs()
    multi = lk.RealignBatch(d["rrs"], d["rhs"], d["vals"], d["haps"], d["hcig"], d["hstart"], d["reads"], d["rcig"], d["rstart"],
                            region_sample_start=d["rss"], sample_read_start=d["srs"], sample_n_evidence=d["nev"])
    # flattened: every sample a region of its own (haplotypes repeated)
    f_rrs, f_rhs, f_haps, f_hcig, f_hstart, f_rstart = [0], [0], [], [], [], []
    for g in range(len(d["rrs"]) - 1):
        for s in range(3):
            sm = d["rss"][g] + s
            f_rrs.append(d["srs"][sm + 1])
            f_haps += d["haps"][d["rhs"][g]:d["rhs"][g + 1]]; f_hcig += d["hcig"][d["rhs"][g]:d["rhs"][g + 1]]
            f_hstart += d["hstart"][d["rhs"][g]:d["rhs"][g + 1]]
            f_rhs.append(len(f_haps)); f_rstart.append(d["rstart"][g])
    flat = lk.RealignBatch(f_rrs, f_rhs, d["vals"], f_haps, f_hcig, f_hstart, d["reads"], d["rcig"], f_rstart, region_n_evidence=d["nev"])
    with lk.Context(0) as ctx:
        mb, mi, mc, mp = ctx.realign_reads_batch(multi)
        fb, fi, fc, fp = ctx.realign_reads_batch(flat)
        assert np.array_equal(mb, fb) and np.array_equal(mi, fi) and mc == fc and np.array_equal(mp, fp)
        assert (mb >= 0).sum() > 10 and (mb < 0).sum() > 0
        # the same matrices left in device memory: the best-allele search runs as a kernel, nothing else changes
        dv = torch.from_numpy(d["vals"]).cuda()
        dev = lk.RealignBatch(d["rrs"], d["rhs"], None, d["haps"], d["hcig"], d["hstart"], d["reads"], d["rcig"], d["rstart"],
                              region_sample_start=d["rss"], sample_read_start=d["srs"], sample_n_evidence=d["nev"],
                              values_device_ptr=dv.data_ptr())
        db, di, dc, dp = ctx.realign_reads_batch(dev)
        assert np.array_equal(db, mb) and np.array_equal(di, mi) and dc == mc and np.array_equal(dp, mp)


def test_pool_and_batcher_carry_samples():
    import threading
    from lorikeet_b200.engine import DevicePool, PairHMMLikelihoodCalculationEngine, PCRErrorModel, RegionBatcher
    b = _sampled_batch(16, seed=4)
    mapq = np.full(b.n_reads, 60, np.uint8)
    mk = lambda c: PairHMMLikelihoodCalculationEngine(c, 10, -4.5, PCRErrorModel.CONSERVATIVE, 18, False, 1.0, 0.02, True, False, False)
    with lk.Context(0) as ctx:
        exp_v, exp_keep = mk(ctx).compute_read_likelihoods(b, mapq)
    import os
    with DevicePool() as pool:
        pv, pk = pool.compute_read_likelihoods(mk(None), b, mapq)
        assert _same(pv, exp_v) and np.array_equal(pk, exp_keep)
        # the same with the batch cut into ~4-region chunks, so that the chunk views re-base the sample arrays
        os.environ["LK_POOL_MIN_CHUNK_CELLS"] = "150000000"
        try:
            pv, pk = pool.compute_read_likelihoods(mk(None), b, mapq)
        finally:
            del os.environ["LK_POOL_MIN_CHUNK_CELLS"]
        assert _same(pv, exp_v) and np.array_equal(pk, exp_keep)
        got = [None] * b.n_regions
        with RegionBatcher(pool, mk(None), max_regions=16, max_wait_us=5000) as bt:
            def work(g):
                r0, r1 = int(b.region_read_start[g]), int(b.region_read_start[g + 1])
                reg = b.region_slice(g, g + 1)
                if g % 4 == 0:  # a caller that does not use samples at all shares launches with callers that do
                    reg.region_sample_start = reg.sample_read_start = None
                got[g] = (bt.submit(reg, mapq[r0:r1], 0), g % 4 == 0)
            th = [threading.Thread(target=work, args=(g,)) for g in range(b.n_regions)]
            [t.start() for t in th]
            [t.join() for t in th]
    with lk.Context(0) as ctx:
        pbase = 0
        for g in range(b.n_regions):
            r0, r1 = int(b.region_read_start[g]), int(b.region_read_start[g + 1])
            n = (r1 - r0) * int(b.region_hap_start[g + 1] - b.region_hap_start[g])
            (v, keep), unsampled = got[g]
            if unsampled:
                reg = b.region_slice(g, g + 1)
                reg.region_sample_start = reg.sample_read_start = None
                ev, ek = mk(ctx).compute_read_likelihoods(reg, mapq[r0:r1])
                assert _same(v, ev) and np.array_equal(keep, ek)
            else:
                assert _same(v, exp_v[pbase:pbase + n]) and np.array_equal(keep, exp_keep[r0:r1])
            pbase += n


def test_engine_to_realign_chain_keeps_the_matrices_on_the_device():
    """N2 as a resident chain: lk_engine_compute_read_likelihoods leaves `values` in HBM, lk_realign_reads_batch picks the best
    haplotype of every surviving read with a kernel on those matrices and aligns the read to it; same answers as the chain
    through host memory."""
    import torch
    from lorikeet_b200 import synth
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine, PCRErrorModel
    b = synth.pairhmm_regions(20, first_region=555, pcr_model=0, base_quality_score_threshold=0)
    rng = np.random.default_rng(2)
    for r in rng.choice(b.n_reads, 30, replace=False):  # some reads the filter removes
        a, e = int(b.read_off[r]), int(b.read_off[r + 1])
        b.read_bases[a:e] = rng.choice(np.frombuffer(b"ACGT", np.uint8), e - a)
    mapq = np.full(b.n_reads, 60, np.uint8)
    with lk.Context(0) as ctx:
        eng = PairHMMLikelihoodCalculationEngine(ctx, 10, -4.5, PCRErrorModel.CONSERVATIVE, 18, False, 1.0, 0.02, True, False, False)
        host_values, keep = eng.compute_read_likelihoods(b, mapq)
        dev_values = torch.empty(b.n_pairs, dtype=torch.float64, device="cuda")
        _, keep2 = eng.compute_read_likelihoods(b, mapq, values_device_ptr=dev_values.data_ptr())
        assert np.array_equal(keep, keep2) and _same(dev_values.cpu().numpy(), host_values)
        # realignment inputs: per region the surviving reads first (the order of the compacted matrix columns), every haplotype
        # aligned to the region's first one with an all-match CIGAR of its own length
        haps, hcig, hstart, reads, rcig, n_ev, rstart = [], [], [], [], [], [], []
        for g in range(b.n_regions):
            for h in range(int(b.region_hap_start[g]), int(b.region_hap_start[g + 1])):
                seq = b.hap_bases[int(b.hap_off[h]):int(b.hap_off[h + 1])]
                haps.append(seq); hcig.append("%dM" % len(seq)); hstart.append(0)
            r0, r1 = int(b.region_read_start[g]), int(b.region_read_start[g + 1])
            order = [r for r in range(r0, r1) if keep[r]] + [r for r in range(r0, r1) if not keep[r]]
            for r in order:
                seq = b.read_bases[int(b.read_off[r]):int(b.read_off[r + 1])]
                reads.append(seq); rcig.append("%dM" % len(seq))
            n_ev.append(int(keep[r0:r1].sum())); rstart.append(1000 * (g + 1))
        common = dict(hap_seqs=haps, hap_cigars=hcig, hap_start_wrt_ref=hstart, read_seqs=reads, read_cigars=rcig, reference_start=rstart,
                      region_n_evidence=n_ev)
        via_host = lk.RealignBatch(b.region_read_start, b.region_hap_start, host_values, **common)
        via_dev = lk.RealignBatch(b.region_read_start, b.region_hap_start, None, values_device_ptr=dev_values.data_ptr(), **common)
        hb, hi, hc, hp = ctx.realign_reads_batch(via_host)
        db, di, dc, dp = ctx.realign_reads_batch(via_dev)
    assert np.array_equal(hb, db) and np.array_equal(hi, di) and hc == dc and np.array_equal(hp, dp)
    assert (hb >= 0).sum() == int(keep.sum()) and (hb < 0).sum() == int((keep == 0).sum())
