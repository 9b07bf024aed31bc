"""GPU tests of lk_pool_* (several devices behind one handle in ONE process): the pooled batch calls and the
pooled batcher must return, bit for bit, what the single-device calls return -- a region's result does not depend on which
device or which chunk computed it.  With one visible GPU the pool has one device; with more, all of them are used."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine(ctx):
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine, PCRErrorModel
    return PairHMMLikelihoodCalculationEngine(ctx, 10, -4.5, PCRErrorModel.CONSERVATIVE, 18, True, 1.0, 0.02, True, False, False)


def _same(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


def test_pool_batch_calls_match_single_device():
    import lorikeet_b200 as lk
    from lorikeet_b200 import synth
    from lorikeet_b200.engine import DevicePool
    n = 700  # > 1e9 cells per chunk floor x several chunks: every device gets work
    raw = synth.pairhmm_regions(n, first_region=77000, pcr_model=0, base_quality_score_threshold=0)
    pre = synth.pairhmm_regions(n, first_region=77000)
    rng = np.random.default_rng(8)
    mapq = rng.integers(20, 61, raw.n_reads).astype(np.uint8)
    ref_idx = rng.integers(0, 3, n).astype(np.int32)
    with lk.Context(0) as ctx:
        eng = _engine(ctx)
        exp_v, exp_keep = eng.compute_read_likelihoods(raw, mapq, ref_idx)
        exp_lk, exp_flags = ctx.pairhmm_batch(pre, want_flags=True)
    with DevicePool() as pool:
        assert pool.device_count >= 1
        v, keep = pool.compute_read_likelihoods(_engine(None), raw, mapq, ref_idx)
        lk2, flags2 = pool.pairhmm_batch(pre, want_flags=True)
    assert _same(v, exp_v) and np.array_equal(keep, exp_keep)
    assert np.array_equal(lk2, exp_lk) and np.array_equal(flags2, exp_flags)


def test_pool_rejects_bad_devices_and_reports_errors():
    import ctypes as C
    from lorikeet_b200 import _lib
    L = _lib.load()
    h = C.c_void_p()
    arr = (C.c_int * 2)(0, 0)  # the same device twice
    assert L.lk_pool_create(arr, 2, C.byref(h)) == _lib.LK_ERR_INVALID
    arr = (C.c_int * 1)(99)
    assert L.lk_pool_create(arr, 1, C.byref(h)) == _lib.LK_ERR_INVALID
    assert L.lk_pool_device_count(None) == 0


def test_pool_batcher_from_many_threads():
    import lorikeet_b200 as lk
    from lorikeet_b200 import synth
    from lorikeet_b200.engine import DevicePool, RegionBatcher
    n = 256
    b = synth.pairhmm_regions(n, first_region=91000, pcr_model=0, base_quality_score_threshold=0)
    mapq = np.full(b.n_reads, 60, np.uint8)
    with lk.Context(0) as ctx:
        exp_v, exp_keep = _engine(ctx).compute_read_likelihoods(b, mapq)
    got = [None] * n
    with DevicePool() as pool:
        with RegionBatcher(pool, _engine(None), max_regions=64, max_wait_us=2000) as bt:
            def work(t, n_threads=32):
                for g in range(t, n, n_threads):
                    r0, r1 = int(b.region_read_start[g]), int(b.region_read_start[g + 1])
                    got[g] = bt.submit(b.region_slice(g, g + 1), mapq[r0:r1], 0)
            th = [threading.Thread(target=work, args=(t,)) for t in range(32)]
            [t.start() for t in th]
            [t.join() for t in th]
            n_batches, n_regions = bt.stats()
        assert n_regions == n and n_batches < n / 2
    pbase = 0
    for g in range(n):
        r0, r1 = int(b.region_read_start[g]), int(b.region_read_start[g + 1])
        npairs = (r1 - r0) * int(b.region_hap_start[g + 1] - b.region_hap_start[g])
        v, keep = got[g]
        assert _same(v, exp_v[pbase:pbase + npairs]) and np.array_equal(keep, exp_keep[r0:r1])
        pbase += npairs


def test_batcher_keeps_each_callers_gcp_and_rejects_bad_requests_at_submit():
    """Two co-batched callers with different gap-continuation constants each get the likelihoods of THEIR constant, and a
    request without the MAPQ array an engine batcher needs is refused at submit instead of failing its batch-mates."""
    import ctypes as C
    import lorikeet_b200 as lk
    from lorikeet_b200 import synth, _lib
    from lorikeet_b200.engine import RegionBatcher
    b = synth.pairhmm_regions(8, first_region=1234)
    with lk.Context(0) as ctx:
        exp = {}
        for gcp in (10, 7):
            bb = lk.PairHmmBatch(b.region_read_start, b.region_hap_start, b.read_off, b.read_bases, b.read_qual, b.read_ins,
                                 b.read_del, b.hap_off, b.hap_bases, read_gcp=None, gcp_constant=gcp)
            exp[gcp] = ctx.pairhmm_batch(bb)
        assert not np.array_equal(exp[10], exp[7])
        got = {}
        with RegionBatcher(ctx, None, max_regions=64, max_wait_us=20000) as bt:
            def work(g):
                gcp = 10 if g % 2 == 0 else 7
                reg = b.region_slice(g, g + 1)
                reg.gcp_constant = gcp
                got[g] = (gcp, bt.submit(reg)[0])
            th = [threading.Thread(target=work, args=(g,)) for g in range(8)]
            [t.start() for t in th]
            [t.join() for t in th]
            n_batches, _ = bt.stats()
        assert n_batches < 8  # regions with different constants did share launches
        for g in range(8):
            gcp, v = got[g]
            assert np.max(np.abs(v - exp[gcp][g * 1024:(g + 1) * 1024]) / np.abs(v)) < 1e-6
        eng = _engine(ctx)
        with RegionBatcher(ctx, eng, max_regions=8, max_wait_us=100) as bt:
            reg = b.region_slice(0, 1)
            s = reg.as_struct()
            vals = np.empty(reg.n_pairs)
            st = bt._L.lk_batcher_submit(bt._h, C.byref(s), None, 0, vals.ctypes.data_as(_lib.f64p), None)
            assert st == _lib.LK_ERR_INVALID
