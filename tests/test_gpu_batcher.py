"""GPU test of the cross-thread batching front-end (lk_batcher): many threads submit one region each, exactly as
Lorikeet's Rayon workers call the engine (assembly_region_walker.rs:140-277); every thread must get back what one big
call returns for its region, and the regions must actually have been merged into few launches."""
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("use_engine", [True, False])
def test_threads_submitting_single_regions_match_one_batch(use_engine):
    import lorikeet_b200 as lk
    from lorikeet_b200 import synth
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine, PCRErrorModel, RegionBatcher
    n = 192
    b = synth.pairhmm_regions(n, first_region=31000, pcr_model=0, base_quality_score_threshold=0)
    rng = np.random.default_rng(2)
    mapq = rng.integers(20, 61, b.n_reads).astype(np.uint8)
    ref_idx = rng.integers(0, 2, n).astype(np.int32)
    with lk.Context(0) as ctx:
        eng = PairHMMLikelihoodCalculationEngine(ctx, 10, -4.5, PCRErrorModel.CONSERVATIVE, 18, True, 1.0, 0.02, True, False, False)
        if use_engine:
            exp_v, exp_aux = eng.compute_read_likelihoods(b, mapq, ref_idx)
        else:
            exp_v, exp_aux = ctx.pairhmm_batch(b, want_flags=True)
        got = [None] * n
        with RegionBatcher(ctx, eng if use_engine else None, max_regions=64, max_wait_us=2000) as bt:
            def work(t, n_threads=16):
                for g in range(t, n, n_threads):
                    r0, r1 = int(b.region_read_start[g]), int(b.region_read_start[g + 1])
                    got[g] = bt.submit(b.region_slice(g, g + 1), mapq[r0:r1], int(ref_idx[g]))
            th = [threading.Thread(target=work, args=(t,)) for t in range(16)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            n_batches, n_regions = bt.stats()
        assert n_regions == n and n_batches < n / 2, (n_batches, n_regions)  # regions were merged
        pbase = 0
        for g in range(n):
            r0, r1 = int(b.region_read_start[g]), int(b.region_read_start[g + 1])
            A = int(b.region_hap_start[g + 1] - b.region_hap_start[g])
            npairs = (r1 - r0) * A
            v, aux = got[g]
            ev = exp_v[pbase:pbase + npairs]
            assert np.array_equal(np.isnan(v), np.isnan(ev))
            assert np.array_equal(v[~np.isnan(v)], ev[~np.isnan(ev)])  # same kernels, same inputs: bit for bit
            assert np.array_equal(aux, exp_aux[r0:r1] if use_engine else exp_aux[pbase:pbase + npairs])
            pbase += npairs


def test_async_submit_keeps_several_regions_in_flight_per_thread():
    # lk_batcher_submit_async / lk_batcher_wait: 4 threads with 24 regions in flight each; results bit for bit those of
    # one big call, and the launches carry far more regions than there are threads
    import lorikeet_b200 as lk
    from lorikeet_b200 import synth
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine, PCRErrorModel, RegionBatcher
    n = 192
    b = synth.pairhmm_regions(n, first_region=52000, pcr_model=0, base_quality_score_threshold=0)
    mapq = np.full(b.n_reads, 60, np.uint8)
    with lk.Context(0) as ctx:
        eng = PairHMMLikelihoodCalculationEngine(ctx, 10, -4.5, PCRErrorModel.CONSERVATIVE, 18, True, 1.0, 0.02, True, False, False)
        exp_v, exp_keep = eng.compute_read_likelihoods(b, mapq)
        got = [None] * n
        with RegionBatcher(ctx, eng, max_regions=256, max_wait_us=2000) as bt:
            def work(t, n_threads=4, depth=24):
                mine = list(range(t, n, n_threads))
                for lo in range(0, len(mine), depth):
                    tickets = []
                    for g in mine[lo:lo + depth]:
                        r0, r1 = int(b.region_read_start[g]), int(b.region_read_start[g + 1])
                        tickets.append((g, bt.submit_async(b.region_slice(g, g + 1), mapq[r0:r1], 0)))
                    for g, tk in tickets:
                        got[g] = bt.wait(tk)
            th = [threading.Thread(target=work, args=(t,)) for t in range(4)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            n_batches, n_regions = bt.stats()
        assert n_regions == n and n_batches <= n // 8, (n_batches, n_regions)
        pbase = 0
        for g in range(n):
            r0, r1 = int(b.region_read_start[g]), int(b.region_read_start[g + 1])
            npairs = (r1 - r0) * int(b.region_hap_start[g + 1] - b.region_hap_start[g])
            v, keep = got[g]
            ev = exp_v[pbase:pbase + npairs]
            assert np.array_equal(np.isnan(v), np.isnan(ev)) and np.array_equal(v[~np.isnan(v)], ev[~np.isnan(ev)])
            assert np.array_equal(keep, exp_keep[r0:r1])
            pbase += npairs
