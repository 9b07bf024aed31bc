"""Developer check: end-to-end cost of the region driver (lk_engine_compute_read_likelihoods: host pre-ops + GPU PairHMM +
host post-ops) next to the bare lk_pairhmm_batch_host call on the same batch."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lorikeet_b200 as lk
from lorikeet_b200 import synth
from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
b = synth.pairhmm_regions(n)
mapq = np.full(b.n_reads, 60, np.uint8)
with lk.Context(0) as ctx:
    eng = PairHMMLikelihoodCalculationEngine(ctx)
    for name, f in (("lk_pairhmm_batch_host", lambda: ctx.pairhmm_batch(b)),
                    ("lk_engine_compute_read_likelihoods", lambda: eng.compute_read_likelihoods(b, mapq))):
        f()
        t = []
        for _ in range(3):
            t0 = time.perf_counter(); f(); t.append(time.perf_counter() - t0)
        print("%-36s regions=%d best %.1f ms  %.1f GCUPS" % (name, n, min(t) * 1e3, b.n_cells / min(t) / 1e9), flush=True)
