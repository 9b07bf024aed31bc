"""Developer check: cfg#3-style long reads (kernel-only)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lorikeet_b200 as lk
from lorikeet_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rl = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
hl = int(sys.argv[3]) if len(sys.argv) > 3 else 12000
b = synth.pairhmm_regions(n, **dict(synth.CFG3, read_len=rl, hap_len=hl))
with lk.Context(0) as ctx:
    res = ctx.pairhmm_upload(b)
    res.run()
    ms = min(res.run() for _ in range(2))
    out, flags = res.download(want_flags=True)
    print("long reads: regions=%d R=%d H=%d cells=%.3e ms=%.1f GCUPS=%.1f rescued=%d neg_inf=%d finite=%s" % (
        n, rl, hl, res.n_cells, ms, res.n_cells / ms / 1e6, int((flags & 1).sum()), int((flags & 2).sum() // 2),
        bool(np.all(np.isfinite(out)))), flush=True)
    if len(sys.argv) > 4:
        from oracle import lk_oracle as O
        t = time.time()
        exp = O.pairhmm_batch(b.region_slice(0, 1), rescaled=True)
        print("oracle 1 region %.1f s; max rel err %.2e" % (time.time() - t, np.max(np.abs(out[:len(exp)] - exp) / np.abs(exp))))
