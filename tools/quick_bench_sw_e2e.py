"""Developer check: lk_sw_batch_host end to end on the cfg#4 generator from page-locked host buffers (what bench.py reports as
other_configs.cfg4_smith_waterman.e2e), several repetitions."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import lorikeet_b200 as lk
from lorikeet_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ro, refs, ao, alts = synth.sw_pairs(n)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
pro, prefs, pao, palts = pin(ro), pin(refs), pin(ao), pin(alts)
outs = (torch.empty((n, 64), dtype=torch.int32).pin_memory().numpy().view(np.uint32),
        torch.empty(n, dtype=torch.int32).pin_memory().numpy(), torch.empty(n, dtype=torch.int32).pin_memory().numpy())
cells = float((np.diff(ro) * np.diff(ao)).sum())
with lk.Context(0) as ctx:
    for i in range(8):
        t0 = time.perf_counter()
        ctx.sw_batch(pro, prefs, pao, palts, [200, -150, -260, -11], "SoftClip", max_ops=64, out=outs)
        dt = time.perf_counter() - t0
        print("e2e %.2f ms  %.0f GCUPS" % (dt * 1e3, cells / dt / 1e9), flush=True)
