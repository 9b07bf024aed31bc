"""Developer check: kernel-only GCUPS of the Smith-Waterman path on the cfg#4 generator."""
import hashlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lorikeet_b200 as lk
from lorikeet_b200 import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ref_len = int(sys.argv[2]) if len(sys.argv) > 2 else 500  # + 2 x 10 N pads
ro, refs, ao, alts = synth.sw_pairs(n, ref_len=ref_len)
with lk.Context(0) as ctx:
    for strategy in ("SoftClip", "InDel"):
        res = ctx.sw_upload(ro, refs, ao, alts, [200, -150, -260, -11], strategy, max_ops=64)
        for _ in range(2):
            res.run()
        ms = min(res.run() for _ in range(5))
        ops, n_ops, offs = res.download()
        valid = np.arange(64)[None, :] < n_ops[:, None]
        sha = hashlib.sha256(ops[valid].tobytes() + n_ops.tobytes() + offs.tobytes()).hexdigest()[:12]
        print("SW %s pairs=%d cells=%.3e best_ms=%.3f GCUPS=%.1f packed_pairs=%d sha=%s" % (strategy, n, res.n_cells, ms, res.n_cells / ms / 1e6, res.packed_pairs, sha), flush=True)
        outs = (np.zeros((n, 64), np.uint32), np.zeros(n, np.int32), np.zeros(n, np.int32))
        for _ in range(3):
            t = time.time()
            ctx.sw_batch(ro, refs, ao, alts, [200, -150, -260, -11], strategy, max_ops=64, out=outs)
            print("   e2e host call (pageable buffers) %.1f ms" % ((time.time() - t) * 1e3))
        res.free()
