"""Developer check: a small Smith-Waterman workload for compute-sanitizer (packed two-pairs-per-warp kernel in its three K
classes, the one-pair fall-back inside it, every overhang strategy)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lorikeet_b200 as lk
from lorikeet_b200 import synth

with lk.Context(0) as ctx:
    for ref_len, n in ((500, 120), (400, 60), (200, 60)):
        ro, refs, ao, alts = synth.sw_pairs(n, ref_len=ref_len)
        alts = alts.copy()
        alts[ao[3] + 30] = ord("R")  # one pair falls back to the one-pair routine
        for strategy in ("SoftClip", "InDel", "LeadingInDel", "Ignore"):
            res = ctx.sw_upload(ro, refs, ao, alts, [200, -150, -260, -11], strategy, max_ops=64)
            res.run()
            ops, n_ops, offs = res.download()
            print(ref_len, strategy, "packed", res.packed_pairs, "of", n, "ops", int(n_ops.sum()), flush=True)
            res.free()
