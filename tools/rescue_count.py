import sys; sys.path.insert(0, "/root/repo")
import numpy as np, lorikeet_b200 as lk
from lorikeet_b200 import synth
b = synth.pairhmm_regions(2000)
out = np.empty(b.n_pairs); fl = np.zeros(b.n_pairs, np.uint8)
with lk.Context(0) as ctx:
    ctx.pairhmm_batch(b, out=out, flags=fl)
print("pairs", b.n_pairs, "rescued", int((fl & 1).sum()), "neg_inf", int((fl & 2).sum()), "min log10", out.min())
