"""Developer check: lk_calculate_cigar_batch end to end (host shortcuts + padding, GPU SW batch, host CIGAR surgery) on
cfg#4-style pairs given un-padded, next to the bare lk_sw_batch_host call on the padded pairs."""
import ctypes as C
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import lorikeet_b200 as lk
from lorikeet_b200 import synth, _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
ro, refs, ao, alts = synth.sw_pairs(n, pad=0)
L = _lib.load()
prm = _lib.SwParams(200, -150, -260, -11)
ops = np.zeros((n, 64), np.uint32)
n_ops = np.zeros(n, np.int32)
p = lambda a, t: a.ctypes.data_as(t)
with lk.Context(0) as ctx:
    for _ in range(4):
        t = time.perf_counter()
        st = L.lk_calculate_cigar_batch(ctx._h, n, p(ro, _lib.i64p), p(refs, _lib.u8p), p(ao, _lib.i64p), p(alts, _lib.u8p),
                                        C.byref(prm), 1, 64, p(ops, _lib.u32p), p(n_ops, _lib.i32p))
        dt = time.perf_counter() - t
        assert st == 0
        print("lk_calculate_cigar_batch pairs=%d  %.1f ms  (%.2f us per pair)  failed=%d" % (n, dt * 1e3, dt * 1e6 / n, int((n_ops < 0).sum())), flush=True)
