"""Driver for tests/emul (lock-step warp emulator running the kernels' own warp routines on CPU).
TEST TOOLING ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

from lorikeet_b200 import _lib
from lorikeet_b200.batch import PairHmmBatch

HERE = os.path.dirname(os.path.abspath(__file__))
_E = None


def lib():
    global _E
    if _E is None:
        d = os.path.join(HERE, "emul")
        subprocess.run(["make", "-C", d], check=True, capture_output=True)
        _E = C.CDLL(os.path.join(d, "liblk_emul.so"))
        _E.emul_pairhmm_batch.restype = C.c_int
        _E.emul_pairhmm_batch.argtypes = [C.POINTER(_lib.PairHmmBatchStruct), C.c_int, C.c_int, C.c_int, _lib.f64p,
                                          _lib.u8p, _lib.i32p]
    return _E


def pairhmm_batch(batch, use_tristate=True, force_k=0, force_f64=False):
    if isinstance(batch, dict):
        batch = PairHmmBatch.from_dict(batch)
    out = np.full(batch.n_pairs, np.nan)
    flags = np.full(batch.n_pairs, 255, dtype=np.uint8)
    n_resc = C.c_int32(0)
    s = batch.as_struct()
    rc = lib().emul_pairhmm_batch(C.byref(s), 1 if use_tristate else 0, force_k, 1 if force_f64 else 0,
                                  out.ctypes.data_as(_lib.f64p), flags.ctypes.data_as(_lib.u8p), C.byref(n_resc))
    assert rc == 0
    return out, flags, n_resc.value
