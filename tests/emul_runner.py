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


def count_fast_tasks(batch, split_target_tasks=0):
    """Warp tasks host_prep builds for the batch (to see the haplotype-axis split of small launches)."""
    if isinstance(batch, dict):
        batch = PairHmmBatch.from_dict(batch)
    L = lib()
    L.emul_count_fast_tasks.restype = C.c_int64
    L.emul_count_fast_tasks.argtypes = [C.POINTER(_lib.PairHmmBatchStruct), C.c_int64]
    s = batch.as_struct()
    return int(L.emul_count_fast_tasks(C.byref(s), int(split_target_tasks)))


def sw_batch(ref_off, refs, alt_off, alts, params, strategy, max_ops=64, force_k=0):
    from oracle.lk_oracle import SW_STRATEGY
    L = lib()
    L.emul_sw_batch.restype = C.c_int
    L.emul_sw_batch.argtypes = [C.c_int64, _lib.i64p, _lib.u8p, _lib.i64p, _lib.u8p, _lib.i32p, C.c_int, C.c_int, C.c_int,
                                _lib.u32p, _lib.i32p, _lib.i32p]
    ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
    alt_off = np.ascontiguousarray(alt_off, dtype=np.int64)
    refs = np.ascontiguousarray(refs, dtype=np.uint8)
    alts = np.ascontiguousarray(alts, dtype=np.uint8)
    n = len(ref_off) - 1
    if isinstance(strategy, str):
        strategy = SW_STRATEGY[strategy]
    prm = np.asarray(params, dtype=np.int32)
    ops = np.zeros((n, max_ops), dtype=np.uint32)
    n_ops = np.zeros(n, dtype=np.int32)
    offs = np.zeros(n, dtype=np.int32)
    p = lambda a, t: a.ctypes.data_as(t)
    rc = L.emul_sw_batch(n, p(ref_off, _lib.i64p), p(refs, _lib.u8p), p(alt_off, _lib.i64p), p(alts, _lib.u8p),
                         p(prm, _lib.i32p), strategy, max_ops, force_k, p(ops, _lib.u32p), p(n_ops, _lib.i32p),
                         p(offs, _lib.i32p))
    assert rc == 0
    return ops, n_ops, offs


def sw_duo_batch(ref_off, refs, alt_off, alts, params, strategy, max_ops=64, force_k=0):
    """Consecutive pairs (2i, 2i+1) through the packed two-pairs-per-warp routine with sw_duo_kernel's dispatch; also returns
    which pairs actually took the packed path."""
    from oracle.lk_oracle import SW_STRATEGY
    L = lib()
    L.emul_sw_duo_batch.restype = C.c_int
    L.emul_sw_duo_batch.argtypes = [C.c_int64, _lib.i64p, _lib.u8p, _lib.i64p, _lib.u8p, _lib.i32p, C.c_int, C.c_int, C.c_int,
                                    _lib.u32p, _lib.i32p, _lib.i32p, _lib.u8p]
    ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
    alt_off = np.ascontiguousarray(alt_off, dtype=np.int64)
    refs = np.ascontiguousarray(refs, dtype=np.uint8)
    alts = np.ascontiguousarray(alts, dtype=np.uint8)
    n = len(ref_off) - 1
    if isinstance(strategy, str):
        strategy = SW_STRATEGY[strategy]
    prm = np.asarray(params, dtype=np.int32)
    ops = np.zeros((n, max_ops), dtype=np.uint32)
    n_ops = np.zeros(n, dtype=np.int32)
    offs = np.zeros(n, dtype=np.int32)
    took = np.zeros(n, dtype=np.uint8)
    p = lambda a, t: a.ctypes.data_as(t)
    rc = L.emul_sw_duo_batch(n, p(ref_off, _lib.i64p), p(refs, _lib.u8p), p(alt_off, _lib.i64p), p(alts, _lib.u8p),
                             p(prm, _lib.i32p), strategy, max_ops, force_k, p(ops, _lib.u32p), p(n_ops, _lib.i32p),
                             p(offs, _lib.i32p), p(took, _lib.u8p))
    assert rc == 0, rc
    return ops, n_ops, offs, took
