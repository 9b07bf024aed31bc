"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module; nothing under lorikeet_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SW_STRATEGY = {"SoftClip": 0, "InDel": 1, "LeadingInDel": 2, "Ignore": 3}
CIGAR_CHARS = "MIDNSHP=X"

u8p = C.POINTER(C.c_uint8)
i64p = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
f64p = C.POINTER(C.c_double)


def build(native=False, quiet=True):
    """Compile the oracle.  native=True builds a -march=native copy for CPU-baseline
    timing on the host it runs on (never shipped)."""
    if native:
        out = os.path.join(HERE, "_build")
        os.makedirs(out, exist_ok=True)
        so = os.path.join(out, "liblk_oracle_native.so")
        cmd = ["gcc", "-O3", "-march=native", "-ffp-contract=off", "-fPIC", "-std=c11", "-D_GNU_SOURCE", "-shared",
               "-o", so, os.path.join(HERE, "lk_oracle.c"), os.path.join(HERE, "lk_oracle_mt.c"),
               os.path.join(HERE, "lk_oracle_simd.c"), "-lm", "-lpthread"]
        subprocess.run(cmd, check=True, capture_output=quiet)
        return so
    subprocess.run(["make", "-C", HERE], check=True, capture_output=quiet)
    return os.path.join(HERE, "liblk_oracle.so")


def _bind(lib):
    lib.lko_qual_to_error_prob.restype = C.c_double
    lib.lko_qual_to_error_prob.argtypes = [C.c_uint8]
    lib.lko_qual_to_prob.restype = C.c_double
    lib.lko_qual_to_prob.argtypes = [C.c_uint8]
    lib.lko_approximate_log10_sum_log10.restype = C.c_double
    lib.lko_approximate_log10_sum_log10.argtypes = [C.c_double, C.c_double]
    lib.lko_match_to_match_prob.restype = C.c_double
    lib.lko_match_to_match_prob.argtypes = [C.c_int, C.c_int]
    lib.lko_qual_to_trans_probs.restype = None
    lib.lko_qual_to_trans_probs.argtypes = [f64p, C.c_uint8, C.c_uint8, C.c_uint8]
    for name in ("lko_pairhmm_forward", "lko_pairhmm_forward_rescaled"):
        f = getattr(lib, name)
        f.restype = C.c_double
        f.argtypes = [u8p, C.c_int, u8p, u8p, u8p, u8p, u8p, C.c_int, C.c_int]
    lib.lko_find_first_position_where_haplotypes_differ.restype = C.c_int
    lib.lko_find_first_position_where_haplotypes_differ.argtypes = [u8p, C.c_int, u8p, C.c_int]
    for name in ("lko_sw_align", "lko_sw_align_full"):
        f = getattr(lib, name)
        f.restype = C.c_int
        f.argtypes = [u8p, C.c_int, u8p, C.c_int, i32p, C.c_int, u32p, C.c_int, i32p]
    lib.lko_last_index_of.restype = C.c_int
    lib.lko_last_index_of.argtypes = [u8p, C.c_int, u8p, C.c_int]
    lib.lko_pcr_error_model_adjusted_qual.restype = C.c_uint8
    lib.lko_pcr_error_model_adjusted_qual.argtypes = [C.c_int, C.c_double]
    lib.lko_find_tandem_repeat_length.restype = C.c_int
    lib.lko_find_tandem_repeat_length.argtypes = [u8p, C.c_int, C.c_int]
    lib.lko_apply_pcr_error_model.restype = None
    lib.lko_apply_pcr_error_model.argtypes = [u8p, C.c_int, u8p, u8p, C.c_int]
    lib.lko_cap_minimum_read_qualities.restype = None
    lib.lko_cap_minimum_read_qualities.argtypes = [u8p, u8p, u8p, C.c_int, C.c_uint8, C.c_uint8, C.c_int]
    lib.lko_normalize_likelihoods.restype = None
    lib.lko_normalize_likelihoods.argtypes = [f64p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int]
    lib.lko_log10_dynamic_read_qual_threshold.restype = C.c_double
    lib.lko_log10_dynamic_read_qual_threshold.argtypes = [u8p, C.c_int, C.c_double]
    lib.lko_log10_min_true_likelihood.restype = C.c_double
    lib.lko_log10_min_true_likelihood.argtypes = [C.c_int, C.c_double, C.c_int]
    lib.lko_filter_poorly_modeled_evidence.restype = C.c_int
    lib.lko_filter_poorly_modeled_evidence.argtypes = [f64p, C.c_int, C.c_int, f64p, u8p]
    lib.lko_init.restype = None
    lib.lko_pairhmm_batch_mt.restype = None
    lib.lko_pairhmm_batch_mt.argtypes = [C.c_int, i64p, i64p, i64p, u8p, u8p, u8p, u8p, u8p, i64p, u8p, i64p,
                                         C.c_int, C.c_int, C.c_int, f64p]
    lib.lko_sw_batch_mt.restype = None
    lib.lko_sw_batch_mt.argtypes = [C.c_int, i64p, u8p, i64p, u8p, i32p, C.c_int, C.c_int, C.c_int, C.c_int,
                                    u32p, i32p, i32p]
    return lib


def lib(path=None):
    global _LIB
    if path is not None:
        return _bind(C.CDLL(path))
    if _LIB is None:
        so = os.path.join(HERE, "liblk_oracle.so")
        if not os.path.exists(so):
            build()
        _LIB = _bind(C.CDLL(so))
        _LIB.lko_init()
    return _LIB


def _u8(a):
    if isinstance(a, (bytes, bytearray)):
        a = np.frombuffer(bytes(a), dtype=np.uint8)
    elif isinstance(a, str):
        a = np.frombuffer(a.encode(), dtype=np.uint8)
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.size == 0:
        a = np.zeros(1, dtype=np.uint8)[:0]
    return a


def _p(a, t):
    return a.ctypes.data_as(t)


def pairhmm_forward(hap, read, qual, ins, dele, gcp, use_tristate=True, rescaled=False):
    hap, read, qual, ins, dele, gcp = map(_u8, (hap, read, qual, ins, dele, gcp))
    fn = lib().lko_pairhmm_forward_rescaled if rescaled else lib().lko_pairhmm_forward
    return fn(_p(hap, u8p), hap.size, _p(read, u8p), _p(qual, u8p), _p(ins, u8p), _p(dele, u8p), _p(gcp, u8p),
              read.size, 1 if use_tristate else 0)


def cigar_to_string(ops):
    return "".join("%d%s" % (int(o) >> 4, CIGAR_CHARS[int(o) & 0xF]) for o in ops)


def sw_align(ref, alt, params, strategy, use_shortcut=True, max_ops=None):
    """Returns (offset, cigar_string).  strategy is a name from SW_STRATEGY or its int."""
    ref, alt = _u8(ref), _u8(alt)
    if isinstance(strategy, str):
        strategy = SW_STRATEGY[strategy]
    if max_ops is None:
        max_ops = 2 * (ref.size + alt.size) + 8
    ops = np.zeros(max_ops, dtype=np.uint32)
    off = C.c_int32(0)
    prm = np.asarray(params, dtype=np.int32)
    fn = lib().lko_sw_align if use_shortcut else lib().lko_sw_align_full
    n = fn(_p(ref, u8p), ref.size, _p(alt, u8p), alt.size, _p(prm, i32p), strategy, _p(ops, u32p), max_ops, C.byref(off))
    if n < 0:
        raise ValueError("oracle sw_align failed with code %d" % n)
    return off.value, cigar_to_string(ops[:n])


def pairhmm_batch(batch, use_tristate=True, rescaled=False, n_threads=None, lib_path=None):
    """batch: object/dict with the packed arrays of lorikeet_b200.batch.PairHmmBatch.  Returns f64 array."""
    L = lib(lib_path)
    if lib_path is not None:
        L.lko_init()
    g = (lambda k: batch[k]) if hasattr(batch, "__getitem__") else (lambda k: getattr(batch, k))
    rrs = np.ascontiguousarray(g("region_read_start"), dtype=np.int64)
    rhs = np.ascontiguousarray(g("region_hap_start"), dtype=np.int64)
    roff = np.ascontiguousarray(g("read_off"), dtype=np.int64)
    hoff = np.ascontiguousarray(g("hap_off"), dtype=np.int64)
    nr = np.diff(rrs)
    nh = np.diff(rhs)
    out_off = np.zeros(len(nr) + 1, dtype=np.int64)
    np.cumsum(nr * nh, out=out_off[1:])
    out = np.zeros(int(out_off[-1]), dtype=np.float64)
    if n_threads is None:
        n_threads = os.cpu_count() or 1
    L.lko_pairhmm_batch_mt(len(nr), _p(rrs, i64p), _p(rhs, i64p), _p(roff, i64p),
                           _p(_u8(g("read_bases")), u8p), _p(_u8(g("read_qual")), u8p), _p(_u8(g("read_ins")), u8p),
                           _p(_u8(g("read_del")), u8p), _p(_u8(g("read_gcp")), u8p), _p(hoff, i64p),
                           _p(_u8(g("hap_bases")), u8p), _p(out_off, i64p), 1 if use_tristate else 0,
                           1 if rescaled else 0, int(n_threads), _p(out, f64p))
    return out


def sw_batch(ref_off, refs, alt_off, alts, params, strategy, use_shortcut=True, max_ops=64, n_threads=None, lib_path=None):
    L = lib(lib_path)
    ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
    alt_off = np.ascontiguousarray(alt_off, dtype=np.int64)
    n = len(ref_off) - 1
    if isinstance(strategy, str):
        strategy = SW_STRATEGY[strategy]
    prm = np.asarray(params, dtype=np.int32)
    ops = np.zeros((n, max_ops), dtype=np.uint32)
    n_ops = np.zeros(n, dtype=np.int32)
    offs = np.zeros(n, dtype=np.int32)
    if n_threads is None:
        n_threads = os.cpu_count() or 1
    L.lko_sw_batch_mt(n, _p(ref_off, i64p), _p(_u8(refs), u8p), _p(alt_off, i64p), _p(_u8(alts), u8p), _p(prm, i32p),
                      strategy, 1 if use_shortcut else 0, max_ops, int(n_threads), _p(ops, u32p), _p(n_ops, i32p),
                      _p(offs, i32p))
    return ops, n_ops, offs


def vector_isa(lib_path=None):
    """ISA the SIMD stand-in picked on this CPU ("avx512 x16 fp32", "avx2+fma x8 fp32" or "sse2 x4 fp32")."""
    L = lib(lib_path)
    L.lko_simd_isa.restype = C.c_char_p
    return L.lko_simd_isa().decode()


def pairhmm_batch_vector(batch, use_tristate=True, n_threads=None, lib_path=None, force_lanes=0, want_redone=False):
    """OUR fp32 SIMD stand-in for the reference's un-vendored gkl AVX path (lk_oracle_simd.c), same contract as
    pairhmm_batch.  force_lanes in (16, 8, 4) pins the variant (tests)."""
    L = lib(lib_path)
    g = lambda k: batch[k] if isinstance(batch, dict) else getattr(batch, k)
    rrs, rhs = np.ascontiguousarray(g("region_read_start"), np.int64), np.ascontiguousarray(g("region_hap_start"), np.int64)
    roff, hoff = np.ascontiguousarray(g("read_off"), np.int64), np.ascontiguousarray(g("hap_off"), np.int64)
    rb, rq, ri, rd, hb = (_u8(g(k)) for k in ("read_bases", "read_qual", "read_ins", "read_del", "hap_bases"))
    gcp = batch["read_gcp"] if isinstance(batch, dict) else batch["read_gcp"]
    gcp = _u8(gcp)
    n_regions = len(rrs) - 1
    pc = np.diff(rrs) * np.diff(rhs)
    out_off = np.zeros(n_regions + 1, dtype=np.int64)
    np.cumsum(pc, out=out_off[1:])
    out = np.empty(int(out_off[-1]), dtype=np.float64)
    if n_threads is None:
        n_threads = os.cpu_count() or 1
    L.lko_simd_force_lanes.restype = None
    L.lko_simd_force_lanes(int(force_lanes))
    L.lko_pairhmm_batch_vector_mt.restype = C.c_int
    i64p = C.POINTER(C.c_int64)
    redone = L.lko_pairhmm_batch_vector_mt(n_regions, _p(rrs, i64p), _p(rhs, i64p), _p(roff, i64p), _p(rb, u8p), _p(rq, u8p), _p(ri, u8p),
                                           _p(rd, u8p), _p(gcp, u8p), _p(hoff, i64p), _p(hb, u8p), _p(out_off, i64p),
                                           1 if use_tristate else 0, int(n_threads), _p(out, f64p))
    L.lko_simd_force_lanes(0)
    return (out, redone) if want_redone else out
