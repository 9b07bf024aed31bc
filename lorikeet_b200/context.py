"""Context = one lk_ctx (one CUDA device).  Thin, typed wrappers over the C ABI."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import LorikeetError

SW_STRATEGY = {"SoftClip": 0, "InDel": 1, "LeadingInDel": 2, "Ignore": 3}
CIGAR_CHARS = "MIDNSHP=X"


def cigar_to_string(ops):
    return "".join("%d%s" % (int(o) >> 4, CIGAR_CHARS[int(o) & 0xF]) for o in ops)


def cigar_from_string(text):
    """'3M1I2M' -> BAM-encoded uint32 elements (len << 4 | op)."""
    import re
    return np.asarray([(int(n) << 4) | CIGAR_CHARS.index(c) for n, c in re.findall(r"(\d+)([MIDNSHP=X])", text)], dtype=np.uint32)


class RealignBatch:
    """Inputs of AssemblyBasedCallerUtils::realign_reads_to_their_best_haplotype for a batch of single-sample regions
    (assembly/assembly_based_caller_utils.rs:208-246); see lk_realign_batch in include/lorikeet_b200.h."""

    def __init__(self, region_read_start, region_hap_start, values, hap_seqs, hap_cigars, hap_start_wrt_ref, read_seqs,
                 read_cigars, reference_start, read_orig_len=None, ref_allele_index=None, region_n_evidence=None,
                 informative_threshold=0.0, region_sample_start=None, sample_read_start=None, sample_n_evidence=None,
                 values_device_ptr=None):
        as_u8 = lambda x: np.frombuffer(x.encode(), np.uint8) if isinstance(x, str) else np.asarray(x, np.uint8)
        self.region_read_start = np.ascontiguousarray(region_read_start, np.int64)
        self.region_hap_start = np.ascontiguousarray(region_hap_start, np.int64)
        # values: host matrices, or (values_device_ptr) the address of the same blocks in the context's device memory
        self.values = None if values is None else np.ascontiguousarray(values, np.float64).ravel()
        self.values_device_ptr = values_device_ptr
        i64 = lambda x: None if x is None else np.ascontiguousarray(x, np.int64)
        self.region_sample_start, self.sample_read_start, self.sample_n_evidence = i64(region_sample_start), i64(sample_read_start), i64(sample_n_evidence)
        haps = [as_u8(h) for h in hap_seqs]
        reads = [as_u8(r) for r in read_seqs]
        hc = [cigar_from_string(c) if isinstance(c, str) else np.asarray(c, np.uint32) for c in hap_cigars]
        rc = [cigar_from_string(c) if isinstance(c, str) else np.asarray(c, np.uint32) for c in read_cigars]
        cat = lambda xs, dt: np.ascontiguousarray(np.concatenate(xs) if len(xs) else np.zeros(0, dt), dt)
        off = lambda xs: np.ascontiguousarray(np.concatenate([[0], np.cumsum([len(x) for x in xs])]), np.int64)
        self.hap_off, self.hap_bases = off(haps), cat(haps, np.uint8)
        self.read_off, self.read_bases = off(reads), cat(reads, np.uint8)
        self.hap_cigar_off, self.hap_cigars = off(hc), cat(hc, np.uint32)
        self.read_cigar_off, self.read_cigars = off(rc), cat(rc, np.uint32)
        self.hap_start_wrt_ref = np.ascontiguousarray(hap_start_wrt_ref, np.int64)
        self.reference_start = np.ascontiguousarray(reference_start, np.int64)
        self.read_orig_len = np.ascontiguousarray(read_orig_len if read_orig_len is not None else [len(r) for r in reads], np.int64)
        self.ref_allele_index = None if ref_allele_index is None else np.ascontiguousarray(ref_allele_index, np.int32)
        self.region_n_evidence = None if region_n_evidence is None else np.ascontiguousarray(region_n_evidence, np.int64)
        self.informative_threshold = float(informative_threshold)

    @property
    def n_reads(self):
        return int(self.region_read_start[-1])

    def as_struct(self):
        p = lambda a, t: None if a is None else a.ctypes.data_as(t)
        S = _lib.RealignBatchStruct
        vals = p(self.values, _lib.f64p) if self.values_device_ptr is None else C.cast(C.c_void_p(int(self.values_device_ptr)), _lib.f64p)
        return S(len(self.region_read_start) - 1, p(self.region_read_start, _lib.i64p), p(self.region_hap_start, _lib.i64p),
                 vals, p(self.region_n_evidence, _lib.i64p), p(self.ref_allele_index, _lib.i32p),
                 p(self.reference_start, _lib.i64p), p(self.hap_off, _lib.i64p), p(self.hap_bases, _lib.u8p),
                 p(self.hap_cigar_off, _lib.i64p), p(self.hap_cigars, _lib.u32p), p(self.hap_start_wrt_ref, _lib.i64p),
                 p(self.read_off, _lib.i64p), p(self.read_bases, _lib.u8p), p(self.read_cigar_off, _lib.i64p),
                 p(self.read_cigars, _lib.u32p), p(self.read_orig_len, _lib.i64p), self.informative_threshold,
                 p(self.region_sample_start, _lib.i64p), p(self.sample_read_start, _lib.i64p), p(self.sample_n_evidence, _lib.i64p))


class Context:
    def __init__(self, device=-1):
        self._L = _lib.load()
        self._h = C.c_void_p()
        st = self._L.lk_create(int(device), C.byref(self._h))
        if st != 0:
            self._h = None
            raise LorikeetError(st, "lk_create failed (a CUDA device is required; there is no CPU fallback)")

    def close(self):
        if getattr(self, "_h", None):
            self._L.lk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, st):
        if st != 0:
            raise LorikeetError(st, (self._L.lk_last_error(self._h) or b"").decode())

    def version(self):
        return self._L.lk_version().decode()

    def set_tristate_correction(self, on):
        self._check(self._L.lk_set_tristate_correction(self._h, 1 if on else 0))

    def set_force_f64(self, on):
        self._check(self._L.lk_set_force_f64(self._h, 1 if on else 0))

    def measure_fp32_fma_tflops(self, duration_ms=300.0):
        """FFMA-only microbenchmark: the fp32 roofline denominator measured on this device."""
        v = C.c_double(0.0)
        self._check(self._L.lk_measure_fp32_fma_tflops(self._h, float(duration_ms), C.byref(v)))
        return v.value

    # ---------------------------------------------------------------- PairHMM
    def pairhmm_batch(self, batch, want_flags=False, out=None, flags=None):
        """lk_pairhmm_batch_host: host arrays in, log10 likelihoods out (layout r*A + h per region).
        `out` / `flags` may be preallocated (e.g. pinned) float64 / uint8 arrays of n_pairs entries."""
        n = batch.n_pairs
        if out is None:
            out = np.empty(n, dtype=np.float64)
        if flags is None:
            flags = np.zeros(n, dtype=np.uint8)
        assert out.dtype == np.float64 and out.size >= n and out.flags.c_contiguous
        assert flags.dtype == np.uint8 and flags.size >= n and flags.flags.c_contiguous
        s = batch.as_struct()
        self._check(self._L.lk_pairhmm_batch_host(self._h, C.byref(s), out.ctypes.data_as(_lib.f64p),
                                                  flags.ctypes.data_as(_lib.u8p)))
        return (out, flags) if want_flags else out

    def pairhmm_forward(self, hap, read, qual, ins, dele, gcp):
        """Single pair, gkl::pairhmm::forward()'s argument list (pair_hmm.rs:358-365)."""
        from .batch import _u8
        hap, read, qual, ins, dele, gcp = map(_u8, (hap, read, qual, ins, dele, gcp))
        out = C.c_double(0.0)
        p = lambda a: a.ctypes.data_as(_lib.u8p)
        self._check(self._L.lk_pairhmm_forward(self._h, p(hap), hap.size, p(read), p(qual), p(ins), p(dele), p(gcp),
                                               read.size, C.byref(out)))
        return out.value

    def pairhmm_upload(self, batch):
        return ResidentPairHmm(self, batch)

    # ---------------------------------------------------------------- Smith-Waterman
    def sw_batch(self, ref_off, refs, alt_off, alts, params, strategy, max_ops=64, out=None):
        """lk_sw_batch_host; out = (ops[n, max_ops] u32, n_ops[n] i32, offsets[n] i32) to reuse (e.g. pinned) buffers."""
        from .batch import _u8, _i64
        ref_off, alt_off, refs, alts = _i64(ref_off), _i64(alt_off), _u8(refs), _u8(alts)
        n = len(ref_off) - 1
        if isinstance(strategy, str):
            strategy = SW_STRATEGY[strategy]
        prm = _lib.SwParams(*[int(x) for x in params])
        if out is not None:
            ops, n_ops, offs = out
            assert ops.dtype == np.uint32 and ops.shape == (n, max_ops) and ops.flags.c_contiguous
            assert n_ops.dtype == np.int32 and n_ops.size == n and offs.dtype == np.int32 and offs.size == n
        else:
            ops = np.zeros((n, max_ops), dtype=np.uint32)
            n_ops = np.zeros(n, dtype=np.int32)
            offs = np.zeros(n, dtype=np.int32)
        self._check(self._L.lk_sw_batch_host(self._h, n, ref_off.ctypes.data_as(_lib.i64p), refs.ctypes.data_as(_lib.u8p),
                                             alt_off.ctypes.data_as(_lib.i64p), alts.ctypes.data_as(_lib.u8p),
                                             C.byref(prm), int(strategy), int(max_ops), ops.ctypes.data_as(_lib.u32p),
                                             n_ops.ctypes.data_as(_lib.i32p), offs.ctypes.data_as(_lib.i32p)))
        return ops, n_ops, offs

    def sw_align(self, ref, alt, params, strategy, cigar_cap=4096):
        """Single pair, gkl::smithwaterman::align()'s shape: returns (offset, "31M20S")."""
        from .batch import _u8
        ref, alt = _u8(ref), _u8(alt)
        if isinstance(strategy, str):
            strategy = SW_STRATEGY[strategy]
        prm = _lib.SwParams(*[int(x) for x in params])
        buf = C.create_string_buffer(cigar_cap)
        off = C.c_int32(0)
        self._check(self._L.lk_sw_align(self._h, ref.ctypes.data_as(_lib.u8p), ref.size, alt.ctypes.data_as(_lib.u8p),
                                        alt.size, C.byref(prm), int(strategy), buf, cigar_cap, C.byref(off)))
        return off.value, buf.value.decode()

    def calculate_cigar_batch(self, ref_off, refs, alt_off, alts, params, strategy="InDel", max_ops=64):
        """CigarUtils::calculate_cigar (reads/cigar_utils.rs:358-457) for a batch: list of CIGAR strings (None where the
        reference returns None)."""
        from .batch import _u8, _i64
        ref_off, alt_off, refs, alts = _i64(ref_off), _i64(alt_off), _u8(refs), _u8(alts)
        n = len(ref_off) - 1
        if isinstance(strategy, str):
            strategy = SW_STRATEGY[strategy]
        prm = _lib.SwParams(*[int(x) for x in params])
        ops = np.zeros((n, max_ops), dtype=np.uint32)
        n_ops = np.zeros(n, dtype=np.int32)
        self._check(self._L.lk_calculate_cigar_batch(self._h, n, ref_off.ctypes.data_as(_lib.i64p), refs.ctypes.data_as(_lib.u8p),
                                                     alt_off.ctypes.data_as(_lib.i64p), alts.ctypes.data_as(_lib.u8p), C.byref(prm),
                                                     int(strategy), int(max_ops), ops.ctypes.data_as(_lib.u32p),
                                                     n_ops.ctypes.data_as(_lib.i32p)))
        out = []
        for i in range(n):
            if n_ops[i] == -3:
                out.append(None)
            elif n_ops[i] < 0:
                raise LorikeetError(_lib.LK_ERR_INVALID, "calculate_cigar failed for pair %d (code %d)" % (i, n_ops[i]))
            else:
                out.append(cigar_to_string(ops[i, :n_ops[i]]))
        return out

    def dangling_end_sw_batch(self, ref_off, refs, alt_off, alts, params, max_ops=64):
        """The alignment step of the read-threading graph's dangling-end merges (read_threading_graph.rs:1408,1471):
        LeadingInDel + remove_trailing_deletions; list of CIGAR strings."""
        from .batch import _u8, _i64
        ref_off, alt_off, refs, alts = _i64(ref_off), _i64(alt_off), _u8(refs), _u8(alts)
        n = len(ref_off) - 1
        prm = _lib.SwParams(*[int(x) for x in params])
        ops = np.zeros((n, max_ops), dtype=np.uint32)
        n_ops = np.zeros(n, dtype=np.int32)
        self._check(self._L.lk_dangling_end_sw_batch(self._h, n, ref_off.ctypes.data_as(_lib.i64p), refs.ctypes.data_as(_lib.u8p),
                                                     alt_off.ctypes.data_as(_lib.i64p), alts.ctypes.data_as(_lib.u8p), C.byref(prm),
                                                     int(max_ops), ops.ctypes.data_as(_lib.u32p), n_ops.ctypes.data_as(_lib.i32p)))
        return [cigar_to_string(ops[i, :n_ops[i]]) if n_ops[i] >= 0 else None for i in range(n)]

    def realign_reads_batch(self, batch, max_ops=64):
        """realign_reads_to_their_best_haplotype: (best_allele[int32], informative[uint8], cigars[list of str or None where
        the read is not realigned], pos[int64])."""
        n = batch.n_reads
        best = np.zeros(n, np.int32)
        info = np.zeros(n, np.uint8)
        ops = np.zeros((max(n, 1), max_ops), np.uint32)
        n_ops = np.zeros(n, np.int32)
        pos = np.zeros(n, np.int64)
        st = batch.as_struct()
        self._check(self._L.lk_realign_reads_batch(self._h, C.byref(st), int(max_ops), best.ctypes.data_as(_lib.i32p),
                                                   info.ctypes.data_as(_lib.u8p), ops.ctypes.data_as(_lib.u32p),
                                                   n_ops.ctypes.data_as(_lib.i32p), pos.ctypes.data_as(_lib.i64p)))
        cigars = []
        for i in range(n):
            if n_ops[i] == -1:
                cigars.append(None)
            elif n_ops[i] < 0:
                raise LorikeetError(_lib.LK_ERR_INVALID, "realignment failed for read %d (code %d)" % (i, n_ops[i]))
            else:
                cigars.append(cigar_to_string(ops[i, :n_ops[i]]))
        return best, info, cigars, pos

    def sw_upload(self, ref_off, refs, alt_off, alts, params, strategy, max_ops=64):
        return ResidentSw(self, ref_off, refs, alt_off, alts, params, strategy, max_ops)


class ResidentPairHmm:
    """A batch staged in HBM (lk_pairhmm_upload) for kernel-only timing."""

    def __init__(self, ctx, batch):
        self.ctx, self.batch = ctx, batch
        self._r = C.c_void_p()
        s = batch.as_struct()
        ctx._check(ctx._L.lk_pairhmm_upload(ctx._h, C.byref(s), C.byref(self._r)))
        self.n_pairs = int(ctx._L.lk_resident_pairs(self._r))
        self.n_cells = int(ctx._L.lk_resident_cells(self._r))

    def run(self):
        ms = C.c_float(0.0)
        self.ctx._check(self.ctx._L.lk_pairhmm_run_resident(self.ctx._h, self._r, C.byref(ms)))
        return ms.value

    @property
    def launch_count(self):
        return int(self.ctx._L.lk_resident_launch_count(self._r))

    @property
    def last_fast_ms(self):
        """Device time of the dominant (fp32 wavefront) kernel launches inside the last run()."""
        return float(self.ctx._L.lk_resident_last_fast_ms(self._r))

    @property
    def out_device_ptr(self):
        return int(self.ctx._L.lk_resident_out_device_ptr(self._r))

    def download(self, want_flags=False):
        out = np.empty(self.n_pairs, dtype=np.float64)
        flags = np.zeros(self.n_pairs, dtype=np.uint8)
        self.ctx._check(self.ctx._L.lk_pairhmm_download(self.ctx._h, self._r, out.ctypes.data_as(_lib.f64p),
                                                        flags.ctypes.data_as(_lib.u8p)))
        return (out, flags) if want_flags else out

    def free(self):
        if self._r:
            self.ctx._L.lk_resident_free(self.ctx._h, self._r)
            self._r = None

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass


class ResidentSw:
    def __init__(self, ctx, ref_off, refs, alt_off, alts, params, strategy, max_ops):
        from .batch import _u8, _i64
        self.ctx = ctx
        ref_off, alt_off, refs, alts = _i64(ref_off), _i64(alt_off), _u8(refs), _u8(alts)
        self.n = len(ref_off) - 1
        self.max_ops = int(max_ops)
        if isinstance(strategy, str):
            strategy = SW_STRATEGY[strategy]
        prm = _lib.SwParams(*[int(x) for x in params])
        self._r = C.c_void_p()
        ctx._check(ctx._L.lk_sw_upload(ctx._h, self.n, ref_off.ctypes.data_as(_lib.i64p), refs.ctypes.data_as(_lib.u8p),
                                       alt_off.ctypes.data_as(_lib.i64p), alts.ctypes.data_as(_lib.u8p), C.byref(prm),
                                       int(strategy), self.max_ops, C.byref(self._r)))
        self.n_cells = int((np.diff(ref_off) * np.diff(alt_off)).sum())

    def run(self):
        ms = C.c_float(0.0)
        self.ctx._check(self.ctx._L.lk_sw_run_resident(self.ctx._h, self._r, C.byref(ms)))
        return ms.value

    @property
    def packed_pairs(self):
        """Pairs of the last run() that the packed two-pairs-per-warp kernel took."""
        return int(self.ctx._L.lk_resident_sw_packed_pairs(self._r))

    def download(self):
        ops = np.zeros((self.n, self.max_ops), dtype=np.uint32)
        n_ops = np.zeros(self.n, dtype=np.int32)
        offs = np.zeros(self.n, dtype=np.int32)
        self.ctx._check(self.ctx._L.lk_sw_download(self.ctx._h, self._r, ops.ctypes.data_as(_lib.u32p),
                                                   n_ops.ctypes.data_as(_lib.i32p), offs.ctypes.data_as(_lib.i32p)))
        return ops, n_ops, offs

    def free(self):
        if self._r:
            self.ctx._L.lk_resident_free(self.ctx._h, self._r)
            self._r = None

    def __del__(self):
        try:
            if self.ctx._h:
                self.free()
        except Exception:
            pass
