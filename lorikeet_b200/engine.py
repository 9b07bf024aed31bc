"""Host mirror of the reference's PairHMMLikelihoodCalculationEngine / PairHMM / SmithWatermanAligner API surface
(src/pair_hmm/pair_hmm_likelihood_calculation_engine.rs, src/pair_hmm/pair_hmm.rs,
src/smith_waterman/smith_waterman_aligner.rs), so the parity tests read like the reference's own tests.
Everything here forwards to the C ABI; there is no computation in Python."""
import ctypes as C

import numpy as np

from . import _lib
from .context import Context

# smith_waterman_aligner.rs:12-25
ORIGINAL_DEFAULT = (3, -1, -4, -3)
STANDARD_NGS = (25, -50, -110, -6)
NEW_SW_PARAMETERS = (200, -150, -260, -11)
ALIGNMENT_TO_BEST_HAPLOTYPE_SW_PARAMETERS = (10, -15, -30, -5)


class PCRErrorModel:  # pair_hmm_likelihood_calculation_engine.rs:60-70
    NONE, HOSTILE, AGGRESSIVE, CONSERVATIVE = 0, 1, 2, 3


class AVXMode:  # :654-672 — the selection seam; the third arm this library adds
    AVX, NONE, GPU = "AVX", "None", "Gpu"


class PairHMMLikelihoodCalculationEngine:
    """Same constructor arguments, in the same order, as PairHMMLikelihoodCalculationEngine::new (:129-141)."""

    def __init__(self, ctx, constant_gcp=10, log10_global_read_mismapping_rate=-4.5, pcr_error_model=PCRErrorModel.CONSERVATIVE,
                 base_quality_score_threshold=18, dynamic_read_disqualification=False, read_disqualification_scale=1.0,
                 expected_error_rate_per_base=0.02, symmetrically_normalize_alleles_to_reference=True,
                 disable_cap_read_qualities_to_mapq=False, modify_soft_clipped_bases=False, avx_mode=AVXMode.GPU):
        assert ctx is None or isinstance(ctx, (Context, DevicePool))
        self.ctx = ctx
        self.args = _lib.EngineArgs(int(constant_gcp), float(log10_global_read_mismapping_rate), int(pcr_error_model),
                                    int(base_quality_score_threshold), 1 if dynamic_read_disqualification else 0,
                                    float(read_disqualification_scale), float(expected_error_rate_per_base),
                                    1 if symmetrically_normalize_alleles_to_reference else 0,
                                    1 if disable_cap_read_qualities_to_mapq else 0, 1 if modify_soft_clipped_bases else 0)

    def compute_read_likelihoods(self, batch, read_mapq, ref_allele_index=None, values=None, keep=None, values_device_ptr=None):
        """compute_read_likelihoods (:195-242) for a batch of single-sample regions.
        Returns (values, keep): values holds per region an [A x R] row-major block ([allele][read], the reference's
        AlleleLikelihoods layout) at the pair offsets of the batch, normalised and filtered; keep[r] = read survived.
        values_device_ptr: address of n_pairs f64 in the context's device memory; the matrices are then left in HBM
        (for a collective) and `values` is returned as None."""
        L = _lib.load()
        n = batch.n_pairs
        if values_device_ptr is not None:
            values_arg = C.cast(C.c_void_p(int(values_device_ptr)), _lib.f64p)
        else:
            if values is None:
                values = np.empty(n, dtype=np.float64)
            assert values.dtype == np.float64 and values.size == n
            values_arg = values.ctypes.data_as(_lib.f64p)
        if keep is None:
            keep = np.ones(batch.n_reads, dtype=np.uint8)
        assert keep.dtype == np.uint8 and keep.size == batch.n_reads
        mapq = np.ascontiguousarray(read_mapq, dtype=np.uint8)
        assert mapq.size == batch.n_reads
        ref = None
        if ref_allele_index is not None:
            ref = np.ascontiguousarray(ref_allele_index, dtype=np.int32)
        s = batch.as_struct()
        st = L.lk_engine_compute_read_likelihoods(self.ctx._h, C.byref(self.args), C.byref(s), mapq.ctypes.data_as(_lib.u8p),
                                                  ref.ctypes.data_as(_lib.i32p) if ref is not None else C.cast(None, _lib.i32p),
                                                  values_arg, keep.ctypes.data_as(_lib.u8p))
        self.ctx._check(st)
        return values, keep

    def modify_read_qualities_device(self, batch):
        """modify_read_qualities (:352-426) as ONE device launch, in place on the batch's quality arenas
        (lk_engine_modify_read_qualities_device); MAPQ capping must be disabled or irrelevant (MAPQ 60 assumed)."""
        L = _lib.load()
        mapq = np.full(batch.n_reads, 60, dtype=np.uint8)
        p = lambda a, t: a.ctypes.data_as(t)
        for a in (batch.read_qual, batch.read_ins, batch.read_del):
            assert a.flags.writeable and a.flags.c_contiguous
        st = L.lk_engine_modify_read_qualities_device(self.ctx._h, C.byref(self.args), batch.n_reads, p(batch.read_off, _lib.i64p),
                                                      p(batch.read_bases, _lib.u8p), p(batch.read_qual, _lib.u8p),
                                                      p(batch.read_ins, _lib.u8p), p(batch.read_del, _lib.u8p), p(mapq, _lib.u8p))
        self.ctx._check(st)
        return batch


class PairHMM:
    """PairHMM (pair_hmm.rs): initialize / quick_initialize (:63-168), compute_log10_likelihoods (:217-341) on packed
    arrays, and the single-pair compute_read_likelihood_given_haplotype_log10 (:404-500) with the reference's argument
    list, asserts and messages.  Haplotype-prefix caching (hap_start_index, recache_read_values, next_haplotype_bases)
    is book-kept exactly as the reference does and, like its gkl arm, never used by the arithmetic."""

    def __init__(self, ctx):
        self.ctx = ctx
        self._log_likelihood_array = np.zeros(0)
        self.initialized = False
        self.max_read_length = 0
        self.max_haplotype_length = 0
        self.m_haplotype_data_array = []
        self.haplotype_to_haplotype_list_index_map = {}
        self.hap_start_index = None
        self.previous_haplotype_length = None
        self.do_exact_log10 = False
        self.logless = True

    @classmethod
    def initialize(cls, ctx, haplotypes, per_sample_read_list):
        """pair_hmm.rs:63-127.  haplotypes: list of byte strings in list order; per_sample_read_list: {sample: [read bases]}.
        Keeps the haplotypes and the haplotype -> list index map (the AVX arm) and the maximum lengths (the scalar arm)."""
        hmm = cls(ctx)
        hmm.m_haplotype_data_array = [bytes(h) for h in haplotypes]
        hmm.haplotype_to_haplotype_list_index_map = {h: i for i, h in enumerate(hmm.m_haplotype_data_array)}
        max_read = max((len(r) for reads in per_sample_read_list.values() for r in reads), default=0)
        max_hap = max((len(h) for h in hmm.m_haplotype_data_array), default=0)
        hmm._set_sizes(max_read, max_hap)
        hmm.initialized = True  # the AVX arm is initialised whatever the sizes (:88)
        return hmm

    @classmethod
    def quick_initialize(cls, ctx, max_read_length, haplotype_max_length):  # pair_hmm.rs:130-168
        hmm = cls(ctx)
        hmm._set_sizes(max_read_length, haplotype_max_length)
        return hmm

    def _set_sizes(self, max_read_length, haplotype_max_length):
        self.max_read_length = int(max_read_length)
        self.max_haplotype_length = int(haplotype_max_length)
        self.initialized = not (max_read_length == 0 and haplotype_max_length == 0)  # :138-141

    def reinitialize(self, max_read_length, haplotype_max_length):  # pair_hmm.rs:193-207 (sizes only, as the reference)
        self._set_sizes(max_read_length, haplotype_max_length)
        self.initialized = True

    def set_do_exact_log10(self, v):  # pair_hmm.rs:172-176, inert
        self.do_exact_log10 = bool(v)

    def set_logless(self, v):  # pair_hmm.rs:178-184, inert
        self.logless = bool(v)

    def do_not_use_tristate_correction(self):  # pair_hmm.rs:189-191
        self.ctx.set_tristate_correction(False)

    def compute_log10_likelihoods(self, batch):
        self._log_likelihood_array = self.ctx.pairhmm_batch(batch)
        return self._log_likelihood_array

    def compute_read_likelihood_given_haplotype_log10(self, haplotype_bases, read_bases, read_quals, insertion_gop,
                                                      deletion_gop, overall_gcp, recache_read_values=True,
                                                      next_haplotype_bases=None):
        """pair_hmm.rs:404-500: one (read, haplotype) pair through lk_pairhmm_forward."""
        assert self.initialized, "Must call initialize before calling compute_read_likelihood_given_haplotype_log10"
        assert len(haplotype_bases) <= self.max_haplotype_length, "Haplotype bases is too long"
        assert len(read_quals) == len(read_bases), "Read bases and read quals aren't the same size"
        assert len(insertion_gop) == len(read_bases), "Read bases and insertion gcp aren't the same size"
        assert len(deletion_gop) == len(read_bases), "Read bases and deletion gcp aren't the same size"
        assert len(overall_gcp) == len(read_bases), "Read bases and overal GCP aren't the same size"
        if recache_read_values:
            self.hap_start_index = 0
        next_start = 0
        if next_haplotype_bases is not None and len(next_haplotype_bases) == len(haplotype_bases):
            next_start = self.find_first_position_where_haplotypes_differ(haplotype_bases, next_haplotype_bases)
        assert self.hap_start_index is not None  # the reference unwraps it (:468)
        result = self.ctx.pairhmm_forward(haplotype_bases, read_bases, read_quals, insertion_gop, deletion_gop, overall_gcp)
        assert result <= 0.0, "PairHmm Log Probability cannot be greater than 0.0"
        self.hap_start_index = 0 if next_start < self.hap_start_index else next_start  # :487-496
        self.previous_haplotype_length = len(haplotype_bases)
        return result

    def get_log_likelihood_array(self):  # pair_hmm.rs:377-379
        return self._log_likelihood_array

    @staticmethod
    def find_first_position_where_haplotypes_differ(h1, h2):  # pair_hmm.rs:706-717 (host-side helper)
        n = min(len(h1), len(h2))
        a = np.frombuffer(bytes(h1[:n]), np.uint8) != np.frombuffer(bytes(h2[:n]), np.uint8)
        return int(np.argmax(a)) if a.any() else n


class SmithWatermanAligner:
    """SmithWatermanAligner::align (smith_waterman_aligner.rs:47-113): returns (cigar string, alignment_offset)."""

    def __init__(self, ctx):
        self.ctx = ctx

    def align(self, reference, alternate, parameters, overhang_strategy):
        off, cigar = self.ctx.sw_align(reference, alternate, parameters, overhang_strategy)
        return cigar, off


class DevicePool:
    """Several GPUs behind one handle in ONE process (lk_pool_*): the batch calls deal contiguous chunks of about
    equal cell counts to one host thread per device; RegionBatcher(pool, ...) batches per-region calls over all devices."""

    def __init__(self, devices=None):
        self._L = _lib.load()
        self._h = C.c_void_p()
        if devices is None:
            st = self._L.lk_pool_create(None, 0, C.byref(self._h))
        else:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            st = self._L.lk_pool_create(arr, len(devices), C.byref(self._h))
        if st != _lib.LK_OK:
            raise _lib.LorikeetError(st, "lk_pool_create failed (no CUDA device? there is no CPU fallback)")

    def _check(self, st):
        if st != _lib.LK_OK:
            raise _lib.LorikeetError(st, (self._L.lk_pool_last_error(self._h) or b"").decode())

    @property
    def device_count(self):
        return int(self._L.lk_pool_device_count(self._h))

    def pairhmm_batch(self, batch, want_flags=False):
        out = np.empty(batch.n_pairs, dtype=np.float64)
        flags = np.zeros(batch.n_pairs, dtype=np.uint8)
        s = batch.as_struct()
        self._check(self._L.lk_pool_pairhmm_batch_host(self._h, C.byref(s), out.ctypes.data_as(_lib.f64p), flags.ctypes.data_as(_lib.u8p)))
        return (out, flags) if want_flags else out

    def compute_read_likelihoods(self, engine, batch, read_mapq, ref_allele_index=None, values=None, keep=None):
        """lk_pool_compute_read_likelihoods with the arguments of `engine` (a PairHMMLikelihoodCalculationEngine)."""
        if values is None:
            values = np.empty(batch.n_pairs, dtype=np.float64)
        if keep is None:
            keep = np.ones(batch.n_reads, dtype=np.uint8)
        mapq = np.ascontiguousarray(read_mapq, dtype=np.uint8)
        ref = None if ref_allele_index is None else np.ascontiguousarray(ref_allele_index, dtype=np.int32)
        s = batch.as_struct()
        self._check(self._L.lk_pool_compute_read_likelihoods(self._h, C.byref(engine.args), C.byref(s), mapq.ctypes.data_as(_lib.u8p),
                                                             ref.ctypes.data_as(_lib.i32p) if ref is not None else C.cast(None, _lib.i32p),
                                                             values.ctypes.data_as(_lib.f64p), keep.ctypes.data_as(_lib.u8p)))
        return values, keep

    def close(self):
        if self._h:
            self._L.lk_pool_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class RegionBatcher:
    """Cross-thread batching front-end (lk_batcher, include/lorikeet_b200.h): worker threads submit ONE region each, as
    Lorikeet's Rayon tasks do (assembly_region_walker.rs:140-277); a dispatcher thread inside the library runs one launch
    for everything pending.  engine=None batches bare PairHMM calls."""

    def __init__(self, ctx, engine=None, max_regions=256, max_wait_us=200):
        """ctx: a Context (one device) or a DevicePool (the dispatcher threads of all its devices share one queue)."""
        self.ctx = ctx
        self.engine = engine
        self._L = _lib.load()
        self._h = C.c_void_p()
        args = C.byref(engine.args) if engine is not None else None
        create = self._L.lk_pool_batcher_create if isinstance(ctx, DevicePool) else self._L.lk_batcher_create
        ctx._check(create(ctx._h, args, int(max_regions), int(max_wait_us), C.byref(self._h)))

    def submit(self, region, read_mapq=None, ref_allele_index=0):
        """Blocking; returns (values, keep) for an engine batcher, (log10 likelihoods [r*A + h], flags) otherwise."""
        assert region.n_regions == 1
        values = np.empty(region.n_pairs, dtype=np.float64)
        aux = np.zeros(region.n_reads if self.engine is not None else region.n_pairs, dtype=np.uint8)
        mapq = None if read_mapq is None else np.ascontiguousarray(read_mapq, dtype=np.uint8)
        s = region.as_struct()
        st = self._L.lk_batcher_submit(self._h, C.byref(s), mapq.ctypes.data_as(_lib.u8p) if mapq is not None else None,
                                       int(ref_allele_index), values.ctypes.data_as(_lib.f64p), aux.ctypes.data_as(_lib.u8p))
        self.ctx._check(st)
        return values, aux

    def submit_async(self, region, read_mapq=None, ref_allele_index=0):
        """Non-blocking (lk_batcher_submit_async): returns a ticket for wait(); several may be in flight per thread."""
        assert region.n_regions == 1
        values = np.empty(region.n_pairs, dtype=np.float64)
        aux = np.zeros(region.n_reads if self.engine is not None else region.n_pairs, dtype=np.uint8)
        mapq = None if read_mapq is None else np.ascontiguousarray(read_mapq, dtype=np.uint8)
        s = region.as_struct()
        t = C.c_void_p()
        st = self._L.lk_batcher_submit_async(self._h, C.byref(s), mapq.ctypes.data_as(_lib.u8p) if mapq is not None else None,
                                             int(ref_allele_index), values.ctypes.data_as(_lib.f64p), aux.ctypes.data_as(_lib.u8p),
                                             C.byref(t))
        self.ctx._check(st)
        return (t, values, aux, region, s, mapq)  # keeps every buffer the library reads or writes alive until wait()

    def wait(self, ticket):
        """Blocks until the ticket's region is done; returns (values, keep) / (log10 likelihoods, flags) as submit()."""
        t, values, aux = ticket[0], ticket[1], ticket[2]
        self.ctx._check(self._L.lk_batcher_wait(self._h, t))
        return values, aux

    def stats(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self._L.lk_batcher_stats(self._h, C.byref(a), C.byref(b))
        return a.value, b.value

    def close(self):
        if self._h:
            self._L.lk_batcher_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
