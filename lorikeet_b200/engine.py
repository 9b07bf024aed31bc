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
        assert isinstance(ctx, Context)
        self.ctx = ctx
        self.args = _lib.EngineArgs(int(constant_gcp), float(log10_global_read_mismapping_rate), int(pcr_error_model),
                                    int(base_quality_score_threshold), 1 if dynamic_read_disqualification else 0,
                                    float(read_disqualification_scale), float(expected_error_rate_per_base),
                                    1 if symmetrically_normalize_alleles_to_reference else 0,
                                    1 if disable_cap_read_qualities_to_mapq else 0, 1 if modify_soft_clipped_bases else 0)

    def compute_read_likelihoods(self, batch, read_mapq, ref_allele_index=None, values=None, keep=None):
        """compute_read_likelihoods (:195-242) for a batch of single-sample regions.
        Returns (values, keep): values holds per region an [A x R] row-major block ([allele][read], the reference's
        AlleleLikelihoods layout) at the pair offsets of the batch, normalised and filtered; keep[r] = read survived."""
        L = _lib.load()
        n = batch.n_pairs
        if values is None:
            values = np.empty(n, dtype=np.float64)
        if keep is None:
            keep = np.ones(batch.n_reads, dtype=np.uint8)
        assert values.dtype == np.float64 and values.size == n and keep.dtype == np.uint8 and keep.size == batch.n_reads
        mapq = np.ascontiguousarray(read_mapq, dtype=np.uint8)
        assert mapq.size == batch.n_reads
        ref = None
        if ref_allele_index is not None:
            ref = np.ascontiguousarray(ref_allele_index, dtype=np.int32)
        s = batch.as_struct()
        st = L.lk_engine_compute_read_likelihoods(self.ctx._h, C.byref(self.args), C.byref(s), mapq.ctypes.data_as(_lib.u8p),
                                                  ref.ctypes.data_as(_lib.i32p) if ref is not None else C.cast(None, _lib.i32p),
                                                  values.ctypes.data_as(_lib.f64p), keep.ctypes.data_as(_lib.u8p))
        self.ctx._check(st)
        return values, keep


class PairHMM:
    """PairHMM::initialize + compute_log10_likelihoods (pair_hmm.rs:63-341) on packed arrays."""

    def __init__(self, ctx):
        self.ctx = ctx
        self._log_likelihood_array = np.zeros(0)

    def do_not_use_tristate_correction(self):  # pair_hmm.rs:189-191
        self.ctx.set_tristate_correction(False)

    def compute_log10_likelihoods(self, batch):
        self._log_likelihood_array = self.ctx.pairhmm_batch(batch)
        return self._log_likelihood_array

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


class RegionBatcher:
    """Cross-thread batching front-end (lk_batcher, include/lorikeet_b200.h): worker threads submit ONE region each, as
    Lorikeet's Rayon tasks do (assembly_region_walker.rs:140-277); a dispatcher thread inside the library runs one launch
    for everything pending.  engine=None batches bare PairHMM calls."""

    def __init__(self, ctx, engine=None, max_regions=256, max_wait_us=200):
        self.ctx = ctx
        self.engine = engine
        self._L = _lib.load()
        self._h = C.c_void_p()
        args = C.byref(engine.args) if engine is not None else None
        ctx._check(self._L.lk_batcher_create(ctx._h, args, int(max_regions), int(max_wait_us), C.byref(self._h)))

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
