"""ctypes binding of liblorikeet_b200.so (the C ABI of include/lorikeet_b200.h).

The library is the product; this module only loads it and declares signatures.  There is no
Python or CPU fallback: if the shared object is missing or no CUDA device is present, calls fail
loudly (LorikeetError).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("LK_SO") or os.path.join(HERE, "liblorikeet_b200.so")  # LK_SO: developer override (kernel variants)

u8p = C.POINTER(C.c_uint8)
i32p = C.POINTER(C.c_int32)
u32p = C.POINTER(C.c_uint32)
i64p = C.POINTER(C.c_int64)
f64p = C.POINTER(C.c_double)
f32p = C.POINTER(C.c_float)

LK_OK, LK_ERR_INVALID, LK_ERR_CUDA, LK_ERR_NOMEM, LK_ERR_CAPACITY = 0, 1, 2, 3, 4
STATUS_NAMES = {0: "LK_OK", 1: "LK_ERR_INVALID", 2: "LK_ERR_CUDA", 3: "LK_ERR_NOMEM", 4: "LK_ERR_CAPACITY"}

# the entry points include/lorikeet_b200.h declares
EXPORTED_SYMBOLS = [
    "lk_create", "lk_destroy", "lk_last_error", "lk_version", "lk_set_tristate_correction", "lk_set_force_f64",
    "lk_pairhmm_batch_dims", "lk_pairhmm_batch_host", "lk_pairhmm_upload", "lk_pairhmm_run_resident",
    "lk_pairhmm_download", "lk_resident_free", "lk_pairhmm_forward",
    "lk_sw_batch_host", "lk_sw_upload", "lk_sw_run_resident", "lk_sw_download", "lk_sw_align",
    "lk_calculate_cigar_batch", "lk_calculate_cigar_shortcut", "lk_calculate_cigar_finish",
    "lk_engine_modify_read_qualities", "lk_engine_modify_read_qualities_device", "lk_engine_compute_read_likelihoods",
    "lk_realign_reads_batch", "lk_best_alleles_breaking_ties", "lk_realign_finish", "lk_cigar_apply_to_cigar",
    "lk_cigar_append_clipped_elements", "lk_read_start_on_reference_haplotype",
    "lk_batcher_create", "lk_batcher_destroy", "lk_batcher_submit", "lk_batcher_stats", "lk_batcher_submit_async",
    "lk_batcher_wait",
    "lk_host_register", "lk_host_unregister", "lk_engine_normalize_and_filter_device", "lk_dangling_end_sw_batch", "lk_find_number_of_repetitions",
    "lk_pool_create", "lk_pool_destroy", "lk_pool_device_count", "lk_pool_ctx", "lk_pool_last_error",
    "lk_pool_set_tristate_correction", "lk_pool_pairhmm_batch_host", "lk_pool_compute_read_likelihoods", "lk_pool_batcher_create",
    "lk_resident_pairs", "lk_resident_cells", "lk_resident_launch_count", "lk_resident_last_fast_ms", "lk_resident_sw_packed_pairs",
    "lk_resident_out_device_ptr", "lk_debug_last_launch_count", "lk_measure_fp32_fma_tflops",
]


class LorikeetError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), message))
        self.status = status


class PairHmmBatchStruct(C.Structure):
    _fields_ = [
        ("n_regions", C.c_int64),
        ("region_read_start", i64p),
        ("region_hap_start", i64p),
        ("read_off", i64p),
        ("read_bases", u8p),
        ("read_qual", u8p),
        ("read_ins", u8p),
        ("read_del", u8p),
        ("read_gcp", u8p),
        ("gcp_constant", C.c_uint8),
        ("hap_off", i64p),
        ("hap_bases", u8p),
        ("region_sample_start", i64p),
        ("sample_read_start", i64p),
    ]


class SwParams(C.Structure):
    _fields_ = [("match_value", C.c_int32), ("mismatch_penalty", C.c_int32),
                ("gap_open_penalty", C.c_int32), ("gap_extend_penalty", C.c_int32)]


class EngineArgs(C.Structure):
    _fields_ = [
        ("constant_gcp", C.c_uint8),
        ("log10_global_read_mismapping_rate", C.c_double),
        ("pcr_error_model", C.c_int32),
        ("base_quality_score_threshold", C.c_uint8),
        ("dynamic_read_disqualification", C.c_int32),
        ("read_disqualification_scale", C.c_double),
        ("expected_error_rate_per_base", C.c_double),
        ("symmetrically_normalize_alleles_to_reference", C.c_int32),
        ("disable_cap_read_qualities_to_mapq", C.c_int32),
        ("modify_soft_clipped_bases", C.c_int32),
    ]


class RealignBatchStruct(C.Structure):
    _fields_ = [
        ("n_regions", C.c_int64),
        ("region_read_start", i64p),
        ("region_hap_start", i64p),
        ("values", f64p),
        ("region_n_evidence", i64p),
        ("ref_allele_index", i32p),
        ("reference_start", i64p),
        ("hap_off", i64p),
        ("hap_bases", u8p),
        ("hap_cigar_off", i64p),
        ("hap_cigars", u32p),
        ("hap_start_wrt_ref", i64p),
        ("read_off", i64p),
        ("read_bases", u8p),
        ("read_cigar_off", i64p),
        ("read_cigars", u32p),
        ("read_orig_len", i64p),
        ("informative_threshold", C.c_double),
        ("region_sample_start", i64p),
        ("sample_read_start", i64p),
        ("sample_n_evidence", i64p),
    ]


_LIB = None


def load():
    """Load the shared library (once).  Raises if it has not been built: run __graft_entry__.build()."""
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        raise LorikeetError(LK_ERR_CUDA, "liblorikeet_b200.so not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'`" % SO_PATH)
    L = C.CDLL(SO_PATH)
    vp = C.c_void_p
    L.lk_create.restype = C.c_int
    L.lk_create.argtypes = [C.c_int, C.POINTER(vp)]
    L.lk_destroy.restype = None
    L.lk_destroy.argtypes = [vp]
    L.lk_last_error.restype = C.c_char_p
    L.lk_last_error.argtypes = [vp]
    L.lk_version.restype = C.c_char_p
    L.lk_set_tristate_correction.restype = C.c_int
    L.lk_set_tristate_correction.argtypes = [vp, C.c_int]
    L.lk_set_force_f64.restype = C.c_int
    L.lk_set_force_f64.argtypes = [vp, C.c_int]
    bp = C.POINTER(PairHmmBatchStruct)
    L.lk_pairhmm_batch_dims.restype = C.c_int
    L.lk_pairhmm_batch_dims.argtypes = [bp, i64p, i64p]
    L.lk_pairhmm_batch_host.restype = C.c_int
    L.lk_pairhmm_batch_host.argtypes = [vp, bp, f64p, u8p]
    L.lk_pairhmm_upload.restype = C.c_int
    L.lk_pairhmm_upload.argtypes = [vp, bp, C.POINTER(vp)]
    L.lk_pairhmm_run_resident.restype = C.c_int
    L.lk_pairhmm_run_resident.argtypes = [vp, vp, f32p]
    L.lk_pairhmm_download.restype = C.c_int
    L.lk_pairhmm_download.argtypes = [vp, vp, f64p, u8p]
    L.lk_resident_free.restype = None
    L.lk_resident_free.argtypes = [vp, vp]
    L.lk_pairhmm_forward.restype = C.c_int
    L.lk_pairhmm_forward.argtypes = [vp, u8p, C.c_int64, u8p, u8p, u8p, u8p, u8p, C.c_int64, f64p]
    sp = C.POINTER(SwParams)
    L.lk_sw_batch_host.restype = C.c_int
    L.lk_sw_batch_host.argtypes = [vp, C.c_int64, i64p, u8p, i64p, u8p, sp, C.c_int, C.c_int32, u32p, i32p, i32p]
    L.lk_sw_upload.restype = C.c_int
    L.lk_sw_upload.argtypes = [vp, C.c_int64, i64p, u8p, i64p, u8p, sp, C.c_int, C.c_int32, C.POINTER(vp)]
    L.lk_sw_run_resident.restype = C.c_int
    L.lk_sw_run_resident.argtypes = [vp, vp, f32p]
    L.lk_sw_download.restype = C.c_int
    L.lk_sw_download.argtypes = [vp, vp, u32p, i32p, i32p]
    L.lk_sw_align.restype = C.c_int
    L.lk_sw_align.argtypes = [vp, u8p, C.c_int64, u8p, C.c_int64, sp, C.c_int, C.c_char_p, C.c_size_t, i32p]
    L.lk_calculate_cigar_batch.restype = C.c_int
    L.lk_calculate_cigar_batch.argtypes = [vp, C.c_int64, i64p, u8p, i64p, u8p, sp, C.c_int, C.c_int32, u32p, i32p]
    L.lk_calculate_cigar_shortcut.restype = C.c_int
    L.lk_calculate_cigar_shortcut.argtypes = [u8p, C.c_int64, u8p, C.c_int64, C.c_int32, u32p, i32p]
    L.lk_calculate_cigar_finish.restype = C.c_int
    L.lk_calculate_cigar_finish.argtypes = [u8p, C.c_int64, u8p, C.c_int64, u32p, C.c_int32, C.c_int32, C.c_int32, u32p, i32p]
    ep = C.POINTER(EngineArgs)
    L.lk_engine_modify_read_qualities.restype = C.c_int
    L.lk_engine_modify_read_qualities.argtypes = [ep, C.c_int64, i64p, u8p, u8p, u8p, u8p, u8p, C.c_int]
    L.lk_engine_modify_read_qualities_device.restype = C.c_int
    L.lk_engine_modify_read_qualities_device.argtypes = [vp, ep, C.c_int64, i64p, u8p, u8p, u8p, u8p, u8p]
    L.lk_engine_compute_read_likelihoods.restype = C.c_int
    L.lk_engine_compute_read_likelihoods.argtypes = [vp, ep, bp, u8p, i32p, f64p, u8p]
    L.lk_realign_reads_batch.restype = C.c_int
    L.lk_realign_reads_batch.argtypes = [vp, C.POINTER(RealignBatchStruct), C.c_int32, i32p, u8p, u32p, i32p, i64p]
    L.lk_best_alleles_breaking_ties.restype = C.c_int
    L.lk_best_alleles_breaking_ties.argtypes = [f64p, C.c_int64, C.c_int64, C.c_int64, i32p, C.c_double, i32p, f64p, u8p]
    L.lk_realign_finish.restype = C.c_int
    L.lk_realign_finish.argtypes = [u8p, C.c_int64, u32p, C.c_int32, C.c_int64, u32p, C.c_int32, C.c_int32, u32p, C.c_int32,
                                    C.c_int64, u8p, C.c_int64, C.c_int64, C.c_int32, u32p, i32p, i64p]
    L.lk_cigar_apply_to_cigar.restype = C.c_int
    L.lk_cigar_apply_to_cigar.argtypes = [u32p, C.c_int32, u32p, C.c_int32, C.c_int32, u32p, i32p]
    L.lk_cigar_append_clipped_elements.restype = C.c_int
    L.lk_cigar_append_clipped_elements.argtypes = [u32p, C.c_int32, u32p, C.c_int32, C.c_int32, u32p, i32p]
    L.lk_read_start_on_reference_haplotype.restype = C.c_int64
    L.lk_read_start_on_reference_haplotype.argtypes = [u32p, C.c_int32, C.c_uint32]
    L.lk_batcher_create.restype = C.c_int
    L.lk_batcher_create.argtypes = [vp, ep, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.lk_batcher_destroy.restype = None
    L.lk_batcher_destroy.argtypes = [vp]
    L.lk_batcher_submit.restype = C.c_int
    L.lk_batcher_submit.argtypes = [vp, bp, u8p, C.c_int32, f64p, u8p]
    L.lk_batcher_submit_async.restype = C.c_int
    L.lk_batcher_submit_async.argtypes = [vp, bp, u8p, C.c_int32, f64p, u8p, C.POINTER(vp)]
    L.lk_batcher_wait.restype = C.c_int
    L.lk_batcher_wait.argtypes = [vp, vp]
    L.lk_batcher_stats.restype = None
    L.lk_batcher_stats.argtypes = [vp, i64p, i64p]
    L.lk_engine_normalize_and_filter_device.restype = C.c_int
    L.lk_engine_normalize_and_filter_device.argtypes = [vp, ep, bp, i32p, f64p, f64p, u8p]
    L.lk_host_register.restype = C.c_int
    L.lk_host_register.argtypes = [vp, C.c_size_t]
    L.lk_host_unregister.restype = C.c_int
    L.lk_host_unregister.argtypes = [vp]
    L.lk_dangling_end_sw_batch.restype = C.c_int
    L.lk_dangling_end_sw_batch.argtypes = [vp, C.c_int64, i64p, u8p, i64p, u8p, C.POINTER(SwParams), C.c_int32, u32p, i32p]
    L.lk_find_number_of_repetitions.restype = C.c_int64
    L.lk_find_number_of_repetitions.argtypes = [u8p, C.c_int64, u8p, C.c_int64, C.c_int]
    # several devices behind one handle
    L.lk_pool_create.restype = C.c_int
    L.lk_pool_create.argtypes = [C.POINTER(C.c_int), C.c_int32, C.POINTER(vp)]
    L.lk_pool_destroy.restype = None
    L.lk_pool_destroy.argtypes = [vp]
    L.lk_pool_device_count.restype = C.c_int32
    L.lk_pool_device_count.argtypes = [vp]
    L.lk_pool_ctx.restype = vp
    L.lk_pool_ctx.argtypes = [vp, C.c_int32]
    L.lk_pool_last_error.restype = C.c_char_p
    L.lk_pool_last_error.argtypes = [vp]
    L.lk_pool_set_tristate_correction.restype = C.c_int
    L.lk_pool_set_tristate_correction.argtypes = [vp, C.c_int]
    L.lk_pool_pairhmm_batch_host.restype = C.c_int
    L.lk_pool_pairhmm_batch_host.argtypes = [vp, bp, f64p, u8p]
    L.lk_pool_compute_read_likelihoods.restype = C.c_int
    L.lk_pool_compute_read_likelihoods.argtypes = [vp, ep, bp, u8p, i32p, f64p, u8p]
    L.lk_pool_batcher_create.restype = C.c_int
    L.lk_pool_batcher_create.argtypes = [vp, ep, C.c_int32, C.c_int32, C.POINTER(vp)]
    # introspection helpers
    L.lk_resident_launch_count.restype = C.c_int32
    L.lk_resident_launch_count.argtypes = [vp]
    L.lk_resident_cells.restype = C.c_int64
    L.lk_resident_cells.argtypes = [vp]
    L.lk_resident_pairs.restype = C.c_int64
    L.lk_resident_pairs.argtypes = [vp]
    L.lk_resident_sw_packed_pairs.restype = C.c_int64
    L.lk_resident_sw_packed_pairs.argtypes = [vp]
    L.lk_resident_last_fast_ms.restype = C.c_float
    L.lk_resident_last_fast_ms.argtypes = [vp]
    L.lk_resident_out_device_ptr.restype = C.c_void_p
    L.lk_resident_out_device_ptr.argtypes = [vp]
    L.lk_measure_fp32_fma_tflops.restype = C.c_int
    L.lk_measure_fp32_fma_tflops.argtypes = [vp, C.c_double, f64p]
    _LIB = L
    return L
