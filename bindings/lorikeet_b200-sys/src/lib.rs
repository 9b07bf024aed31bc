//! Raw bindings of `include/lorikeet_b200.h`, declaration for declaration and in the header's order.
//!
//! NOT compiled in the build container of this repository (no `cargo` / `rustc` there); `tests/test_bindings.py`
//! checks this file against the header instead: every function the header declares is declared here with the same
//! number of arguments, and every struct has the same fields in the same order.  The C ABI itself is exercised end to
//! end through ctypes by `tests/`.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct lk_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct lk_resident {
    _private: [u8; 0],
}
#[repr(C)]
pub struct lk_batcher {
    _private: [u8; 0],
}
#[repr(C)]
pub struct lk_ticket {
    _private: [u8; 0],
}
#[repr(C)]
pub struct lk_pool {
    _private: [u8; 0],
}

/// `enum lk_status`
pub type lk_status = c_int;
pub const LK_OK: lk_status = 0;
pub const LK_ERR_INVALID: lk_status = 1;
pub const LK_ERR_CUDA: lk_status = 2;
pub const LK_ERR_NOMEM: lk_status = 3;
pub const LK_ERR_CAPACITY: lk_status = 4;

/// `enum lk_sw_overhang` — replaces `gkl::smithwaterman::OverhangStrategy`
pub type lk_sw_overhang = c_int;
pub const LK_SW_SOFTCLIP: lk_sw_overhang = 0;
pub const LK_SW_INDEL: lk_sw_overhang = 1;
pub const LK_SW_LEADING_INDEL: lk_sw_overhang = 2;
pub const LK_SW_IGNORE: lk_sw_overhang = 3;

/// per-pair result flags of the PairHMM entry points
pub const LK_FLAG_RESCUED: u8 = 1;
pub const LK_FLAG_NEG_INF: u8 = 2;

/// `enum lk_pcr_model` — `PCRErrorModel`
pub type lk_pcr_model = c_int;
pub const LK_PCR_NONE: lk_pcr_model = 0;
pub const LK_PCR_HOSTILE: lk_pcr_model = 1;
pub const LK_PCR_AGGRESSIVE: lk_pcr_model = 2;
pub const LK_PCR_CONSERVATIVE: lk_pcr_model = 3;

/// replaces `gkl::smithwaterman::Parameters`
#[repr(C)]
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub struct lk_sw_params {
    pub match_value: i32,
    pub mismatch_penalty: i32,
    pub gap_open_penalty: i32,
    pub gap_extend_penalty: i32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct lk_pairhmm_batch {
    pub n_regions: i64,
    pub region_read_start: *const i64,
    pub region_hap_start: *const i64,
    pub read_off: *const i64,
    pub read_bases: *const u8,
    pub read_qual: *const u8,
    pub read_ins: *const u8,
    pub read_del: *const u8,
    pub read_gcp: *const u8,
    pub gcp_constant: u8,
    pub hap_off: *const i64,
    pub hap_bases: *const u8,
    pub region_sample_start: *const i64,
    pub sample_read_start: *const i64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct lk_realign_batch {
    pub n_regions: i64,
    pub region_read_start: *const i64,
    pub region_hap_start: *const i64,
    pub values: *const f64,
    pub region_n_evidence: *const i64,
    pub ref_allele_index: *const i32,
    pub reference_start: *const i64,
    pub hap_off: *const i64,
    pub hap_bases: *const u8,
    pub hap_cigar_off: *const i64,
    pub hap_cigars: *const u32,
    pub hap_start_wrt_ref: *const i64,
    pub read_off: *const i64,
    pub read_bases: *const u8,
    pub read_cigar_off: *const i64,
    pub read_cigars: *const u32,
    pub read_orig_len: *const i64,
    pub informative_threshold: f64,
    pub region_sample_start: *const i64,
    pub sample_read_start: *const i64,
    pub sample_n_evidence: *const i64,
}

/// arguments of `PairHMMLikelihoodCalculationEngine::new`, in the same order
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct lk_engine_args {
    pub constant_gcp: u8,
    pub log10_global_read_mismapping_rate: f64,
    pub pcr_error_model: i32,
    pub base_quality_score_threshold: u8,
    pub dynamic_read_disqualification: i32,
    pub read_disqualification_scale: f64,
    pub expected_error_rate_per_base: f64,
    pub symmetrically_normalize_alleles_to_reference: i32,
    pub disable_cap_read_qualities_to_mapq: i32,
    pub modify_soft_clipped_bases: i32,
}

extern "C" {
    // ---- context
    pub fn lk_create(device: c_int, out_ctx: *mut *mut lk_ctx) -> lk_status;
    pub fn lk_destroy(ctx: *mut lk_ctx);
    pub fn lk_last_error(ctx: *const lk_ctx) -> *const c_char;
    pub fn lk_version() -> *const c_char;
    pub fn lk_set_tristate_correction(ctx: *mut lk_ctx, use_tristate: c_int) -> lk_status;
    pub fn lk_set_force_f64(ctx: *mut lk_ctx, force_f64: c_int) -> lk_status;

    pub fn lk_host_register(ptr: *mut c_void, bytes: usize) -> lk_status;
    pub fn lk_host_unregister(ptr: *mut c_void) -> lk_status;

    // ---- PairHMM
    pub fn lk_pairhmm_batch_dims(batch: *const lk_pairhmm_batch, n_pairs: *mut i64, n_cells: *mut i64) -> lk_status;
    pub fn lk_pairhmm_batch_host(ctx: *mut lk_ctx, batch: *const lk_pairhmm_batch, out: *mut f64, flags: *mut u8) -> lk_status;
    pub fn lk_pairhmm_upload(ctx: *mut lk_ctx, batch: *const lk_pairhmm_batch, out_res: *mut *mut lk_resident) -> lk_status;
    pub fn lk_pairhmm_run_resident(ctx: *mut lk_ctx, res: *mut lk_resident, kernel_ms: *mut f32) -> lk_status;
    pub fn lk_pairhmm_download(ctx: *mut lk_ctx, res: *mut lk_resident, out: *mut f64, flags: *mut u8) -> lk_status;
    pub fn lk_resident_free(ctx: *mut lk_ctx, res: *mut lk_resident);
    pub fn lk_resident_pairs(res: *const lk_resident) -> i64;
    pub fn lk_resident_cells(res: *const lk_resident) -> i64;
    pub fn lk_resident_launch_count(res: *const lk_resident) -> i32;
    pub fn lk_resident_last_fast_ms(res: *const lk_resident) -> f32;
    pub fn lk_resident_sw_packed_pairs(res: *const lk_resident) -> i64;
    pub fn lk_resident_out_device_ptr(res: *const lk_resident) -> *mut c_void;
    pub fn lk_debug_last_launch_count(ctx: *const lk_ctx) -> i32;
    pub fn lk_measure_fp32_fma_tflops(ctx: *mut lk_ctx, duration_ms: f64, tflops: *mut f64) -> lk_status;
    pub fn lk_pairhmm_forward(
        ctx: *mut lk_ctx,
        hap: *const u8,
        hap_len: i64,
        read: *const u8,
        qual: *const u8,
        ins: *const u8,
        del: *const u8,
        gcp: *const u8,
        read_len: i64,
        out_log10: *mut f64,
    ) -> lk_status;

    // ---- Smith-Waterman
    pub fn lk_sw_batch_host(
        ctx: *mut lk_ctx,
        n_pairs: i64,
        ref_off: *const i64,
        refs: *const u8,
        alt_off: *const i64,
        alts: *const u8,
        params: *const lk_sw_params,
        strategy: lk_sw_overhang,
        max_ops: i32,
        cigar_ops: *mut u32,
        n_ops: *mut i32,
        alignment_offset: *mut i32,
    ) -> lk_status;
    pub fn lk_sw_upload(
        ctx: *mut lk_ctx,
        n_pairs: i64,
        ref_off: *const i64,
        refs: *const u8,
        alt_off: *const i64,
        alts: *const u8,
        params: *const lk_sw_params,
        strategy: lk_sw_overhang,
        max_ops: i32,
        out_res: *mut *mut lk_resident,
    ) -> lk_status;
    pub fn lk_sw_run_resident(ctx: *mut lk_ctx, res: *mut lk_resident, kernel_ms: *mut f32) -> lk_status;
    pub fn lk_sw_download(ctx: *mut lk_ctx, res: *mut lk_resident, cigar_ops: *mut u32, n_ops: *mut i32, alignment_offset: *mut i32) -> lk_status;
    pub fn lk_sw_align(
        ctx: *mut lk_ctx,
        reference: *const u8,
        ref_len: i64,
        alt: *const u8,
        alt_len: i64,
        params: *const lk_sw_params,
        strategy: lk_sw_overhang,
        cigar_text: *mut c_char,
        cigar_cap: usize,
        alignment_offset: *mut i32,
    ) -> lk_status;

    // ---- haplotype -> reference CIGARs (CigarUtils::calculate_cigar)
    pub fn lk_calculate_cigar_batch(
        ctx: *mut lk_ctx,
        n_pairs: i64,
        ref_off: *const i64,
        refs: *const u8,
        alt_off: *const i64,
        alts: *const u8,
        params: *const lk_sw_params,
        strategy: lk_sw_overhang,
        max_ops: i32,
        cigar_ops: *mut u32,
        n_ops: *mut i32,
    ) -> lk_status;
    pub fn lk_calculate_cigar_shortcut(
        reference: *const u8,
        ref_len: i64,
        alt: *const u8,
        alt_len: i64,
        max_ops: i32,
        out_ops: *mut u32,
        out_n_ops: *mut i32,
    ) -> c_int;
    pub fn lk_calculate_cigar_finish(
        reference: *const u8,
        ref_len: i64,
        alt: *const u8,
        alt_len: i64,
        sw_ops: *const u32,
        sw_n_ops: i32,
        sw_offset: i32,
        max_ops: i32,
        out_ops: *mut u32,
        out_n_ops: *mut i32,
    ) -> lk_status;

    // ---- dangling-end merges of the read-threading graph
    pub fn lk_dangling_end_sw_batch(
        ctx: *mut lk_ctx,
        n_pairs: i64,
        ref_off: *const i64,
        refs: *const u8,
        alt_off: *const i64,
        alts: *const u8,
        params: *const lk_sw_params,
        max_ops: i32,
        cigar_ops: *mut u32,
        n_ops: *mut i32,
    ) -> lk_status;

    // ---- read -> best-haplotype realignment
    pub fn lk_realign_reads_batch(
        ctx: *mut lk_ctx,
        batch: *const lk_realign_batch,
        max_ops: i32,
        best_allele: *mut i32,
        informative: *mut u8,
        cigar_ops: *mut u32,
        n_ops: *mut i32,
        pos: *mut i64,
    ) -> lk_status;
    pub fn lk_best_alleles_breaking_ties(
        values: *const f64,
        n_alleles: i64,
        n_reads: i64,
        n_evidence: i64,
        priorities: *const i32,
        informative_threshold: f64,
        best_allele: *mut i32,
        confidence: *mut f64,
        informative: *mut u8,
    ) -> lk_status;
    pub fn lk_realign_finish(
        read: *const u8,
        read_len: i64,
        orig_cigar: *const u32,
        orig_n: i32,
        orig_read_len: i64,
        sw_ops: *const u32,
        sw_n_ops: i32,
        sw_offset: i32,
        hap_cigar: *const u32,
        hap_cigar_n: i32,
        hap_start_wrt_ref: i64,
        ref_hap: *const u8,
        ref_hap_len: i64,
        reference_start: i64,
        max_ops: i32,
        out_ops: *mut u32,
        out_n_ops: *mut i32,
        out_pos: *mut i64,
    ) -> lk_status;
    pub fn lk_cigar_apply_to_cigar(
        first_to_second: *const u32,
        n12: i32,
        second_to_third: *const u32,
        n23: i32,
        max_ops: i32,
        out_ops: *mut u32,
        out_n_ops: *mut i32,
    ) -> lk_status;
    pub fn lk_cigar_append_clipped_elements(
        cigar: *const u32,
        n: i32,
        original: *const u32,
        n_orig: i32,
        max_ops: i32,
        out_ops: *mut u32,
        out_n_ops: *mut i32,
    ) -> lk_status;
    pub fn lk_read_start_on_reference_haplotype(hap_vs_ref_cigar: *const u32, n: i32, read_start_on_haplotype: u32) -> i64;

    // ---- engine pre/post-ops
    pub fn lk_engine_modify_read_qualities(
        args: *const lk_engine_args,
        n_reads: i64,
        read_off: *const i64,
        read_bases: *const u8,
        read_qual: *mut u8,
        read_ins: *mut u8,
        read_del: *mut u8,
        read_mapq: *const u8,
        n_threads: c_int,
    ) -> lk_status;
    pub fn lk_engine_modify_read_qualities_device(
        ctx: *mut lk_ctx,
        args: *const lk_engine_args,
        n_reads: i64,
        read_off: *const i64,
        read_bases: *const u8,
        read_qual: *mut u8,
        read_ins: *mut u8,
        read_del: *mut u8,
        read_mapq: *const u8,
    ) -> lk_status;
    pub fn lk_engine_compute_read_likelihoods(
        ctx: *mut lk_ctx,
        args: *const lk_engine_args,
        batch: *const lk_pairhmm_batch,
        read_mapq: *const u8,
        ref_allele_index: *const i32,
        values: *mut f64,
        keep: *mut u8,
    ) -> lk_status;
    pub fn lk_engine_normalize_and_filter_device(
        ctx: *mut lk_ctx,
        args: *const lk_engine_args,
        batch: *const lk_pairhmm_batch,
        ref_allele_index: *const i32,
        likelihoods: *const f64,
        values: *mut f64,
        keep: *mut u8,
    ) -> lk_status;
    pub fn lk_find_number_of_repetitions(
        repeat_unit: *const u8,
        unit_len: i64,
        test_string: *const u8,
        test_len: i64,
        leading_repeats: c_int,
    ) -> i64;

    // ---- cross-thread batching front-end
    pub fn lk_batcher_create(ctx: *mut lk_ctx, args: *const lk_engine_args, max_regions: i32, max_wait_us: i32, out: *mut *mut lk_batcher) -> lk_status;
    pub fn lk_batcher_destroy(b: *mut lk_batcher);
    pub fn lk_batcher_submit(
        b: *mut lk_batcher,
        region: *const lk_pairhmm_batch,
        read_mapq: *const u8,
        ref_allele_index: i32,
        values: *mut f64,
        keep_or_flags: *mut u8,
    ) -> lk_status;
    pub fn lk_batcher_stats(b: *mut lk_batcher, n_batches: *mut i64, n_regions: *mut i64);
    pub fn lk_batcher_submit_async(
        b: *mut lk_batcher,
        region: *const lk_pairhmm_batch,
        read_mapq: *const u8,
        ref_allele_index: i32,
        values: *mut f64,
        keep_or_flags: *mut u8,
        out_ticket: *mut *mut lk_ticket,
    ) -> lk_status;
    pub fn lk_batcher_wait(b: *mut lk_batcher, ticket: *mut lk_ticket) -> lk_status;

    // ---- several devices in one process
    pub fn lk_pool_create(devices: *const c_int, n_devices: i32, out_pool: *mut *mut lk_pool) -> lk_status;
    pub fn lk_pool_destroy(pool: *mut lk_pool);
    pub fn lk_pool_device_count(pool: *const lk_pool) -> i32;
    pub fn lk_pool_ctx(pool: *mut lk_pool, i: i32) -> *mut lk_ctx;
    pub fn lk_pool_last_error(pool: *const lk_pool) -> *const c_char;
    pub fn lk_pool_set_tristate_correction(pool: *mut lk_pool, use_tristate: c_int) -> lk_status;
    pub fn lk_pool_pairhmm_batch_host(pool: *mut lk_pool, batch: *const lk_pairhmm_batch, out: *mut f64, flags: *mut u8) -> lk_status;
    pub fn lk_pool_compute_read_likelihoods(
        pool: *mut lk_pool,
        args: *const lk_engine_args,
        batch: *const lk_pairhmm_batch,
        read_mapq: *const u8,
        ref_allele_index: *const i32,
        values: *mut f64,
        keep: *mut u8,
    ) -> lk_status;
    pub fn lk_pool_batcher_create(pool: *mut lk_pool, args: *const lk_engine_args, max_regions: i32, max_wait_us: i32, out: *mut *mut lk_batcher) -> lk_status;
}
