/*
 * lorikeet_b200.h — C ABI of the B200-native read<->haplotype likelihood hot path.
 *
 * This is the drop-in boundary for rhysnewell/Lorikeet's two vector entry points
 * (reference paths relative to the Lorikeet checkout @ 3fbbdec):
 *
 *   gkl::pairhmm::forward()        used at src/pair_hmm/pair_hmm.rs:348-366
 *   gkl::smithwaterman::align()    used at src/smith_waterman/smith_waterman_aligner.rs:56-61
 *
 * plus batch forms of both (one FFI call per active region / per batch of regions),
 * and the engine-level pre/post-ops of
 * src/pair_hmm/pair_hmm_likelihood_calculation_engine.rs:195-242.
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every buffer passed in; the
 *     library keeps no caller pointer after a call returns.
 *   - every entry point returns an lk_status (0 = OK); nothing unwinds or aborts
 *     across the ABI.  lk_last_error(ctx) gives a message for the last failure on ctx.
 *   - an lk_ctx is bound to one CUDA device and is thread-safe (calls serialise on an
 *     internal lock; use one ctx per host thread for concurrency).
 *   - all compute happens on the GPU.  There is no CPU fallback: without a CUDA
 *     device lk_create fails with LK_ERR_CUDA.
 */
#ifndef LORIKEET_B200_H
#define LORIKEET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lk_ctx lk_ctx;

typedef enum lk_status {
    LK_OK = 0,
    LK_ERR_INVALID = 1,   /* bad argument: null pointer, length mismatch, empty sequence ... */
    LK_ERR_CUDA = 2,      /* CUDA runtime failure (message in lk_last_error)                  */
    LK_ERR_NOMEM = 3,     /* host or device allocation failed                                 */
    LK_ERR_CAPACITY = 4   /* output buffer too small (e.g. max_ops for a CIGAR)               */
} lk_status;

/* Replaces gkl::smithwaterman::OverhangStrategy (smith_waterman_aligner.rs:71-72,146-148;
 * shape mirrored at src/smith_waterman/bindings.rs:5-32). */
typedef enum lk_sw_overhang {
    LK_SW_SOFTCLIP = 0,
    LK_SW_INDEL = 1,
    LK_SW_LEADING_INDEL = 2,
    LK_SW_IGNORE = 3
} lk_sw_overhang;

/* Replaces gkl::smithwaterman::Parameters (smith_waterman_aligner.rs:12-25, bindings.rs:34-66). */
typedef struct lk_sw_params {
    int32_t match_value;
    int32_t mismatch_penalty;
    int32_t gap_open_penalty;
    int32_t gap_extend_penalty;
} lk_sw_params;

/* per-pair result flags of the PairHMM entry points */
enum {
    LK_FLAG_RESCUED = 1,   /* fp32 pass under-flowed or was out of its safe domain; value is from the f64 pass */
    LK_FLAG_NEG_INF = 2    /* even f64 under-flowed: value is -inf (the reference returns log10(0) here too)  */
};

/* ---------------------------------------------------------------- context */

/* device < 0 selects the current CUDA device. */
lk_status lk_create(int device, lk_ctx **out_ctx);
void lk_destroy(lk_ctx *ctx);
const char *lk_last_error(const lk_ctx *ctx);
const char *lk_version(void);
/* Mirrors PairHMM::do_not_use_tristate_correction (pair_hmm.rs:189-191): mismatch prior err/1 instead of err/3. */
lk_status lk_set_tristate_correction(lk_ctx *ctx, int use_tristate);
/* 0 (default): fp32 scaled-linear first, f64 only for flagged pairs (the gkl convention);
 * 1: every pair in f64 with the reference's own operation order (slow; for validation). */
lk_status lk_set_force_f64(lk_ctx *ctx, int force_f64);

/* Optional: page-lock memory the caller owns (input arenas, result arrays) so that the host-buffer entry points copy by
 * asynchronous DMA (cudaHostRegister, portable across the devices of a pool).  Pageable memory works everywhere too. */
lk_status lk_host_register(void *ptr, size_t bytes);
lk_status lk_host_unregister(void *ptr);

/* ---------------------------------------------------------------- PairHMM */

/*
 * A batch of active regions.  Region g owns reads [region_read_start[g], region_read_start[g+1])
 * and haplotypes [region_hap_start[g], region_hap_start[g+1]).  Read r occupies bytes
 * [read_off[r], read_off[r+1]) of each of the four read arenas; haplotype h occupies
 * [hap_off[h], hap_off[h+1]) of hap_bases.  read_gcp may be NULL, in which case gcp_constant is
 * the gap-continuation penalty of every base (PairHMMInputScoreImputator::gap_continuation_penalties,
 * pair_hmm_likelihood_calculation_engine.rs:649-651).
 */
typedef struct lk_pairhmm_batch {
    int64_t n_regions;
    const int64_t *region_read_start; /* n_regions + 1 */
    const int64_t *region_hap_start;  /* n_regions + 1 */
    const int64_t *read_off;          /* n_reads + 1   */
    const uint8_t *read_bases;
    const uint8_t *read_qual;
    const uint8_t *read_ins;
    const uint8_t *read_del;
    const uint8_t *read_gcp;          /* per-base, or NULL */
    uint8_t gcp_constant;
    const int64_t *hap_off;           /* n_haps + 1    */
    const uint8_t *hap_bases;
    /* Optional: several SAMPLES per region.  The reference computes one [A x R_s] matrix per sample over the region's
     * shared haplotypes and normalises / filters each on its own (pair_hmm_likelihood_calculation_engine.rs:201-239,
     * 321-340).  Region g owns samples [region_sample_start[g], region_sample_start[g+1]); sample s owns reads
     * [sample_read_start[s], sample_read_start[s+1]); the samples of a region tile its reads in order.  Both NULL: every
     * region is one sample.  Only the region driver (and the batcher in engine mode) looks at them: haplotypes are staged
     * and encoded once per region however many samples share them; lk_pairhmm_batch_host's [r*A + h] layout is unaffected. */
    const int64_t *region_sample_start; /* n_regions + 1, or NULL */
    const int64_t *sample_read_start;   /* n_samples + 1, or NULL */
} lk_pairhmm_batch;

/* Number of (read, haplotype) pairs = sum_g reads_g * haps_g, and sum of R*H over them. */
lk_status lk_pairhmm_batch_dims(const lk_pairhmm_batch *batch, int64_t *n_pairs, int64_t *n_cells);

/*
 * Batch form of PairHMM::compute_likelihoods (pair_hmm.rs:345-375): for every region, for every
 * read r, for every haplotype h: out[base_g + r*A_g + h] = log10 P(read | hap), the layout of
 * m_log_likelihood_array (pair_hmm.rs:262-263,349-369); base_g is the running sum of
 * reads*haps over earlier regions.  The batch pointers are HOST pointers; the call stages, runs and
 * copies back (chunked, overlapped) and returns when out is complete.
 * flags may be NULL; otherwise one LK_FLAG_* byte per pair.
 * out / flags (and values / keep of lk_engine_compute_read_likelihoods) may also point into memory of
 * the context's own device (unified addressing; the copies use cudaMemcpyDefault): a multi-GPU caller
 * that hands the results to a collective keeps them in HBM instead of bouncing them through the host.
 */
lk_status lk_pairhmm_batch_host(lk_ctx *ctx, const lk_pairhmm_batch *batch, double *out, uint8_t *flags);

/*
 * Device-resident variant used for kernel-only timing: lk_pairhmm_upload stages a batch into
 * HBM once; lk_pairhmm_run_resident launches the kernels on the resident copy (no H2D/D2H) and
 * returns after the stream has drained; lk_pairhmm_download copies results back.
 */
typedef struct lk_resident lk_resident;
lk_status lk_pairhmm_upload(lk_ctx *ctx, const lk_pairhmm_batch *batch, lk_resident **out_res);
lk_status lk_pairhmm_run_resident(lk_ctx *ctx, lk_resident *res, float *kernel_ms /* may be NULL */);
lk_status lk_pairhmm_download(lk_ctx *ctx, lk_resident *res, double *out, uint8_t *flags);
void lk_resident_free(lk_ctx *ctx, lk_resident *res);
/* Introspection of a resident batch and of the last host call; what bench.py and a multi-GPU dispatcher need. */
int64_t lk_resident_pairs(const lk_resident *res);            /* (read, haplotype) pairs / SW pairs staged            */
int64_t lk_resident_cells(const lk_resident *res);            /* sum of R*H (PairHMM) or n*m (SW) over them           */
int32_t lk_resident_launch_count(const lk_resident *res);     /* kernels launched by the last run of this batch       */
float lk_resident_last_fast_ms(const lk_resident *res);       /* device time of the dominant kernel(s) in the last
                                                                 TIMED run (kernel_ms != NULL): fp32 wavefront + band */
void *lk_resident_out_device_ptr(const lk_resident *res);     /* device address of the f64 result block [n_pairs], so
                                                                 the results can be handed to a collective in place   */
int64_t lk_resident_sw_packed_pairs(const lk_resident *res);  /* SW pairs of the last run that the packed two-pairs-per-warp
                                                                 kernel took (-1: not an SW batch)                      */
int32_t lk_debug_last_launch_count(const lk_ctx *ctx);        /* kernels launched by the last host-buffer call on ctx */
/* FFMA-only microbenchmark on the context's device for `duration_ms`: the measured fp32 roofline denominator. */
lk_status lk_measure_fp32_fma_tflops(lk_ctx *ctx, double duration_ms, double *tflops);

/*
 * Single-pair convenience wrapper with exactly the argument list of the function pointer
 * returned by gkl::pairhmm::forward() (pair_hmm.rs:358-365,
 * tests/vector_pair_hmm_unit_tests.rs:51-59): (hap, read, qual, ins, del, gcp) -> log10 lk.
 * All five read arrays have read_len entries.
 */
lk_status lk_pairhmm_forward(lk_ctx *ctx, const uint8_t *hap, int64_t hap_len, const uint8_t *read,
                             const uint8_t *qual, const uint8_t *ins, const uint8_t *del, const uint8_t *gcp,
                             int64_t read_len, double *out_log10);

/* ---------------------------------------------------------------- Smith-Waterman */

/*
 * Batch form of SmithWatermanAligner::align (smith_waterman_aligner.rs:47-113) for n_pairs
 * independent (reference, alternate) pairs sharing one parameter set and overhang strategy.
 * Pair p: ref bytes [ref_off[p], ref_off[p+1]) of refs, alt likewise.
 * Results, per pair: n_ops[p] CIGAR operations in forward order in cigar_ops[p*max_ops ..],
 * BAM-encoded (len << 4 | op; M=0 I=1 D=2 S=4), and alignment_offset[p]
 * (SmithWatermanAlignmentResult, smith_waterman_aligner.rs:455-458).
 * n_ops[p] = -1: empty sequence (the reference asserts, :65-68); -2: more than max_ops operations.
 */
lk_status lk_sw_batch_host(lk_ctx *ctx, int64_t n_pairs, const int64_t *ref_off, const uint8_t *refs,
                           const int64_t *alt_off, const uint8_t *alts, const lk_sw_params *params,
                           lk_sw_overhang strategy, int32_t max_ops, uint32_t *cigar_ops, int32_t *n_ops,
                           int32_t *alignment_offset);

/* Resident variant for kernel-only timing (same contract as the PairHMM one). */
lk_status lk_sw_upload(lk_ctx *ctx, int64_t n_pairs, const int64_t *ref_off, const uint8_t *refs,
                       const int64_t *alt_off, const uint8_t *alts, const lk_sw_params *params,
                       lk_sw_overhang strategy, int32_t max_ops, lk_resident **out_res);
lk_status lk_sw_run_resident(lk_ctx *ctx, lk_resident *res, float *kernel_ms);
lk_status lk_sw_download(lk_ctx *ctx, lk_resident *res, uint32_t *cigar_ops, int32_t *n_ops,
                         int32_t *alignment_offset);

/*
 * Single-pair wrapper with the argument list of the function pointer returned by
 * gkl::smithwaterman::align() (smith_waterman_aligner.rs:56-61): the CIGAR is returned as ASCII
 * text ("31M20S") in cigar_text (NUL-terminated, capacity cigar_cap) exactly as the reference
 * parses it with CigarString::try_from, plus the alignment offset.
 */
lk_status lk_sw_align(lk_ctx *ctx, const uint8_t *ref, int64_t ref_len, const uint8_t *alt, int64_t alt_len,
                      const lk_sw_params *params, lk_sw_overhang strategy, char *cigar_text, size_t cigar_cap,
                      int32_t *alignment_offset);

/* ---------------------------------------------------------------- haplotype -> reference CIGARs */

/*
 * Batch form of CigarUtils::calculate_cigar (src/reads/cigar_utils.rs:358-457): per pair the trivial cases
 * (empty alternate; equal lengths with <= 2 mismatches) are resolved on the host, every other pair is wrapped in
 * 10 x 'N' (SW_PAD, :10,387-398) and aligned in ONE Smith-Waterman batch on the GPU, then trimmed
 * (AlignmentUtils::trim_cigar_by_bases), left-aligned (left_align_indels) and given back its leading / trailing
 * deletions exactly as the reference does.  Results as in lk_sw_batch_host; n_ops[p] = -3 when the reference
 * returns None (is_s_w_failure, :469-487), -4 where the reference would panic, -2 on capacity.
 */
lk_status lk_calculate_cigar_batch(lk_ctx *ctx, int64_t n_pairs, const int64_t *ref_off, const uint8_t *refs,
                                   const int64_t *alt_off, const uint8_t *alts, const lk_sw_params *params,
                                   lk_sw_overhang strategy, int32_t max_ops, uint32_t *cigar_ops, int32_t *n_ops);
/* The two host-side halves of the above, usable on their own: returns 1 if the pair needs no alignment. */
int lk_calculate_cigar_shortcut(const uint8_t *ref, int64_t ref_len, const uint8_t *alt, int64_t alt_len, int32_t max_ops,
                                uint32_t *out_ops, int32_t *out_n_ops);
/* From the SW result (BAM-encoded ops + offset) on the 10 x 'N' padded pair to the final CIGAR. */
lk_status lk_calculate_cigar_finish(const uint8_t *ref, int64_t ref_len, const uint8_t *alt, int64_t alt_len,
                                    const uint32_t *sw_ops, int32_t sw_n_ops, int32_t sw_offset, int32_t max_ops,
                                    uint32_t *out_ops, int32_t *out_n_ops);

/* ---------------------------------------------------------------- dangling-end merges of the read-threading graph (N4) */

/*
 * The alignment step of ReadThreadingGraph's dangling-tail / dangling-head recovery
 * (src/read_threading/read_threading_graph.rs:1370-1486; align calls at :1408-1414 and :1471-1477): each dangling path
 * (alts) is aligned to its reference path (refs) with OverhangStrategy::LeadingInDel and the dangling-end parameters the
 * caller passes (STANDARD_NGS by default), then AlignmentUtils::remove_trailing_deletions (alignment_utils.rs:702-707)
 * drops a final deletion.  Only the CIGAR is used by the callers.  All the dangling ends of a region's graph (or of many
 * regions) go in ONE Smith-Waterman launch.  Results as in lk_sw_batch_host (n_ops -1 empty sequence, -2 capacity).
 */
lk_status lk_dangling_end_sw_batch(lk_ctx *ctx, int64_t n_pairs, const int64_t *ref_off, const uint8_t *refs,
                                   const int64_t *alt_off, const uint8_t *alts, const lk_sw_params *params, int32_t max_ops,
                                   uint32_t *cigar_ops, int32_t *n_ops);

/* ---------------------------------------------------------------- read -> best-haplotype realignment (N2) */

/*
 * Batch form of AssemblyBasedCallerUtils::realign_reads_to_their_best_haplotype
 * (src/assembly/assembly_based_caller_utils.rs:208-246); regions may hold several samples (see the struct).
 * `values` may be HOST memory or memory of the context's device (e.g. the buffer lk_engine_compute_read_likelihoods
 * just filled): in the second case the best-allele search runs as a kernel on the matrices where they are and only one
 * int32 + one flag per read come back to the host.
 * Per read: the best haplotype with the reference's tie-breaking (AlleleLikelihoods::search_best_allele with
 * priorities, src/model/allele_likelihoods.rs:457-554; priority = is_ref + 1 - cigar elements,
 * assembly_based_caller_utils.rs:187-195), then AlignmentUtils::create_read_aligned_to_ref
 * (src/reads/alignment_utils.rs:40-164): ONE Smith-Waterman batch on the GPU (haplotype vs read, SoftClip,
 * ALIGNMENT_TO_BEST_HAPLOTYPE_SW_PARAMETERS), then the read->haplotype CIGAR is mapped through the
 * haplotype->reference CIGAR (apply_cigar_to_cigar), left-aligned and given back the read's clips.
 */
typedef struct lk_realign_batch {
    int64_t n_regions;
    const int64_t *region_read_start;  /* [n_regions + 1] */
    const int64_t *region_hap_start;   /* [n_regions + 1] */
    const double *values;              /* per region an [A_g x R_g] row-major block, consecutive (the layout
                                          lk_engine_compute_read_likelihoods returns) */
    const int64_t *region_n_evidence;  /* reads of region g that are still evidence (<= R_g); NULL = all */
    const int32_t *ref_allele_index;   /* index of the reference haplotype in region g; NULL = 0 */
    const int64_t *reference_start;    /* per region: padded_reference_loc.get_start() */
    /* haplotypes */
    const int64_t *hap_off;            /* [n_haps + 1] */
    const uint8_t *hap_bases;
    const int64_t *hap_cigar_off;      /* [n_haps + 1] into hap_cigars */
    const uint32_t *hap_cigars;        /* BAM-encoded (len << 4 | op) haplotype -> reference CIGARs */
    const int64_t *hap_start_wrt_ref;  /* Haplotype::alignment_start_hap_wrt_ref */
    /* reads */
    const int64_t *read_off;           /* [n_reads + 1] */
    const uint8_t *read_bases;         /* WITHOUT soft clips (ReadClipper::hard_clip_soft_clipped_bases is the caller's) */
    const int64_t *read_cigar_off;     /* [n_reads + 1] into read_cigars */
    const uint32_t *read_cigars;       /* the reads' original CIGARs (their clips are re-attached) */
    const int64_t *read_orig_len;      /* original read lengths (with soft clips) */
    double informative_threshold;      /* 0 = LOG_10_INFORMATIVE_THRESHOLD (0.2) */
    /* Optional: several samples per region, as in lk_pairhmm_batch: `values` then holds per region one [A_g x R_s] block
     * per sample, consecutive (what lk_engine_compute_read_likelihoods returns for such a batch), and sample_n_evidence
     * (may be NULL = all) counts the reads of each SAMPLE that are still evidence; region_n_evidence is then ignored. */
    const int64_t *region_sample_start; /* [n_regions + 1], or NULL */
    const int64_t *sample_read_start;   /* [n_samples + 1], or NULL */
    const int64_t *sample_n_evidence;   /* [n_samples], or NULL */
} lk_realign_batch;

/*
 * Outputs per read: best_allele (index inside its region, -1 = none), informative (BestAllele::is_informative; may be
 * NULL), the new CIGAR (max_ops elements per read) with n_ops >= 0, or n_ops = -1 read not realigned, -2 capacity,
 * -4 where the reference would panic; pos = new alignment start.
 */
lk_status lk_realign_reads_batch(lk_ctx *ctx, const lk_realign_batch *batch, int32_t max_ops, int32_t *best_allele,
                                 uint8_t *informative, uint32_t *cigar_ops, int32_t *n_ops, int64_t *pos);

/* Host-side pieces of the above, usable (and tested) on their own. */
lk_status lk_best_alleles_breaking_ties(const double *values, int64_t n_alleles, int64_t n_reads, int64_t n_evidence,
                                        const int32_t *priorities /* may be NULL */, double informative_threshold,
                                        int32_t *best_allele, double *confidence /* may be NULL */,
                                        uint8_t *informative /* may be NULL */);
lk_status lk_realign_finish(const uint8_t *read, int64_t read_len, const uint32_t *orig_cigar, int32_t orig_n,
                            int64_t orig_read_len, const uint32_t *sw_ops, int32_t sw_n_ops, int32_t sw_offset,
                            const uint32_t *hap_cigar, int32_t hap_cigar_n, int64_t hap_start_wrt_ref,
                            const uint8_t *ref_hap, int64_t ref_hap_len, int64_t reference_start, int32_t max_ops,
                            uint32_t *out_ops, int32_t *out_n_ops, int64_t *out_pos);
/* AlignmentUtils::apply_cigar_to_cigar (alignment_utils.rs:240-281); out_n_ops = -4 where the reference panics */
lk_status lk_cigar_apply_to_cigar(const uint32_t *first_to_second, int32_t n12, const uint32_t *second_to_third, int32_t n23,
                                  int32_t max_ops, uint32_t *out_ops, int32_t *out_n_ops);
/* AlignmentUtils::append_clipped_elements_from_cigar_to_cigar (:173-205) */
lk_status lk_cigar_append_clipped_elements(const uint32_t *cigar, int32_t n, const uint32_t *original, int32_t n_orig,
                                           int32_t max_ops, uint32_t *out_ops, int32_t *out_n_ops);
/* AlignmentUtils::read_start_on_reference_haplotype (:283-312); -1 where the reference panics */
int64_t lk_read_start_on_reference_haplotype(const uint32_t *hap_vs_ref_cigar, int32_t n, uint32_t read_start_on_haplotype);

/* ---------------------------------------------------------------- engine pre/post-ops (host side) */

/* PCRErrorModel (pair_hmm_likelihood_calculation_engine.rs:60-70) */
typedef enum lk_pcr_model { LK_PCR_NONE = 0, LK_PCR_HOSTILE = 1, LK_PCR_AGGRESSIVE = 2, LK_PCR_CONSERVATIVE = 3 } lk_pcr_model;

/* Arguments of PairHMMLikelihoodCalculationEngine::new (:129-167), in the same order. */
typedef struct lk_engine_args {
    uint8_t constant_gcp;
    double log10_global_read_mismapping_rate;
    int32_t pcr_error_model;                 /* lk_pcr_model */
    uint8_t base_quality_score_threshold;
    int32_t dynamic_read_disqualification;
    double read_disqualification_scale;
    double expected_error_rate_per_base;
    int32_t symmetrically_normalize_alleles_to_reference;
    int32_t disable_cap_read_qualities_to_mapq;
    int32_t modify_soft_clipped_bases;       /* accepted for signature parity; clipping itself is the caller's */
} lk_engine_args;

/*
 * modify_read_qualities (:352-426) on packed arenas, in place: apply_pcr_error_model (:502-526),
 * cap_minimum_read_qualities (:428-458).  read_mapq has one entry per read.  Host-side, threaded.
 */
lk_status lk_engine_modify_read_qualities(const lk_engine_args *args, int64_t n_reads, const int64_t *read_off,
                                          const uint8_t *read_bases, uint8_t *read_qual, uint8_t *read_ins,
                                          uint8_t *read_del, const uint8_t *read_mapq, int n_threads);

/* VariantContextUtils::find_number_of_repetitions (src/model/variant_context_utils.rs:240-337), the repeat counter under
 * the PCR error model: how many times repeat_unit occurs back to back at the start (leading_repeats != 0) or at the end of
 * test_string.  Host utility; -1 on bad arguments. */
int64_t lk_find_number_of_repetitions(const uint8_t *repeat_unit, int64_t unit_len, const uint8_t *test_string,
                                      int64_t test_len, int leading_repeats);

/* The same pre-ops as one device launch (the kernel the region driver runs on its staged copies): upload, launch,
 * download, in place on the caller's arrays.  Byte-identical to lk_engine_modify_read_qualities. */
lk_status lk_engine_modify_read_qualities_device(lk_ctx *ctx, const lk_engine_args *args, int64_t n_reads,
                                                 const int64_t *read_off, const uint8_t *read_bases, uint8_t *read_qual,
                                                 uint8_t *read_ins, uint8_t *read_del, const uint8_t *read_mapq);

/*
 * Region driver = compute_read_likelihoods (:195-242) for a batch of regions (one sample each, or several: see
 * lk_pairhmm_batch.region_sample_start; the [A_g x R_s] blocks of a region's samples then follow one another):
 * modify_read_qualities on copies -> PairHMM on the GPU -> transpose into the reference's
 * [allele][read] matrices -> normalize_likelihoods (allele_likelihoods.rs:378-444) ->
 * filter_poorly_modeled_evidence (:925-1018, thresholds :244-319 computed from the ORIGINAL
 * qualities and lengths, reproducing the "HMMQuals" key quirk).
 * values: per region an [A_g x R_g] row-major f64 block at the same base offsets as
 * lk_pairhmm_batch_host; removed reads' columns are compacted away and the tail set to NaN as the
 * reference does.  keep[r] (per read, may be NULL) = 1 if the read survived the filter.
 * ref_allele_index[g] (may be NULL = 0 for all) is the index of the reference haplotype in region g.
 * Reads of a region without haplotypes are all filtered (keep = 0; the maximum over no alleles is -inf).
 * Known divergence (non-default path): with modify_soft_clipped_bases == 0 the reference computes the disqualification
 * thresholds from the ORIGINAL evidence read, soft-clipped bases included (read.len(), read.qual(); the "HMMQuals" key
 * quirk); this entry point sees the reads as passed in (hard-clipped by the caller), so for soft-clipped reads
 * ceil(len * rate) and the dynamic mean / variance sums use the clipped length and qualities.
 */
lk_status lk_engine_compute_read_likelihoods(lk_ctx *ctx, const lk_engine_args *args, const lk_pairhmm_batch *batch,
                                             const uint8_t *read_mapq, const int32_t *ref_allele_index,
                                             double *values, uint8_t *keep);

/*
 * The post-ops of the region driver on their own: AlleleLikelihoods::normalize_likelihoods (allele_likelihoods.rs:378-444)
 * and filter_poorly_modeled_evidence (:925-1018, thresholds ...engine.rs:244-319) as the device epilogue, applied to
 * likelihoods the caller supplies in lk_pairhmm_batch_host's layout (per region [r*A + h]); values / keep as in
 * lk_engine_compute_read_likelihoods.  Of the batch only the region / sample structure, read_off and read_qual (the
 * ORIGINAL qualities the thresholds are computed from) are read.
 */
lk_status lk_engine_normalize_and_filter_device(lk_ctx *ctx, const lk_engine_args *args, const lk_pairhmm_batch *batch,
                                                const int32_t *ref_allele_index, const double *likelihoods, double *values,
                                                uint8_t *keep);

/* ---------------------------------------------------------------- cross-thread batching front-end (N4) */

/*
 * Lorikeet calls the engine from Rayon workers, one region per task (src/assembly/assembly_region_walker.rs:140-277).
 * A batcher keeps those call sites unchanged: lk_batcher_submit blocks the calling thread, a dispatcher thread
 * concatenates the regions pending from all threads (at most max_regions, waiting at most max_wait_us for more) and
 * runs ONE launch for them.  args != NULL: the region driver (lk_engine_compute_read_likelihoods; values = [A x R]
 * block, keep_or_flags = keep mask per read); args == NULL: lk_pairhmm_batch_host ([r*A + h], flags per pair).
 * `region` must describe exactly one region (n_regions == 1).
 */
typedef struct lk_batcher lk_batcher;
lk_status lk_batcher_create(lk_ctx *ctx, const lk_engine_args *args /* may be NULL */, int32_t max_regions,
                            int32_t max_wait_us, lk_batcher **out);
void lk_batcher_destroy(lk_batcher *b);
lk_status lk_batcher_submit(lk_batcher *b, const lk_pairhmm_batch *region, const uint8_t *read_mapq,
                            int32_t ref_allele_index, double *values, uint8_t *keep_or_flags /* may be NULL */);
void lk_batcher_stats(lk_batcher *b, int64_t *n_batches, int64_t *n_regions);
/*
 * Non-blocking form: queue a region and return at once with a ticket; lk_batcher_wait blocks until that region's
 * results are in `values` / `keep_or_flags`, returns its status and releases the ticket.  Lets one worker keep several
 * regions in flight (submit the likelihoods of region i+1.. while genotyping region i), which is what fills the GPU when
 * there are fewer worker threads than ~100.  The region's arrays and the output buffers must stay valid until the wait
 * returns; every ticket must be waited on exactly once, before lk_batcher_destroy.
 */
typedef struct lk_ticket lk_ticket;
lk_status lk_batcher_submit_async(lk_batcher *b, const lk_pairhmm_batch *region, const uint8_t *read_mapq,
                                  int32_t ref_allele_index, double *values, uint8_t *keep_or_flags /* may be NULL */,
                                  lk_ticket **out_ticket);
lk_status lk_batcher_wait(lk_batcher *b, lk_ticket *ticket);

/* ---------------------------------------------------------------- several devices in one process */

/*
 * Lorikeet is ONE process whose Rayon workers call the engine (src/assembly/assembly_region_walker.rs:140-277,
 * nested under src/haplotype/haplotype_caller_engine.rs:949-951).  A pool puts several GPUs behind one handle: one
 * lk_ctx (three streams, chunk pipeline) and its dispatcher threads per device; regions are independent, so they are
 * dealt to the devices with no data-path exchange and no collective.  devices == NULL (or n_devices <= 0): every
 * visible device.  The batch entry points cut the batch into contiguous chunks of about equal CELL counts and deal them
 * dynamically (one host thread per device pulls the next chunk when its device is free); results land in the caller's
 * arrays exactly as from the single-device calls (bit-identical: a region's result does not depend on its batch).
 */
typedef struct lk_pool lk_pool;
lk_status lk_pool_create(const int *devices, int32_t n_devices, lk_pool **out_pool);
void lk_pool_destroy(lk_pool *pool);
int32_t lk_pool_device_count(const lk_pool *pool);
lk_ctx *lk_pool_ctx(lk_pool *pool, int32_t i);               /* the i-th device's context (owned by the pool) */
const char *lk_pool_last_error(const lk_pool *pool);
lk_status lk_pool_set_tristate_correction(lk_pool *pool, int use_tristate);
lk_status lk_pool_pairhmm_batch_host(lk_pool *pool, const lk_pairhmm_batch *batch, double *out, uint8_t *flags);
lk_status lk_pool_compute_read_likelihoods(lk_pool *pool, const lk_engine_args *args, const lk_pairhmm_batch *batch,
                                           const uint8_t *read_mapq, const int32_t *ref_allele_index, double *values,
                                           uint8_t *keep);
/* The cross-thread batcher over every device of the pool: lk_batcher_submit / _submit_async / _wait / _stats / _destroy
 * as above; the dispatcher threads of all devices pull from one queue, so a free device takes the next share. */
lk_status lk_pool_batcher_create(lk_pool *pool, const lk_engine_args *args /* may be NULL */, int32_t max_regions,
                                 int32_t max_wait_us, lk_batcher **out);

#ifdef __cplusplus
}
#endif
#endif /* LORIKEET_B200_H */
