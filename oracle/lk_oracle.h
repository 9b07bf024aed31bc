/*
 * lk_oracle.h — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the reference's in-tree scalar algorithms for the
 * read<->haplotype likelihood hot path (PairHMM forward + Smith-Waterman with
 * CIGAR traceback + the engine's pre/post-ops).  Every function cites the
 * reference file:line it follows (paths relative to the reference checkout,
 * rhysnewell/Lorikeet @ 3fbbdec).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may link or call this library.  The product path
 * (lorikeet_b200/) never does.
 *
 * PARITY PINNING: the reference itself is Rust and cannot be compiled in the
 * build container (no cargo/rustc); its "vector" kernels live in the
 * un-vendored crate gkl ^0.1.1.  This restatement is pinned against the
 * reference's own fixtures instead (tests/golden/, see tests/test_oracle_*.py):
 *   - 104 golden rows of tests/resources/pairhmm-testdata.txt (1e-5 abs),
 *   - the SW known-answer (offset, CIGAR) cases of
 *     tests/smith_waterman_aligner_unit_tests.rs:229-400,
 *   - analytic PairHMM families of tests/pair_hmm_unit_tests.rs.
 */
#ifndef LK_ORACLE_H
#define LK_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- quality / math helpers (utils/quality_utils.rs, utils/math_utils.rs) ---- */
double lko_qual_to_error_prob(uint8_t qual);                 /* quality_utils.rs:98-104  */
double lko_qual_to_prob(uint8_t qual);                       /* quality_utils.rs:82-88   */
double lko_approximate_log10_sum_log10(double a, double b);  /* math_utils.rs:314-334    */

/* ---- PairHMMModel (pair_hmm/pair_hmm_model.rs) ---- */
double lko_match_to_match_prob(int ins_qual, int del_qual);  /* pair_hmm_model.rs:442-461 */
/* dest[0..5] = MM, GM(indel->match), MI, II, MD, DD          pair_hmm_model.rs:142-156  */
void lko_qual_to_trans_probs(double dest[6], uint8_t ins_qual, uint8_t del_qual, uint8_t gcp);

/* ---- PairHMM scalar forward (pair_hmm/pair_hmm.rs:503-615, 626-654) ----
 * Returns log10 P(read | haplotype).  use_tristate != 0 => mismatch prior is
 * err/3 (production, gkl); 0 => err/1 (do_not_use_tristate_correction). */
double lko_pairhmm_forward(const uint8_t *hap, int hap_len,
                           const uint8_t *read, const uint8_t *qual,
                           const uint8_t *ins, const uint8_t *del,
                           const uint8_t *gcp, int read_len, int use_tristate);

/* Same recurrence, but every read row is renormalised by a power of two with
 * the exponent carried in an integer, so it does not underflow for long reads
 * (documented deviation: the reference's un-rescaled f64 returns -inf once a
 * path probability drops below ~1e-615). */
double lko_pairhmm_forward_rescaled(const uint8_t *hap, int hap_len,
                                    const uint8_t *read, const uint8_t *qual,
                                    const uint8_t *ins, const uint8_t *del,
                                    const uint8_t *gcp, int read_len, int use_tristate);

/* pair_hmm.rs:706-717 */
int lko_find_first_position_where_haplotypes_differ(const uint8_t *h1, int n1,
                                                    const uint8_t *h2, int n2);

/* Region driver, AVX-arm layout (pair_hmm.rs:345-375): out[r*A + h].
 * Offsets are CSR-style byte offsets (n+1 entries). */
void lko_pairhmm_region(int n_reads, const int64_t *read_off,
                        const uint8_t *read_bases, const uint8_t *read_qual,
                        const uint8_t *read_ins, const uint8_t *read_del,
                        const uint8_t *read_gcp,
                        int n_haps, const int64_t *hap_off, const uint8_t *hap_bases,
                        int use_tristate, int rescaled, double *out);

/* ---- Smith-Waterman (smith_waterman/smith_waterman_aligner.rs) ---- */
enum { LKO_SW_SOFTCLIP = 0, LKO_SW_INDEL = 1, LKO_SW_LEADING_INDEL = 2, LKO_SW_IGNORE = 3 };
/* BAM op codes used in cigar_ops: (len << 4) | op */
enum { LKO_CIGAR_M = 0, LKO_CIGAR_I = 1, LKO_CIGAR_D = 2, LKO_CIGAR_S = 4 };

/* smith_waterman_aligner.rs:47-113 (AVXMode::None arm): substring shortcut for
 * SoftClip/Ignore, else calculate_matrix (:124-263) + calculate_cigar (:273-442).
 * params = {match, mismatch, gap_open, gap_extend}.
 * Returns number of cigar ops written (forward order), or -1 on empty input,
 * -2 if max_ops is too small. */
int lko_sw_align(const uint8_t *ref, int ref_len, const uint8_t *alt, int alt_len,
                 const int32_t params[4], int strategy,
                 uint32_t *cigar_ops, int max_ops, int32_t *alignment_offset);
/* Same without the substring shortcut (what a pure matrix implementation gives). */
int lko_sw_align_full(const uint8_t *ref, int ref_len, const uint8_t *alt, int alt_len,
                      const int32_t params[4], int strategy,
                      uint32_t *cigar_ops, int max_ops, int32_t *alignment_offset);
/* alignment_utils.rs:717-735; returns -1 if not found */
int lko_last_index_of(const uint8_t *ref, int ref_len, const uint8_t *query, int query_len);

/* ---- engine pre-ops (pair_hmm_likelihood_calculation_engine.rs) ---- */
/* :186-193 */
uint8_t lko_pcr_error_model_adjusted_qual(int repeat_length, double rate_factor);
/* :528-611, returns the repeat length (second tuple element) */
int lko_find_number_of_repetitions(const uint8_t *unit, int unit_len, const uint8_t *test, int test_len, int leading);
int lko_find_tandem_repeat_length(const uint8_t *bases, int len, int offset);
/* :502-526; pcr_model = 0 None, 1 Hostile, 2 Aggressive, 3 Conservative */
void lko_apply_pcr_error_model(const uint8_t *bases, int len, uint8_t *ins, uint8_t *del, int pcr_model);
/* :428-466 */
void lko_cap_minimum_read_qualities(uint8_t *qual, uint8_t *ins, uint8_t *del, int len,
                                    uint8_t mapq, uint8_t base_quality_score_threshold,
                                    int disable_cap_to_mapq);

/* ---- engine post-ops ---- */
/* allele_likelihoods.rs:378-444 + search_best_allele :457-554 (priorities = None).
 * values is [A x R] row-major ([allele][read]).  ref_allele_index < 0 = none. */
void lko_normalize_likelihoods(double *values, int n_alleles, int n_reads,
                               double max_likelihood_difference_cap,
                               int symmetrically_normalize_alleles_to_reference,
                               int ref_allele_index);
/* pair_hmm_likelihood_calculation_engine.rs:261-291 */
double lko_log10_dynamic_read_qual_threshold(const uint8_t *qual, int len, double dynamic_constant);
/* :293-319 */
double lko_log10_min_true_likelihood(int read_len, double max_error_per_base, int cap_likelihoods);
/* allele_likelihoods.rs:925-1018: keep[r]=0 for removed reads; matrix columns
 * compacted in place, tail set to NaN.  thresholds[r] is the per-read
 * log10_min_true_likelihood closure value.  Returns new read count. */
int lko_filter_poorly_modeled_evidence(double *values, int n_alleles, int n_reads,
                                       const double *thresholds, uint8_t *keep);

/* ---- multi-threaded batch drivers (lk_oracle_mt.c) ---- */
/* lk_oracle_simd.c: OUR fp32 SIMD stand-in for the un-vendored gkl AVX path (baseline infrastructure, not gkl) */
void lko_simd_force_lanes(int lanes);
int lko_simd_supported(int lanes);
const char *lko_simd_isa(void);
double lko_pairhmm_forward_vector(const uint8_t *hap, int hap_len, const uint8_t *read, const uint8_t *qual, const uint8_t *ins,
                                  const uint8_t *del, const uint8_t *gcp, int read_len, int use_tristate, int *redone);
int lko_pairhmm_batch_vector_mt(int n_regions, const int64_t *region_read_start, const int64_t *region_hap_start,
                                const int64_t *read_off, const uint8_t *read_bases, const uint8_t *read_qual,
                                const uint8_t *read_ins, const uint8_t *read_del, const uint8_t *read_gcp,
                                const int64_t *hap_off, const uint8_t *hap_bases, const int64_t *out_off,
                                int use_tristate, int n_threads, double *out);
void lko_init(void); /* force lazy tables; call once before threads */
/* One region per worker task (assembly_region_walker.rs:140).  Region g owns reads
 * [region_read_start[g], region_read_start[g+1]) and haplotypes likewise; its
 * [reads x haps] block starts at out[out_off[g]] laid out r*A + h. */
void lko_pairhmm_batch_mt(int n_regions, const int64_t *region_read_start, const int64_t *region_hap_start,
                          const int64_t *read_off, const uint8_t *read_bases, const uint8_t *read_qual,
                          const uint8_t *read_ins, const uint8_t *read_del, const uint8_t *read_gcp,
                          const int64_t *hap_off, const uint8_t *hap_bases, const int64_t *out_off,
                          int use_tristate, int rescaled, int n_threads, double *out);
void lko_sw_batch_mt(int n_pairs, const int64_t *ref_off, const uint8_t *refs, const int64_t *alt_off,
                     const uint8_t *alts, const int32_t params[4], int strategy, int use_shortcut,
                     int max_ops, int n_threads, uint32_t *cigar_ops, int32_t *n_ops, int32_t *offsets);

#ifdef __cplusplus
}
#endif
#endif
