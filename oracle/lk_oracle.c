/*
 * lk_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See lk_oracle.h for scope and parity-pinning notes.  Plain C11, no deps.
 * Build: make -C oracle   (gcc -O3 -ffp-contract=off: no FMA contraction, so
 * the f64 arithmetic rounds like the reference's un-fused Rust expressions).
 */
#include "lk_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------ */
/* utils/quality_utils.rs                                                    */
/* ------------------------------------------------------------------------ */

#define LKO_MAX_QUAL 254          /* quality_utils.rs:24 */
#define LKO_MIN_USABLE_Q_SCORE 6  /* quality_utils.rs:23 */

/* quality_utils.rs:98-104: 10^(qual / -10) */
double lko_qual_to_error_prob(uint8_t qual) { return pow(10.0, ((double)qual) / -10.0); }

/* quality_utils.rs:82-88 */
double lko_qual_to_prob(uint8_t qual) { return 1.0 - lko_qual_to_error_prob(qual); }

/* ------------------------------------------------------------------------ */
/* utils/math_utils.rs: JacobianLogTable (:484-498), cache (:8-14),          */
/* approximate_log10_sum_log10 (:314-334)                                    */
/* ------------------------------------------------------------------------ */

#define JAC_MAX_TOLERANCE 8.0
#define JAC_TABLE_STEP 0.0001
#define JAC_INV_STEP (1.0 / JAC_TABLE_STEP)

static double *jac_cache = NULL;
static int jac_cache_len = 0;

static void jac_init(void) {
    if (jac_cache) return;
    int n = (int)((JAC_MAX_TOLERANCE / JAC_TABLE_STEP) + 1.0);
    double *c = (double *)malloc(sizeof(double) * (size_t)n);
    for (int k = 0; k < n; k++) c[k] = log10(1.0 + pow(10.0, -((double)k) * JAC_TABLE_STEP));
    jac_cache_len = n;
    jac_cache = c;
}

static double jac_get(double difference) {
    /* Rust f64::round = half away from zero = C round() */
    size_t index = (size_t)round(difference * JAC_INV_STEP);
    return jac_cache[index];
}

double lko_approximate_log10_sum_log10(double a, double b) {
    jac_init();
    if (a > b) return lko_approximate_log10_sum_log10(b, a);
    if (a == -INFINITY) return b;
    double diff = b - a;
    return b + (diff < JAC_MAX_TOLERANCE ? jac_get(diff) : 0.0);
}

/* ------------------------------------------------------------------------ */
/* pair_hmm/pair_hmm_model.rs                                                */
/* ------------------------------------------------------------------------ */

static double *m2m_prob = NULL; /* triangular, pair_hmm_model.rs:34,47-78 */

static void model_init(void) {
    if (m2m_prob) return;
    jac_init();
    const double LN10 = log(10.0);
    const double INV_LN10 = 1.0 / LN10;
    size_t n = ((size_t)(LKO_MAX_QUAL + 1) * (size_t)(LKO_MAX_QUAL + 2)) >> 1;
    double *t = (double *)malloc(sizeof(double) * n);
    size_t offset = 0;
    for (int i = 0; i <= LKO_MAX_QUAL; i++) {
        for (int j = 0; j <= i; j++) {
            double log10_sum = lko_approximate_log10_sum_log10(-0.1 * (double)i, -0.1 * (double)j);
            double p = pow(10.0, log10_sum);
            double l10 = log1p(-(p < 1.0 ? p : 1.0)) * INV_LN10;
            t[offset + (size_t)j] = pow(10.0, l10);
        }
        offset += (size_t)i + 1;
    }
    m2m_prob = t;
}

double lko_match_to_match_prob(int ins_qual, int del_qual) {
    model_init();
    int min_qual, max_qual;
    if (ins_qual <= del_qual) { min_qual = ins_qual; max_qual = del_qual; }
    else { min_qual = del_qual; max_qual = ins_qual; }
    if (LKO_MAX_QUAL < max_qual) {
        return 1.0 - pow(10.0, lko_approximate_log10_sum_log10(-0.1 * (double)min_qual, -0.1 * (double)max_qual));
    }
    return m2m_prob[(((size_t)max_qual * ((size_t)max_qual + 1)) >> 1) + (size_t)min_qual];
}

enum { T_MM = 0, T_GM = 1, T_MI = 2, T_II = 3, T_MD = 4, T_DD = 5 }; /* pair_hmm_model.rs:88-113 */

void lko_qual_to_trans_probs(double dest[6], uint8_t ins_qual, uint8_t del_qual, uint8_t gcp) {
    dest[T_MM] = lko_match_to_match_prob(ins_qual, del_qual);
    dest[T_MI] = lko_qual_to_error_prob(ins_qual);
    dest[T_MD] = lko_qual_to_error_prob(del_qual);
    dest[T_GM] = lko_qual_to_prob(gcp);
    double tmp = lko_qual_to_error_prob(gcp);
    dest[T_II] = tmp;
    dest[T_DD] = tmp;
}

/* ------------------------------------------------------------------------ */
/* pair_hmm/pair_hmm.rs — scalar forward                                     */
/* ------------------------------------------------------------------------ */

/* per-quality tables so the inner loop does no pow(); values are bit-identical
 * to calling the scalar helpers above. */
static double q_err[256], q_ok[256];
static int q_tables_ready = 0;
static void q_tables_init(void) {
    if (q_tables_ready) return;
    for (int q = 0; q < 256; q++) { q_err[q] = lko_qual_to_error_prob((uint8_t)q); q_ok[q] = 1.0 - q_err[q]; }
    q_tables_ready = 1;
}

static double forward_impl(const uint8_t *hap, int H, const uint8_t *read, const uint8_t *qual,
                           const uint8_t *ins, const uint8_t *del, const uint8_t *gcp, int R,
                           int use_tristate, int rescale) {
    model_init();
    q_tables_init();
    /* pair_hmm.rs:16-17 */
    const double INITIAL_CONDITION = ldexp(1.0, 1020);
    const double INITIAL_CONDITION_LOG10 = log10(INITIAL_CONDITION);
    const double tri = use_tristate ? 3.0 : 1.0; /* pair_hmm.rs:53,647-651 */

    size_t W = (size_t)H + 1;
    /* two rolling rows of each of M, I, D: the reference keeps whole matrices
     * (pair_hmm.rs:144-148) but only rows i-1 and i are ever read. */
    double *buf = (double *)calloc(6 * W, sizeof(double));
    double *Mp = buf, *Ip = buf + W, *Dp = buf + 2 * W;
    double *Mc = buf + 3 * W, *Ic = buf + 4 * W, *Dc = buf + 5 * W;

    /* pair_hmm.rs:515-529: row 0 of the deletion matrix (all columns incl. 0) */
    double initial_value = INITIAL_CONDITION / (double)H;
    for (size_t j = 0; j < W; j++) Dp[j] = initial_value;

    long shift = 0; /* power-of-two exponent removed so far (rescaled variant only) */
    double t[6];
    for (int i = 1; i <= R; i++) {
        lko_qual_to_trans_probs(t, ins[i - 1], del[i - 1], gcp[i - 1]); /* pair_hmm.rs:682-694 */
        uint8_t x = read[i - 1];
        double p_match = q_ok[qual[i - 1]];            /* pair_hmm.rs:644 */
        double p_mis = q_err[qual[i - 1]] / tri;       /* pair_hmm.rs:646-651 */
        Mc[0] = 0.0; Ic[0] = 0.0; Dc[0] = 0.0;          /* column 0 of rows >= 1 stays zero */
        double rowmax = 0.0;
        for (int j = 1; j <= H; j++) {
            uint8_t y = hap[j - 1];
            double prior = (x == y || x == 'N' || y == 'N') ? p_match : p_mis;
            /* pair_hmm.rs:575-591, same association order */
            double m = prior * (Mp[j - 1] * t[T_MM] + Ip[j - 1] * t[T_GM] + Dp[j - 1] * t[T_GM]);
            double ins_v = Mp[j] * t[T_MI] + Ip[j] * t[T_II];
            double del_v = Mc[j - 1] * t[T_MD] + Dc[j - 1] * t[T_DD];
            Mc[j] = m; Ic[j] = ins_v; Dc[j] = del_v;
            if (rescale) {
                if (m > rowmax) rowmax = m;
                if (ins_v > rowmax) rowmax = ins_v;
                if (del_v > rowmax) rowmax = del_v;
            }
        }
        if (rescale && rowmax > 0.0) {
            int e;
            frexp(rowmax, &e); /* rowmax = f * 2^e, f in [0.5,1) */
            int k = 1000 - e;  /* bring the row maximum back to ~2^1000 */
            if (k > 0) {
                for (int j = 0; j <= H; j++) { Mc[j] = ldexp(Mc[j], k); Ic[j] = ldexp(Ic[j], k); Dc[j] = ldexp(Dc[j], k); }
                shift += k;
            }
        }
        double *tmp;
        tmp = Mp; Mp = Mc; Mc = tmp;
        tmp = Ip; Ip = Ic; Ic = tmp;
        tmp = Dp; Dp = Dc; Dc = tmp;
    }
    /* pair_hmm.rs:598-614 */
    double final_sum = 0.0;
    if (R >= 1) {
        for (int j = 1; j <= H; j++) final_sum += Mp[j] + Ip[j];
    }
    free(buf);
    return log10(final_sum) - INITIAL_CONDITION_LOG10 - (double)shift * log10(2.0);
}

double lko_pairhmm_forward(const uint8_t *hap, int hap_len, const uint8_t *read, const uint8_t *qual,
                           const uint8_t *ins, const uint8_t *del, const uint8_t *gcp, int read_len,
                           int use_tristate) {
    return forward_impl(hap, hap_len, read, qual, ins, del, gcp, read_len, use_tristate, 0);
}

double lko_pairhmm_forward_rescaled(const uint8_t *hap, int hap_len, const uint8_t *read, const uint8_t *qual,
                                    const uint8_t *ins, const uint8_t *del, const uint8_t *gcp, int read_len,
                                    int use_tristate) {
    return forward_impl(hap, hap_len, read, qual, ins, del, gcp, read_len, use_tristate, 1);
}

int lko_find_first_position_where_haplotypes_differ(const uint8_t *h1, int n1, const uint8_t *h2, int n2) {
    int n = n1 < n2 ? n1 : n2;
    for (int i = 0; i < n; i++) if (h1[i] != h2[i]) return i;
    return n;
}

void lko_pairhmm_region(int n_reads, const int64_t *read_off, const uint8_t *read_bases,
                        const uint8_t *read_qual, const uint8_t *read_ins, const uint8_t *read_del,
                        const uint8_t *read_gcp, int n_haps, const int64_t *hap_off,
                        const uint8_t *hap_bases, int use_tristate, int rescaled, double *out) {
    /* pair_hmm.rs:349-369: for read { for hap { forward } } -> out[r*A + h] */
    for (int r = 0; r < n_reads; r++) {
        int64_t ro = read_off[r];
        int R = (int)(read_off[r + 1] - ro);
        for (int h = 0; h < n_haps; h++) {
            int64_t ho = hap_off[h];
            int H = (int)(hap_off[h + 1] - ho);
            out[(size_t)r * (size_t)n_haps + (size_t)h] =
                forward_impl(hap_bases + ho, H, read_bases + ro, read_qual + ro, read_ins + ro,
                             read_del + ro, read_gcp + ro, R, use_tristate, rescaled);
        }
    }
}

/* ------------------------------------------------------------------------ */
/* smith_waterman/smith_waterman_aligner.rs                                  */
/* ------------------------------------------------------------------------ */

#define SW_MATRIX_MIN_CUTOFF (-100000000) /* smith_waterman_aligner.rs:31 */

int lko_last_index_of(const uint8_t *ref, int ref_len, const uint8_t *query, int query_len) {
    /* alignment_utils.rs:717-735 */
    for (long r = (long)ref_len - (long)query_len; r >= 0; r--) {
        int q = 0;
        while (q < query_len && ref[r + q] == query[q]) q++;
        if (q == query_len) return (int)r;
    }
    return -1;
}

/* calculate_matrix, smith_waterman_aligner.rs:124-263. sw/btrack are (n+1)x(m+1) row-major. */
static void sw_calculate_matrix(const uint8_t *ref, int ref_len, const uint8_t *alt, int alt_len,
                                int32_t *sw, int32_t *btrack, int strategy, const int32_t params[4]) {
    const int nrow = ref_len + 1, ncol = alt_len + 1;
    const int32_t w_match = params[0], w_mismatch = params[1], w_open = params[2], w_extend = params[3];
    const int32_t low_init_value = INT32_MIN / 2; /* :139 */
    int32_t *best_gap_v = (int32_t *)malloc(sizeof(int32_t) * (size_t)(ncol + 1));
    int32_t *gap_size_v = (int32_t *)calloc((size_t)(ncol + 1), sizeof(int32_t));
    int32_t *best_gap_h = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nrow + 1));
    int32_t *gap_size_h = (int32_t *)calloc((size_t)(nrow + 1), sizeof(int32_t));
    for (int j = 0; j <= ncol; j++) best_gap_v[j] = low_init_value;
    for (int i = 0; i <= nrow; i++) best_gap_h[i] = low_init_value;

    /* :146-179 */
    if (strategy == LKO_SW_INDEL || strategy == LKO_SW_LEADING_INDEL) {
        int32_t cur = w_open;
        if (ncol > 1) sw[1] = cur;
        for (int j = 2; j < ncol; j++) { cur += w_extend; sw[j] = cur; }
        cur = w_open;
        if (nrow > 1) sw[(size_t)1 * ncol] = cur;
        for (int i = 2; i < nrow; i++) { cur += w_extend; sw[(size_t)i * ncol] = cur; }
    }

    /* :189-262 */
    for (int i = 1; i < nrow; i++) {
        uint8_t a_base = ref[i - 1];
        int32_t *cur_row = sw + (size_t)i * ncol;
        const int32_t *prev_row = sw + (size_t)(i - 1) * ncol;
        int32_t *bt_row = btrack + (size_t)i * ncol;
        for (int j = 1; j < ncol; j++) {
            uint8_t b_base = alt[j - 1];
            int32_t step_diag = prev_row[j - 1] + (a_base == b_base ? w_match : w_mismatch);

            int32_t prev_gap = prev_row[j] + w_open;
            best_gap_v[j] += w_extend;
            if (prev_gap > best_gap_v[j]) { best_gap_v[j] = prev_gap; gap_size_v[j] = 1; }
            else gap_size_v[j] += 1;
            int32_t step_down = best_gap_v[j];
            int32_t kd = gap_size_v[j];

            prev_gap = cur_row[j - 1] + w_open;
            best_gap_h[i] += w_extend;
            if (prev_gap > best_gap_h[i]) { best_gap_h[i] = prev_gap; gap_size_h[i] = 1; }
            else gap_size_h[i] += 1;
            int32_t step_right = best_gap_h[i];
            int32_t ki = gap_size_h[i];

            int diag_highest_or_equal = (step_diag >= step_down) && (step_diag >= step_right);
            if (diag_highest_or_equal) {
                cur_row[j] = step_diag > SW_MATRIX_MIN_CUTOFF ? step_diag : SW_MATRIX_MIN_CUTOFF;
                bt_row[j] = 0;
            } else if (step_right >= step_down) {
                cur_row[j] = step_right > SW_MATRIX_MIN_CUTOFF ? step_right : SW_MATRIX_MIN_CUTOFF;
                bt_row[j] = -ki;
            } else {
                cur_row[j] = step_down > SW_MATRIX_MIN_CUTOFF ? step_down : SW_MATRIX_MIN_CUTOFF;
                bt_row[j] = kd;
            }
        }
    }
    free(best_gap_v); free(gap_size_v); free(best_gap_h); free(gap_size_h);
}

enum { ST_MATCH = 0, ST_INS = 1, ST_DEL = 2, ST_CLIP = 3 };
static uint32_t sw_make_element(int state, uint32_t length) {
    /* :444-451 */
    static const uint32_t opcode[4] = { LKO_CIGAR_M, LKO_CIGAR_I, LKO_CIGAR_D, LKO_CIGAR_S };
    return (length << 4) | opcode[state];
}

/* calculate_cigar, smith_waterman_aligner.rs:273-442 */
static int sw_calculate_cigar(const int32_t *sw, const int32_t *btrack, int ref_length, int alt_length,
                              int strategy, uint32_t *ops_out, int max_ops, int32_t *alignment_offset) {
    const int ncol = alt_length + 1, nrow = ref_length + 1;
    int p1 = 0, p2;
    int32_t max_score = INT32_MIN;
    int32_t segment_length = 0;

    if (strategy == LKO_SW_INDEL) {
        p1 = ref_length; p2 = alt_length;
    } else {
        p2 = alt_length;
        for (int i = 1; i < nrow; i++) { /* :301-307 */
            int32_t cur = sw[(size_t)i * ncol + alt_length];
            if (cur >= max_score) { p1 = i; max_score = cur; }
        }
        if (strategy != LKO_SW_LEADING_INDEL) { /* :310-331 */
            const int32_t *bottom = sw + (size_t)ref_length * ncol;
            for (int j = 1; j < ncol; j++) {
                int32_t cur = bottom[j];
                if (cur > max_score ||
                    (cur == max_score && abs(ref_length - j) < abs(p1 - p2))) {
                    p1 = ref_length; p2 = j; max_score = cur;
                    segment_length = alt_length - j;
                }
            }
        }
    }

    /* reversed cigar accumulates in a growable temp */
    int cap = 64, n = 0;
    uint32_t *lce = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)cap);
#define LCE_PUSH(v) do { if (n == cap) { cap *= 2; lce = (uint32_t *)realloc(lce, sizeof(uint32_t) * (size_t)cap); } lce[n++] = (v); } while (0)

    if (segment_length > 0 && strategy == LKO_SW_SOFTCLIP) { /* :336-346 */
        LCE_PUSH(sw_make_element(ST_CLIP, (uint32_t)segment_length));
        segment_length = 0;
    }

    int state = ST_MATCH;
    for (;;) { /* :353-401 */
        int32_t btr = btrack[(size_t)p1 * ncol + p2];
        int new_state;
        int32_t step_length = 1;
        if (btr > 0) { new_state = ST_DEL; step_length = btr; }
        else if (btr < 0) { new_state = ST_INS; step_length = -btr; }
        else new_state = ST_MATCH;

        if (new_state == ST_MATCH) { p1 -= 1; p2 -= 1; }
        else if (new_state == ST_INS) p2 -= step_length;
        else p1 -= step_length;

        if (new_state == state) segment_length += step_length;
        else {
            if (segment_length > 0) LCE_PUSH(sw_make_element(state, (uint32_t)segment_length));
            segment_length = step_length;
            state = new_state;
        }
        if (p1 <= 0 || p2 <= 0) break;
    }

    /* :411-437 */
    if (strategy == LKO_SW_SOFTCLIP) {
        LCE_PUSH(sw_make_element(state, (uint32_t)segment_length));
        if (p2 > 0) LCE_PUSH(sw_make_element(ST_CLIP, (uint32_t)p2));
        *alignment_offset = p1;
    } else if (strategy == LKO_SW_IGNORE) {
        LCE_PUSH(sw_make_element(state, (uint32_t)(segment_length + p2)));
        *alignment_offset = p1 - p2;
    } else {
        LCE_PUSH(sw_make_element(state, (uint32_t)segment_length));
        if (p1 > 0) LCE_PUSH(sw_make_element(ST_DEL, (uint32_t)p1));
        else if (p2 > 0) LCE_PUSH(sw_make_element(ST_INS, (uint32_t)p2));
        *alignment_offset = 0;
    }
#undef LCE_PUSH
    int ret;
    if (n > max_ops) ret = -2;
    else { for (int k = 0; k < n; k++) ops_out[k] = lce[n - 1 - k]; ret = n; } /* :440 reverse */
    free(lce);
    return ret;
}

int lko_sw_align_full(const uint8_t *ref, int ref_len, const uint8_t *alt, int alt_len,
                      const int32_t params[4], int strategy, uint32_t *cigar_ops, int max_ops,
                      int32_t *alignment_offset) {
    if (ref_len <= 0 || alt_len <= 0) return -1; /* :65-68 assert */
    size_t cells = (size_t)(ref_len + 1) * (size_t)(alt_len + 1);
    int32_t *sw = (int32_t *)calloc(cells, sizeof(int32_t));
    int32_t *bt = (int32_t *)calloc(cells, sizeof(int32_t));
    sw_calculate_matrix(ref, ref_len, alt, alt_len, sw, bt, strategy, params);
    int n = sw_calculate_cigar(sw, bt, ref_len, alt_len, strategy, cigar_ops, max_ops, alignment_offset);
    free(sw); free(bt);
    return n;
}

int lko_sw_align(const uint8_t *ref, int ref_len, const uint8_t *alt, int alt_len,
                 const int32_t params[4], int strategy, uint32_t *cigar_ops, int max_ops,
                 int32_t *alignment_offset) {
    if (ref_len <= 0 || alt_len <= 0) return -1;
    if (strategy == LKO_SW_SOFTCLIP || strategy == LKO_SW_IGNORE) { /* :70-80,102-107 */
        int idx = lko_last_index_of(ref, ref_len, alt, alt_len);
        if (idx >= 0) {
            if (max_ops < 1) return -2;
            cigar_ops[0] = ((uint32_t)alt_len << 4) | LKO_CIGAR_M;
            *alignment_offset = idx;
            return 1;
        }
    }
    return lko_sw_align_full(ref, ref_len, alt, alt_len, params, strategy, cigar_ops, max_ops, alignment_offset);
}

/* ------------------------------------------------------------------------ */
/* pair_hmm_likelihood_calculation_engine.rs — pre-ops                       */
/* ------------------------------------------------------------------------ */

#define ENG_MAX_STR_UNIT_LENGTH 20 /* :98  */
#define ENG_MAX_REPEAT_LENGTH 100  /* :99  */
#define ENG_MIN_ADJUSTED_QSCORE 6  /* :100 */
#define ENG_INITIAL_QSCORE 40.0    /* :105 */

uint8_t lko_pcr_error_model_adjusted_qual(int repeat_length, double rate_factor) {
    /* :186-193.  Rust `as usize` saturates: negative / NaN -> 0. */
    double v = ENG_INITIAL_QSCORE - exp((double)repeat_length / (rate_factor * M_PI)) + 1.0;
    size_t u = (v > 0.0) ? (size_t)v : 0; /* truncation toward zero like `as usize` */
    if (!(v == v)) u = 0;
    size_t r = u > ENG_MIN_ADJUSTED_QSCORE ? u : ENG_MIN_ADJUSTED_QSCORE;
    return (uint8_t)r;
}

/* variant_context_utils.rs:276-337 */
static int find_number_of_repetitions_main(const uint8_t *unit_full, int unit_off, int unit_len,
                                           const uint8_t *test_full, int test_off, int test_len,
                                           int leading) {
    if (test_len == 0) return 0;
    long length_difference = (long)test_len - (long)unit_len;
    int num = 0;
    if (leading) {
        for (long start = 0; start <= length_difference; start += unit_len) {
            if (memcmp(test_full + start + test_off, unit_full + unit_off, (size_t)unit_len) == 0) num++;
            else return num;
        }
    } else {
        for (long start = length_difference; start >= 0; start -= unit_len) {
            if (memcmp(test_full + start + test_off, unit_full + unit_off, (size_t)unit_len) == 0) num++;
            else return num;
        }
    }
    return num;
}

/* VariantContextUtils::find_number_of_repetitions, variant_context_utils.rs:240-257 (the wrapper the reference's own
 * known answers call, tests/variant_context_utils_unit_tests.rs:22-140) */
int lko_find_number_of_repetitions(const uint8_t *unit, int unit_len, const uint8_t *test, int test_len, int leading) {
    if (test_len == 0) return 0;
    return find_number_of_repetitions_main(unit, 0, unit_len, test, 0, test_len, leading);
}

int lko_find_tandem_repeat_length(const uint8_t *bases, int len, int offset) {
    /* :528-611 */
    int max_bw = 0;
    const uint8_t *best_bw_unit = bases + offset;
    int best_bw_len = 1;
    for (int str = 1; str <= ENG_MAX_STR_UNIT_LENGTH; str++) {
        if (offset + 1 - str < 0) break;
        max_bw = find_number_of_repetitions_main(bases, offset + 1 - str, str, bases, 0, offset + 1, 0);
        if (max_bw > 1) { best_bw_unit = bases + (offset + 1 - str); best_bw_len = str; break; }
    }
    const uint8_t *best_unit = best_bw_unit;
    int best_len = best_bw_len;
    int max_rl = max_bw;

    if (offset < len - 1) {
        const uint8_t *best_fw_unit = bases + offset + 1;
        int best_fw_len = 1;
        int max_fw = 0;
        for (int str = 1; str <= ENG_MAX_STR_UNIT_LENGTH; str++) {
            if (offset + str + 1 > len) break;
            max_fw = find_number_of_repetitions_main(bases, offset + 1, str, bases, offset + 1, len - offset - 1, 1);
            if (max_fw > 1) { best_fw_unit = bases + offset + 1; best_fw_len = str; break; }
        }
        if (best_fw_len == best_len && memcmp(best_fw_unit, best_unit, (size_t)best_len) == 0) {
            max_rl = max_bw + max_fw;
        } else {
            /* find_number_of_repetitions(fw_unit, bases[0..=offset], trailing) :595-602 */
            max_bw = find_number_of_repetitions_main(best_fw_unit, 0, best_fw_len, bases, 0, offset + 1, 0);
            max_rl = max_fw + max_bw;
        }
    }
    if (max_rl > ENG_MAX_REPEAT_LENGTH) max_rl = ENG_MAX_REPEAT_LENGTH;
    return max_rl;
}

void lko_apply_pcr_error_model(const uint8_t *bases, int len, uint8_t *ins, uint8_t *del, int pcr_model) {
    /* :502-526 (+ cache :169-184) */
    if (pcr_model == 0) return;
    uint8_t cache[ENG_MAX_REPEAT_LENGTH + 1];
    for (int i = 0; i <= ENG_MAX_REPEAT_LENGTH; i++) cache[i] = lko_pcr_error_model_adjusted_qual(i, (double)pcr_model);
    for (int i = 1; i < len; i++) {
        int rl = lko_find_tandem_repeat_length(bases, len, i - 1);
        uint8_t c = cache[rl];
        if (c < ins[i - 1]) ins[i - 1] = c;
        if (c < del[i - 1]) del[i - 1] = c;
    }
}

void lko_cap_minimum_read_qualities(uint8_t *qual, uint8_t *ins, uint8_t *del, int len, uint8_t mapq,
                                    uint8_t base_quality_score_threshold, int disable_cap_to_mapq) {
    /* :428-466 */
    for (int i = 0; i < len; i++) {
        if (!disable_cap_to_mapq && mapq < qual[i]) qual[i] = mapq;
        if (qual[i] < base_quality_score_threshold) qual[i] = LKO_MIN_USABLE_Q_SCORE;
        if (ins[i] < LKO_MIN_USABLE_Q_SCORE) ins[i] = LKO_MIN_USABLE_Q_SCORE;
        if (del[i] < LKO_MIN_USABLE_Q_SCORE) del[i] = LKO_MIN_USABLE_Q_SCORE;
    }
}

/* ------------------------------------------------------------------------ */
/* post-ops                                                                  */
/* ------------------------------------------------------------------------ */

void lko_normalize_likelihoods(double *values, int A, int R, double cap_diff, int symmetric, int ref_idx) {
    /* allele_likelihoods.rs:378-444 */
    if (cap_diff == -INFINITY) return;
    if (A == 0 || A == 1) return;
    for (int r = 0; r < R; r++) {
        /* search_best_allele(.., can_be_reference = symmetric, priorities None) :457-554 */
        int can_be_reference = symmetric;
        int best = (can_be_reference || ref_idx != 0) ? 0 : 1;
        double best_lk = values[(size_t)best * R + r];
        double second = -INFINITY;
        for (int a = best + 1; a < A; a++) {
            if (!can_be_reference && ref_idx == a) continue;
            double c = values[(size_t)a * R + r];
            if (c > best_lk) { second = best_lk; best = a; best_lk = c; }
            else if (c > second) { second = c; }
        }
        best_lk = values[(size_t)best * R + r];
        double worst_cap = best_lk + cap_diff;
        for (int a = 0; a < A; a++) {
            double *v = &values[(size_t)a * R + r];
            if (*v < worst_cap) *v = worst_cap;
        }
    }
}

/* pair_hmm_likelihood_calculation_engine.rs:23-39 */
static const double dyn_table[] = {
    1.0, 5.996842844, 0.196616587, 2.0, 5.870018422, 1.388545569, 3.0, 5.401558531, 5.641990128,
    4.0, 4.818940919, 10.33176216, 5.0, 4.218758304, 14.25799688, 6.0, 3.646319832, 17.02880749,
    7.0, 3.122346753, 18.64537883, 8.0, 2.654731979, 19.27521677, 9.0, 2.244479156, 19.13584613,
    10.0, 1.88893867, 18.43922003, 11.0, 1.583645342, 17.36842261, 12.0, 1.3233807, 16.07088712,
    13.0, 1.102785365, 14.65952563, 14.0, 0.916703025, 13.21718577, 15.0, 0.760361881, 11.80207947,
    16.0, 0.629457387, 10.45304833, 17.0, 0.520175654, 9.194183767, 18.0, 0.42918208, 8.038657241,
    19.0, 0.353590663, 6.991779595, 20.0, 0.290923699, 6.053379213, 21.0, 0.23906788, 5.219610436,
    22.0, 0.196230431, 4.484302033, 23.0, 0.160897421, 3.839943445, 24.0, 0.131795374, 3.27839108,
    25.0, 0.1078567, 2.791361596, 26.0, 0.088189063, 2.370765375, 27.0, 0.072048567, 2.008921719,
    28.0, 0.058816518, 1.698687797, 29.0, 0.047979438, 1.433525748, 30.0, 0.039111985, 1.207526336,
    31.0, 0.031862437, 1.015402928, 32.0, 0.025940415, 0.852465956, 33.0, 0.021106532, 0.714585285,
    34.0, 0.017163711, 0.598145851, 35.0, 0.013949904, 0.500000349, 36.0, 0.011332027, 0.41742159,
    37.0, 0.009200898, 0.348056286, 38.0, 0.007467036, 0.289881373, 39.0, 0.006057179, 0.241163527,
    40.0, 0.004911394, 0.200422214
};

double lko_log10_dynamic_read_qual_threshold(const uint8_t *qual, int len, double k) {
    /* :261-291 */
    double sum_mean = 0.0, sum_var = 0.0;
    for (int i = 0; i < len; i++) {
        int bq = qual[i];
        int entry = (bq <= 1) ? 0 : ((bq < 40 ? bq : 40) - 1);
        int mean_off = entry * 3 + 1;
        sum_mean += dyn_table[mean_off];
        sum_var += dyn_table[mean_off + 1];
    }
    double threshold = sum_mean + k * sqrt(sum_var);
    return threshold * -0.1;
}

double lko_log10_min_true_likelihood(int read_len, double max_error_per_base, int cap_likelihoods) {
    /* :293-319 */
    double max_errors = ceil((double)read_len * max_error_per_base);
    if (cap_likelihoods && max_errors > 2.0) max_errors = 2.0;
    return max_errors * -4.0;
}

int lko_filter_poorly_modeled_evidence(double *values, int A, int R, const double *thresholds, uint8_t *keep) {
    /* allele_likelihoods.rs:925-1018 */
    int removed = 0;
    for (int r = 0; r < R; r++) {
        double best = -INFINITY; /* maximum_likelihood_over_all_alleles :1024-1040 */
        for (int a = 0; a < A; a++) if (values[(size_t)a * R + r] > best) best = values[(size_t)a * R + r];
        keep[r] = !(best < thresholds[r]);
        if (!keep[r]) removed++;
    }
    if (removed == 0) return R;
    int new_count = R - removed;
    int w = 0;
    for (int r = 0; r < R; r++) {
        if (!keep[r]) continue;
        for (int a = 0; a < A; a++) values[(size_t)a * R + w] = values[(size_t)a * R + r];
        w++;
    }
    for (int a = 0; a < A; a++)
        for (int r = new_count; r < R; r++) values[(size_t)a * R + r] = NAN;
    return new_count;
}
