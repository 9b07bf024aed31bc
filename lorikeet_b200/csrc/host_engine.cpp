// host_engine.cpp — host side of PairHMMLikelihoodCalculationEngine (the per-read byte work and the
// per-region post-ops that bracket the GPU PairHMM call).
//
// Reference: src/pair_hmm/pair_hmm_likelihood_calculation_engine.rs
//   new / initialize_pcr_error_model / get_error_model_adjusted_qual   :129-193
//   compute_read_likelihoods                                           :195-242
//   dynamic / static disqualification thresholds                       :244-319 (table :23-39)
//   modify_read_qualities, cap_minimum_read_qualities                  :352-466
//   apply_pcr_error_model, find_tandem_repeat_units                    :502-611
// and src/model/allele_likelihoods.rs normalize_likelihoods :378-444, search_best_allele :457-554,
// filter_poorly_modeled_evidence :925-1018; src/model/variant_context_utils.rs:240-348.
//
// Kept on the host on purpose (SURVEY 8a/a2,a7): the work is O(reads * R * 20) byte compares with side effects on
// the caller's read lists; the GPU part is the (read x haplotype) likelihood block in between.
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../../include/lorikeet_b200.h"
#include "lk_api_internal.h"

namespace {

constexpr int kMaxStrUnitLength = 20;   // :98
constexpr int kMaxRepeatLength = 100;   // :99
constexpr int kMinAdjustedQscore = 6;   // :100
constexpr double kInitialQscore = 40.0; // :105
constexpr uint8_t kMinUsableQ = 6;      // quality_utils.rs:23

// max(6, (40 - e^{n/(rate*pi)} + 1) as usize): the Rust float->usize cast saturates negatives to 0 (:186-193)
uint8_t adjusted_qual(int repeat_length, double rate_factor) {
    const double v = kInitialQscore - std::exp((double)repeat_length / (rate_factor * M_PI)) + 1.0;
    const long t = (v > 0.0 && v == v) ? (long)v : 0;
    return (uint8_t)std::max<long>(kMinAdjustedQscore, t);
}

struct Span {
    const uint8_t *p;
    int n;
    bool operator==(const Span &o) const { return n == o.n && memcmp(p, o.p, (size_t)n) == 0; }
};

// VariantContextUtils::find_number_of_repetitions_main (variant_context_utils.rs:276-337)
int count_repeats(Span unit, const uint8_t *test, int test_len, bool leading) {
    if (test_len == 0) return 0;
    const long diff = (long)test_len - unit.n;
    int reps = 0;
    if (leading) {
        for (long s = 0; s <= diff; s += unit.n) {
            if (memcmp(test + s, unit.p, (size_t)unit.n) != 0) break;
            reps++;
        }
    } else {
        for (long s = diff; s >= 0; s -= unit.n) {
            if (memcmp(test + s, unit.p, (size_t)unit.n) != 0) break;
            reps++;
        }
    }
    return reps;
}

// find_tandem_repeat_units(...).1 (:528-611)
int tandem_repeat_length(const uint8_t *bases, int len, int offset) {
    int max_bw = 0;
    Span best_bw{ bases + offset, 1 };
    for (int str = 1; str <= kMaxStrUnitLength; str++) {
        if (offset + 1 < str) break;
        const Span unit{ bases + offset + 1 - str, str };
        max_bw = count_repeats(unit, bases, offset + 1, false);
        if (max_bw > 1) { best_bw = unit; break; }
    }
    int max_rl = max_bw;
    if (offset < len - 1) {
        Span best_fw{ bases + offset + 1, 1 };
        int max_fw = 0;
        for (int str = 1; str <= kMaxStrUnitLength; str++) {
            if (offset + str + 1 > len) break;
            const Span unit{ bases + offset + 1, str };
            max_fw = count_repeats(unit, bases + offset + 1, len - offset - 1, true);
            if (max_fw > 1) { best_fw = unit; break; }
        }
        if (best_fw == best_bw) {
            max_rl = max_bw + max_fw;
        } else {
            max_bw = count_repeats(best_fw, bases, offset + 1, false);
            max_rl = max_fw + max_bw;
        }
    }
    return std::min(max_rl, kMaxRepeatLength);
}

void modify_one_read(const uint8_t *cache, bool use_pcr, const uint8_t *bases, uint8_t *qual, uint8_t *ins, uint8_t *del,
                     int len, uint8_t mapq, uint8_t bq_threshold, bool cap_to_mapq) {
    if (use_pcr) {  // apply_pcr_error_model :502-526 (note: the last base is never touched)
        for (int i = 1; i < len; i++) {
            const uint8_t c = cache[tandem_repeat_length(bases, len, i - 1)];
            ins[i - 1] = std::min(ins[i - 1], c);
            del[i - 1] = std::min(del[i - 1], c);
        }
    }
    for (int i = 0; i < len; i++) {  // cap_minimum_read_qualities :428-458
        if (cap_to_mapq) qual[i] = std::min(qual[i], mapq);
        if (qual[i] < bq_threshold) qual[i] = kMinUsableQ;
        if (ins[i] < kMinUsableQ) ins[i] = kMinUsableQ;
        if (del[i] < kMinUsableQ) del[i] = kMinUsableQ;
    }
}

// :23-39, rows (baseQ, mean, variance) for baseQ 1..40
const double kDynTable[] = {
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

// calculate_log10_dynamic_read_qual_threshold :261-291
double dynamic_threshold(const uint8_t *qual, int len, double k) {
    double mean = 0.0, var = 0.0;
    for (int i = 0; i < len; i++) {
        const int bq = qual[i];
        const int entry = bq <= 1 ? 0 : std::min(40, bq) - 1;
        mean += kDynTable[entry * 3 + 1];
        var += kDynTable[entry * 3 + 2];
    }
    return (mean + k * std::sqrt(var)) * -0.1;
}

// log10_min_true_likelihood :293-319
double min_true_likelihood(int len, double max_error_per_base, bool cap) {
    double e = std::ceil((double)len * max_error_per_base);
    if (cap) e = std::min(2.0, e);
    return e * -4.0;
}

template <class F>
void parallel_for(int64_t n, int n_threads, F f) {
    if (n_threads <= 1 || n < 2) { for (int64_t i = 0; i < n; i++) f(i); return; }
    std::atomic<int64_t> next{ 0 };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++)
        th.emplace_back([&]() {
            for (;;) {
                const int64_t i0 = next.fetch_add(64);
                if (i0 >= n) break;
                for (int64_t i = i0; i < std::min(n, i0 + 64); i++) f(i);
            }
        });
    for (auto &t : th) t.join();
}

}  // namespace

extern "C" lk_status lk_engine_modify_read_qualities(const lk_engine_args *args, int64_t n_reads, const int64_t *read_off,
                                                     const uint8_t *read_bases, uint8_t *read_qual, uint8_t *read_ins,
                                                     uint8_t *read_del, const uint8_t *read_mapq, int n_threads) {
    if (!args || n_reads < 0) return LK_ERR_INVALID;
    if (n_reads == 0) return LK_OK;
    if (!read_off || !read_bases || !read_qual || !read_ins || !read_del) return LK_ERR_INVALID;
    if (!args->disable_cap_read_qualities_to_mapq && !read_mapq) return LK_ERR_INVALID;
    if (args->pcr_error_model < 0 || args->pcr_error_model > 3) return LK_ERR_INVALID;
    uint8_t cache[kMaxRepeatLength + 1];
    memset(cache, 0, sizeof cache);
    const bool use_pcr = args->pcr_error_model != LK_PCR_NONE;
    if (use_pcr)  // initialize_pcr_error_model :169-184
        for (int i = 0; i <= kMaxRepeatLength; i++) cache[i] = adjusted_qual(i, (double)args->pcr_error_model);
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    const bool cap = !args->disable_cap_read_qualities_to_mapq;
    parallel_for(n_reads, n_threads, [&](int64_t r) {
        const int64_t o = read_off[r];
        modify_one_read(cache, use_pcr, read_bases + o, read_qual + o, read_ins + o, read_del + o,
                        (int)(read_off[r + 1] - o), cap ? read_mapq[r] : 255, args->base_quality_score_threshold, cap);
    });
    return LK_OK;
}

// normalize_likelihoods for one [A x R] block (allele_likelihoods.rs:378-444; search_best_allele :457-554 with priorities None)
static void normalize_block(double *v, int A, int R, double cap_diff, bool symmetric, int ref_idx) {
    if (cap_diff == -std::numeric_limits<double>::infinity() || A < 2) return;
    for (int r = 0; r < R; r++) {
        const bool can_be_ref = symmetric;
        int best = (can_be_ref || ref_idx != 0) ? 0 : 1;
        double best_lk = v[(size_t)best * R + r];
        for (int a = best + 1; a < A; a++) {
            if (!can_be_ref && a == ref_idx) continue;
            if (v[(size_t)a * R + r] > best_lk) { best = a; best_lk = v[(size_t)a * R + r]; }
        }
        const double worst = best_lk + cap_diff;
        for (int a = 0; a < A; a++)
            if (v[(size_t)a * R + r] < worst) v[(size_t)a * R + r] = worst;
    }
}

// filter_poorly_modeled_evidence + remove_evidence_by_index for one block (:925-1018)
static int filter_block(double *v, int A, int R, const double *thr, uint8_t *keep) {
    int kept = 0;
    for (int r = 0; r < R; r++) {
        double best = -std::numeric_limits<double>::infinity();
        for (int a = 0; a < A; a++) best = std::max(best, v[(size_t)a * R + r]);
        const bool remove = best < thr[r];
        if (keep) keep[r] = remove ? 0 : 1;
        if (!remove) {
            if (kept != r)
                for (int a = 0; a < A; a++) v[(size_t)a * R + kept] = v[(size_t)a * R + r];
            kept++;
        }
    }
    if (kept < R)
        for (int a = 0; a < A; a++)
            for (int r = kept; r < R; r++) v[(size_t)a * R + r] = std::numeric_limits<double>::quiet_NaN();
    return kept;
}

extern "C" lk_status lk_engine_compute_read_likelihoods(lk_ctx *ctx, const lk_engine_args *args, const lk_pairhmm_batch *batch,
                                                        const uint8_t *read_mapq, const int32_t *ref_allele_index,
                                                        double *values, uint8_t *keep) {
    if (!ctx || !args || !batch) return LK_ERR_INVALID;
    int64_t n_pairs = 0;
    lk_status st = lk_pairhmm_batch_dims(batch, &n_pairs, nullptr);
    if (st != LK_OK) return lk_fail(ctx, st, "invalid batch");
    if (n_pairs > 0 && !values) return lk_fail(ctx, LK_ERR_INVALID, "null values");
    const int64_t n_reads = batch->region_read_start[batch->n_regions];
    const int64_t nbytes = n_reads ? batch->read_off[n_reads] : 0;
    // processed copies; the originals stay untouched for the thresholds (the "HMMQuals" key quirk: the capped
    // qualities are stored under a key nobody reads, so :268,299 fall back to the ORIGINAL quals / length)
    std::vector<uint8_t> q(batch->read_qual, batch->read_qual + nbytes), qi(batch->read_ins, batch->read_ins + nbytes),
        qd(batch->read_del, batch->read_del + nbytes);
    st = lk_engine_modify_read_qualities(args, n_reads, batch->read_off, batch->read_bases, q.data(), qi.data(), qd.data(),
                                         read_mapq, 0);
    if (st != LK_OK) return lk_fail(ctx, st, "modify_read_qualities failed");
    lk_pairhmm_batch pb = *batch;
    pb.read_qual = q.data(); pb.read_ins = qi.data(); pb.read_del = qd.data();
    pb.read_gcp = nullptr;  // gap_continuation_penalties is vec![constant_gcp; len] (:649-651)
    pb.gcp_constant = args->constant_gcp;
    std::vector<double> lk((size_t)n_pairs);
    st = lk_pairhmm_batch_host(ctx, &pb, lk.data(), nullptr);
    if (st != LK_OK) return st;
    // per region: [r*A + h] -> [a][r] (pair_hmm.rs:247-266), normalise, filter
    std::vector<int64_t> base((size_t)batch->n_regions + 1, 0);
    for (int64_t g = 0; g < batch->n_regions; g++)
        base[(size_t)g + 1] = base[(size_t)g] + (batch->region_read_start[g + 1] - batch->region_read_start[g]) *
                                                    (batch->region_hap_start[g + 1] - batch->region_hap_start[g]);
    parallel_for(batch->n_regions, (int)std::max(1u, std::thread::hardware_concurrency()), [&](int64_t g) {
        const int64_t r0 = batch->region_read_start[g];
        const int R = (int)(batch->region_read_start[g + 1] - r0);
        const int A = (int)(batch->region_hap_start[g + 1] - batch->region_hap_start[g]);
        if (R == 0 || A == 0) return;
        double *v = values + base[(size_t)g];
        const double *src = lk.data() + base[(size_t)g];
        for (int r = 0; r < R; r++)
            for (int a = 0; a < A; a++) v[(size_t)a * R + r] = src[(size_t)r * A + a];
        normalize_block(v, A, R, args->log10_global_read_mismapping_rate,
                        args->symmetrically_normalize_alleles_to_reference != 0, ref_allele_index ? ref_allele_index[g] : 0);
        std::vector<double> thr((size_t)R);
        for (int r = 0; r < R; r++) {
            const int64_t o = batch->read_off[r0 + r];
            const int len = (int)(batch->read_off[r0 + r + 1] - o);
            if (args->dynamic_read_disqualification) {  // dynamic_log10_min_likelihood_model :244-259
                const double dyn = dynamic_threshold(batch->read_qual + o, len, args->read_disqualification_scale);
                const double stat = min_true_likelihood(len, args->expected_error_rate_per_base, false);
                thr[(size_t)r] = dyn < stat ? dyn : stat;
            } else {
                thr[(size_t)r] = min_true_likelihood(len, args->expected_error_rate_per_base, true);
            }
        }
        filter_block(v, A, R, thr.data(), keep ? keep + r0 : nullptr);
    });
    return LK_OK;
}
