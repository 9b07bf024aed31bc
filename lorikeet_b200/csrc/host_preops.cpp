// host_preops.cpp — host side of PairHMMLikelihoodCalculationEngine::modify_read_qualities (the per-read byte work that
// precedes the PairHMM call).  Pure host code, no CUDA: it is compiled into the product library AND into the separate
// generator library liblorikeet_b200_synth.so (so bench.py's reference arm never maps the product library).
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
// lk_engine_modify_read_qualities is the stand-alone host utility (in place on the caller's arrays); the region driver
// lk_engine_compute_read_likelihoods runs the same pre-ops and the post-ops on the device (engine_kernels.cu, N3).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../../include/lorikeet_b200.h"
#include "host_preops.h"

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

// VariantContextUtils::find_number_of_repetitions (variant_context_utils.rs:240-257) as a host utility: the repeat counter the
// PCR error model is built on, exposed so that it can be pinned on the reference's own known answers
// (tests/variant_context_utils_unit_tests.rs:22-140).  -1: bad arguments (the reference asserts a non-empty unit).
extern "C" int64_t lk_find_number_of_repetitions(const uint8_t *repeat_unit, int64_t unit_len, const uint8_t *test_string,
                                                 int64_t test_len, int leading_repeats) {
    if (!repeat_unit || unit_len <= 0 || test_len < 0 || (test_len > 0 && !test_string) || unit_len > 0x7fffffff || test_len > 0x7fffffff) return -1;
    return count_repeats(Span{ repeat_unit, (int)unit_len }, test_string, (int)test_len, leading_repeats != 0);
}

void lk_engine_fill_pcr_cache(const lk_engine_args *args, uint8_t cache[101]) {  // initialize_pcr_error_model :169-184
    memset(cache, 0, 101);
    if (args->pcr_error_model != LK_PCR_NONE)
        for (int i = 0; i <= kMaxRepeatLength; i++) cache[i] = adjusted_qual(i, (double)args->pcr_error_model);
}

