// host_model.cpp — host-side tables of the PairHMM model, uploaded once per context.
//
// Follows the reference's definitions:
//   QualityUtils::qual_to_error_prob            src/utils/quality_utils.rs:98-104   (10^(-q/10))
//   JacobianLogTable / approximate_log10_sum    src/utils/math_utils.rs:8-14,314-334,484-498
//   PairHMMModel::new / match_to_match_prob     src/pair_hmm/pair_hmm_model.rs:47-78,442-461
#include "host_model.h"

#include <cmath>
#include <vector>

namespace lk {

namespace {
constexpr double kMaxTolerance = 8.0;   // math_utils.rs:487
constexpr double kTableStep = 0.0001;   // :492
constexpr int kMaxQual = 254;           // quality_utils.rs:24

struct Jacobian {
    std::vector<double> cache;
    Jacobian() {
        const int n = (int)((kMaxTolerance / kTableStep) + 1.0);
        cache.resize((size_t)n);
        for (int k = 0; k < n; k++) cache[(size_t)k] = std::log10(1.0 + std::pow(10.0, -(double)k * kTableStep));
    }
    double approx_log10_sum(double a, double b) const {
        if (a > b) std::swap(a, b);
        if (a == -INFINITY) return b;
        const double diff = b - a;
        if (diff >= kMaxTolerance) return b;
        return b + cache[(size_t)std::llround(diff * (1.0 / kTableStep))];
    }
};
}  // namespace

void build_model_tables(ModelTables &t) {
    t.qerr_d.resize(256);
    t.qerr_f.resize(256);
    for (int q = 0; q < 256; q++) {
        t.qerr_d[(size_t)q] = std::pow(10.0, (double)q / -10.0);
        t.qerr_f[(size_t)q] = (float)t.qerr_d[(size_t)q];
    }
    // triangular table over 0..255: rows 0..254 follow PairHMMModel::new; row 255 (> MAX_QUAL) uses the
    // closed form of match_to_match_prob's first branch (pair_hmm_model.rs:453-458).
    const Jacobian jac;
    const double inv_ln10 = 1.0 / std::log(10.0);
    t.m2m.assign((size_t)(256 * 257 / 2), 0.0);
    size_t offset = 0;
    for (int i = 0; i <= 255; i++) {
        for (int j = 0; j <= i; j++) {
            const double log10_sum = jac.approx_log10_sum(-0.1 * (double)i, -0.1 * (double)j);
            const double p = std::pow(10.0, log10_sum);
            double v;
            if (i <= kMaxQual) v = std::pow(10.0, std::log1p(-std::fmin(1.0, p)) * inv_ln10);
            else v = 1.0 - p;
            t.m2m[offset + (size_t)j] = v;
        }
        offset += (size_t)i + 1;
    }
}

}  // namespace lk
