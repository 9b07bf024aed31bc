// best_allele.h — AlleleLikelihoods::search_best_allele(sample, r, can_be_reference = true, Some(priorities)) for ONE column of
// an [A x R] row-major block (src/model/allele_likelihoods.rs:457-554) and BestAllele::new / is_informative (:1141-1166).
// One definition shared by the host entry point lk_best_alleles_breaking_ties (host_cigar.cpp) and the device kernel that
// runs the same search on matrices that stayed in HBM (engine_kernels.cu): plain IEEE f64 compares and subtractions, so both
// make the same decisions bit for bit.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LK_BA_HD __host__ __device__ __forceinline__
#else
#define LK_BA_HD inline
#endif

namespace lk {

LK_BA_HD void best_allele_of_read(const double *values, int64_t n_alleles, int64_t n_reads, int64_t r, const int32_t *priorities,
                                  double informative_threshold, int32_t *best_allele, double *confidence, uint8_t *informative) {
    if (n_alleles == 0) {  // BestAllele with allele None
        *best_allele = -1;
        if (confidence) *confidence = 0.0;
        if (informative) *informative = 0;
        return;
    }
    const double *col = values + r;
    int64_t best = 0, second = 0;
    double best_lk = col[0], second_lk = -INFINITY;
    for (int64_t a = 1; a < n_alleles; a++) {
        const double c = col[a * n_reads];
        if (c > best_lk) { second = best; best = a; second_lk = best_lk; best_lk = c; }
        else if (c > second_lk) { second = a; second_lk = c; }
    }
    if (priorities && (best_lk - second_lk) < informative_threshold) {
        int32_t best_pr = priorities[best], second_pr = priorities[second];
        for (int64_t a = 0; a < n_alleles; a++) {
            const double c = col[a * n_reads];
            if (a == best || (best_lk - c) > informative_threshold) continue;
            const int32_t pr = priorities[a];
            if (pr > best_pr) { second = best; best = a; second_pr = best_pr; best_pr = pr; }
            else if (pr > second_pr) { second = a; second_pr = pr; }
        }
    }
    best_lk = col[best * n_reads];
    second_lk = (second != best) ? col[second * n_reads] : -INFINITY;
    const double diff = best_lk - second_lk;
    const double conf = (fabs(diff) < 2.220446049250313e-16) ? 0.0 : diff;  // f64::EPSILON
    *best_allele = (int32_t)best;
    if (confidence) *confidence = conf;
    if (informative) *informative = conf > informative_threshold ? 1 : 0;
}

// One [A x R] block of a realignment batch, as the device kernel sees it.
struct BestAlleleBlock {
    int64_t values_off;   // first element of the block inside `values`
    int64_t first_read;   // batch-global index of the block's first read
    int64_t prio_off;     // first priority of the block's region inside `priorities`
    int32_t n_alleles, n_reads, n_evidence, pad;
};

}  // namespace lk
