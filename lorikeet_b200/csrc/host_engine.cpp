// host_engine.cpp — the region driver compute_read_likelihoods behind the C ABI.
//
// Reference: src/pair_hmm/pair_hmm_likelihood_calculation_engine.rs
//   compute_read_likelihoods                                           :195-242
//   dynamic / static disqualification thresholds                       :244-319 (table :23-39)
// and src/model/allele_likelihoods.rs normalize_likelihoods :378-444, search_best_allele :457-554,
// filter_poorly_modeled_evidence :925-1018.  The per-read pre-ops are in host_preops.cpp (host utility) and
// engine_kernels.cu (what the driver runs).
#include <cstring>

#include "../../include/lorikeet_b200.h"
#include "host_preops.h"
#include "lk_api_internal.h"

// Region driver.  Everything runs on the device inside the PairHMM chunk pipeline (lk_api.cu: lk_pairhmm_pipeline with a
// hook): engine_modify_reads_kernel on the staged copies (the caller's arrays stay untouched), the PairHMM kernels,
// engine_epilogue_kernel (transpose to [allele][read], normalise, thresholds from the ORIGINAL qualities and lengths --
// the "HMMQuals" key quirk: the capped qualities are stored under a key nobody reads, so :268,299 fall back to the
// original quals / length -- filter, compact, NaN tail).  See engine_kernels.cu.
extern "C" lk_status lk_engine_compute_read_likelihoods(lk_ctx *ctx, const lk_engine_args *args, const lk_pairhmm_batch *batch,
                                                        const uint8_t *read_mapq, const int32_t *ref_allele_index,
                                                        double *values, uint8_t *keep) {
    if (!ctx || !args || !batch) return LK_ERR_INVALID;
    int64_t n_pairs = 0;
    lk_status st = lk_pairhmm_batch_dims(batch, &n_pairs, nullptr);
    if (st != LK_OK) return lk_fail(ctx, st, "invalid batch");
    if (n_pairs > 0 && !values) return lk_fail(ctx, LK_ERR_INVALID, "null values");
    if (!args->disable_cap_read_qualities_to_mapq && !read_mapq && batch->region_read_start[batch->n_regions] > 0)
        return lk_fail(ctx, LK_ERR_INVALID, "read_mapq is required unless capping to MAPQ is disabled");
    if (args->pcr_error_model < 0 || args->pcr_error_model > 3) return lk_fail(ctx, LK_ERR_INVALID, "bad PCR error model");
    if ((batch->region_sample_start != nullptr) != (batch->sample_read_start != nullptr))
        return lk_fail(ctx, LK_ERR_INVALID, "region_sample_start and sample_read_start go together");
    if (batch->region_sample_start) {  // the samples of a region tile its reads in order (:219-224 loops the samples)
        const int64_t *rs = batch->region_sample_start, *sr = batch->sample_read_start;
        if (rs[0] != 0) return lk_fail(ctx, LK_ERR_INVALID, "region_sample_start must start at 0");
        for (int64_t g = 0; g < batch->n_regions; g++) {
            if (rs[g + 1] < rs[g]) return lk_fail(ctx, LK_ERR_INVALID, "region_sample_start must be non-decreasing");
            const int64_t r0 = batch->region_read_start[g], r1 = batch->region_read_start[g + 1];
            if (rs[g + 1] == rs[g]) {
                if (r1 != r0) return lk_fail(ctx, LK_ERR_INVALID, "a region with reads needs at least one sample");
                continue;
            }
            if (sr[rs[g]] != r0 || sr[rs[g + 1]] != r1) return lk_fail(ctx, LK_ERR_INVALID, "samples must tile their region's reads");
            for (int64_t sm = rs[g]; sm < rs[g + 1]; sm++)
                if (sr[sm + 1] < sr[sm]) return lk_fail(ctx, LK_ERR_INVALID, "sample_read_start must be non-decreasing");
        }
    }
    lk_engine_hook hook;
    memset(&hook, 0, sizeof hook);
    hook.args = args;
    hook.read_mapq = args->disable_cap_read_qualities_to_mapq ? nullptr : read_mapq;
    hook.ref_allele_index = ref_allele_index;
    hook.values = values;
    hook.keep = keep;
    lk_engine_fill_pcr_cache(args, hook.pcr_cache);
    lk_pairhmm_batch pb = *batch;
    pb.read_gcp = nullptr;  // gap_continuation_penalties is vec![constant_gcp; len] (:649-651)
    pb.gcp_constant = args->constant_gcp;
    return lk_pairhmm_pipeline(ctx, &pb, nullptr, nullptr, &hook);
}
