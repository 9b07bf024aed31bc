// lk_api_internal.h — host-side structures behind the opaque handles of include/lorikeet_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <string>
#include <vector>

#include "../../include/lorikeet_b200.h"
#include "host_prep.h"
#include "host_preops.h"

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    bool view = false;  // p points into another DevBuf (the merged metadata block): never freed here
    cudaError_t reserve(size_t bytes);
    void release();
    void set_view(void *q) { if (!view && p) cudaFree(p); p = q; cap = 0; view = true; }
};

struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes);
    void release();
};

struct lk_resident;
constexpr int LK_PIPE_DEPTH = 3;  // == number of streams in lk_ctx

struct lk_ctx {
    int device = 0;
    int n_sm = 148;
    cudaStream_t stream[3] = { nullptr, nullptr, nullptr };
    std::mutex mu;
    std::string err;
    bool tristate = true;
    bool force_f64 = false;
    int force_k = 0;
    int64_t split_target_tasks = 0;  // launches with fewer warp tasks are cut along the haplotype axis (build_prep)
    int32_t last_launch_count = 0;
    DevBuf d_qerr_f, d_qerr_d, d_m2m;
    DevBuf d_ba_blocks, d_ba_prio, d_ba_best, d_ba_info;  // grow-only buffers of the device best-allele search (N2)
    // reusable workspaces of lk_pairhmm_batch_host's chunk pipeline, one per stream.  Three, not two: chunk i+2's H2D then
    // waits on chunk i-1 (long finished) instead of chunk i, whose small f64 pass can be held back behind chunk i+1's
    // main kernel and would otherwise delay the copy, and with it chunk i+2's kernels, past the end of chunk i+1.
    lk_resident *ws[3] = { nullptr, nullptr, nullptr };
    lk_resident *sw_ws[3] = { nullptr, nullptr, nullptr };  // same for lk_sw_batch_host
    // page-locked, grow-only staging of the SW batches the callers in host_cigar.cpp assemble (padded pairs in, raw CIGARs out),
    // so that lk_sw_batch_host's chunk copies are truly asynchronous; one such batch at a time per context
    std::mutex stage_mu;
    PinnedBuf stage_refs, stage_alts, stage_ops, stage_n, stage_off;
};

struct lk_resident {
    int kind = 0;  // 0 PairHMM, 1 Smith-Waterman
    // ---- PairHMM ----
    int64_t n_regions = 0, n_reads = 0, n_haps = 0, n_pairs = 0, n_cells = 0;
    int h_max = 0;
    bool vargcp = false, all_f64 = false;
    uint8_t gcp_constant = 10;
    int32_t n_direct = 0;
    int32_t counter_init[3] = { 0, 0, 0 };
    int32_t n_band = 0;
    int band_blocks = 0;
    int64_t band_stride = 0;
    int plain_blocks = 0;
    int64_t scratch_stride = 0;
    int32_t launch_count = 0;
    float last_fast_ms = 0.f;
    cudaEvent_t ev_fast0 = nullptr, ev_fast1 = nullptr;
    std::vector<lk::FastLaunch> launches;
    lk::HostPrep prep;
    PinnedBuf pinned;                 // metadata staging (so its H2D copies are truly asynchronous)
    cudaEvent_t ev_staged = nullptr;  // pinned arena may be rewritten once this has fired
    int64_t read_bytes_staged = 0;
    int64_t read_byte0 = 0, hap_byte0 = 0;  // first arena byte staged: device pointers are biased by these
    DevBuf read_bases, read_qual, read_ins, read_del, read_gcp, read_off, read_region, read_out_base, read_flags;
    DevBuf region_hap_start, hap_off, hap_bases, hidx, henc_off, hap_flags;
    DevBuf tasks, task_haps, groups, out, flags, counters, rescue_pairs, scratch, band_pairs, band_scratch;
    // engine pre/post-ops on the device (N3)
    DevBuf region_read_start, block_region, read_mapq, ref_allele_index, qual_orig, values, keep;
    DevBuf meta;            // every metadata array of the staged chunk, one H2D copy; the DevBufs above are views into it
    int64_t n_blocks = 0;   // (region, sample) blocks of the staged chunk
    bool sampled = false;   // several samples per region: block_region is in use
    // ---- Smith-Waterman ----
    struct SwLaunch { int K; int32_t first; int32_t count; bool clamp; int duo; };  // duo: 0 one pair per warp, 1 packed (equal reference lengths), 2 packed (any)
    std::vector<SwLaunch> sw_launches;
    int sw_blocks = 0;
    int64_t sw_scratch_words = 0;
    int64_t sw_n_pairs = 0, sw_cells = 0, sw_packed_pairs = 0;
    int32_t sw_max_ops = 0, sw_strategy = 0;
    lk_sw_params sw_params = { 0, 0, 0, 0 };
    int64_t sw_ref_byte0 = 0, sw_alt_byte0 = 0;  // first arena bytes staged: device base pointers are biased by these
    std::vector<int32_t> sw_order_host;          // lives with the workspace so the queued H2D copy has a stable source
    DevBuf sw_refs, sw_alts, sw_ref_off, sw_alt_off, sw_bt, sw_bt_off, sw_ops, sw_nops, sw_offsets, sw_order;
};

// Region driver hook of the PairHMM chunk pipeline: modify_read_qualities before, normalise / filter after, on the device.
struct lk_engine_hook {
    const lk_engine_args *args;
    const uint8_t *read_mapq;          // per read of the whole batch (may be null when capping is disabled)
    const int32_t *ref_allele_index;   // per region of the whole batch, may be null
    double *values;                    // out, pair-indexed like `out` of lk_pairhmm_batch_host
    uint8_t *keep;                     // out, per read, may be null
    uint8_t pcr_cache[101];
};
lk_status lk_pairhmm_pipeline(lk_ctx *ctx, const lk_pairhmm_batch *batch, double *out, uint8_t *flags, const lk_engine_hook *hook);

lk_status lk_fail(lk_ctx *ctx, lk_status st, const std::string &msg);
lk_status lk_cuda_fail(lk_ctx *ctx, cudaError_t e, const char *what);

// best-allele search on likelihood matrices in device memory (lk_sw_api.cu); best_allele / informative are host arrays
namespace lk { struct BestAlleleBlock; }
lk_status lk_best_alleles_device(lk_ctx *ctx, const double *d_values, const lk::BestAlleleBlock *blocks, int64_t n_blocks,
                                 const int32_t *priorities, int64_t n_priorities, double informative_threshold, int64_t n_reads,
                                 int32_t *best_allele, uint8_t *informative);
