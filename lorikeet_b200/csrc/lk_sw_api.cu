// lk_sw_api.cu — Smith-Waterman entry points of the C ABI (include/lorikeet_b200.h).
//
// Replaces gkl::smithwaterman::align() as called from SmithWatermanAligner::align
// (src/smith_waterman/smith_waterman_aligner.rs:47-113); results are bit-identical to the reference's
// in-tree scalar path (calculate_matrix :124-263 + calculate_cigar :273-442), which the reference's own
// test_avx_mode asserts to equal the gkl path (tests/smith_waterman_aligner_unit_tests.rs:1001-1101).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "host_parallel.h"
#include "lk_api_internal.h"
#include "lk_internal.h"
#include "sw16_core.cuh"

using namespace lk;

namespace {

const int kSwKs[] = { LK_SW_KS };
constexpr int kNumSwKs = (int)(sizeof(kSwKs) / sizeof(kSwKs[0]));

// A/B measurements: LK_SW_DUO=0 in the environment keeps every pair on the one-pair-per-warp kernel, LK_SW_DUO=2 sends every
// packed launch to the any-lengths kernel
int sw_duo_mode() {
    static const int mode = [] { const char *e = getenv("LK_SW_DUO"); return (e && e[0] == '0') ? 0 : ((e && e[0] == '2') ? 2 : 1); }();
    return mode;
}
bool sw_duo_enabled() { return sw_duo_mode() != 0; }

int sw_choose_k(int n) {
    for (int i = 0; i < kNumSwKs; i++)
        if (n <= 32 * kSwKs[i]) return kSwKs[i];
    return kSwKs[kNumSwKs - 1];  // banded
}

template <class T>
cudaError_t up(DevBuf &b, const T *src, size_t n, cudaStream_t st) {
    cudaError_t e = b.reserve(std::max<size_t>(n * sizeof(T), 16));
    if (e != cudaSuccess || n == 0) return e;
    return cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
}

// Stage pairs [p0, p1) of the caller's arrays into workspace R on stream st: classify by K, order each class longest
// first, queue the H2D copies.  Asynchronous (no stream sync); R may be a reused workspace whose earlier work on st
// has completed.  Pair indices inside R are relative to p0; the byte arenas keep the caller's absolute offsets and
// the device base pointers are biased instead (sw_fill_args).
lk_status sw_stage_range(lk_ctx *ctx, lk_resident *R, int64_t p0, int64_t p1, const int64_t *ref_off, const uint8_t *refs,
                         const int64_t *alt_off, const uint8_t *alts, const lk_sw_params *params, lk_sw_overhang strategy,
                         int32_t max_ops, cudaStream_t st) {
    const int64_t n_pairs = p1 - p0;
    R->kind = 1;
    R->sw_n_pairs = n_pairs; R->sw_max_ops = max_ops; R->sw_strategy = (int32_t)strategy; R->sw_params = *params;
    R->sw_launches.clear();
    R->sw_ref_byte0 = n_pairs ? ref_off[p0] : 0;
    R->sw_alt_byte0 = n_pairs ? alt_off[p0] : 0;
    // one pass: K class, cell count and scratch need of every pair
    std::vector<int32_t> &order = R->sw_order_host;
    order.resize((size_t)n_pairs);
    std::vector<uint8_t> cls((size_t)n_pairs);
    std::vector<int64_t> cells((size_t)n_pairs);
    int64_t max_words = 64, nm_max[kNumSwKs], cells_max[kNumSwKs], count[kNumSwKs], m_max[kNumSwKs];
    for (int ki = 0; ki < kNumSwKs; ki++) { nm_max[ki] = 0; cells_max[ki] = 1; count[ki] = 0; m_max[ki] = 0; }
    R->sw_cells = 0;
    for (int64_t i = 0; i < n_pairs; i++) {
        const int64_t p = p0 + i;
        const int n = (int)(ref_off[p + 1] - ref_off[p]), m = (int)(alt_off[p + 1] - alt_off[p]);
        const int K = sw_choose_k(n);
        int ki = 0;
        while (kSwKs[ki] != K) ki++;
        cls[(size_t)i] = (uint8_t)ki;
        cells[(size_t)i] = (int64_t)n * m;
        R->sw_cells += cells[(size_t)i];
        count[ki]++;
        nm_max[ki] = std::max<int64_t>(nm_max[ki], (int64_t)n + m);
        cells_max[ki] = std::max(cells_max[ki], cells[(size_t)i]);
        m_max[ki] = std::max<int64_t>(m_max[ki], m);
        if (n > 0 && m > 0) max_words = std::max(max_words, sw_scratch_words(n, m, K));
    }
    // order array: one contiguous run per K class, longest pairs first inside a class (better tail balance of the
    // persistent grid), by a counting sort: the order only has to be roughly descending.  A class the packed kernel serves
    // (sw16_core.cuh: two consecutive pairs of the order per warp, best when their references have the same length) is ordered by
    // reference length first, then by alternate length in 16 steps, so that neighbours qualify; the others by cell count
    // in 256 steps.
    {
        const bool duo_weights = sw_duo_enabled() && sw16_weights_ok(params->match_value, params->mismatch_penalty,
                                                                     params->gap_open_penalty, params->gap_extend_penalty);
        bool duo[kNumSwKs];
        int nb[kNumSwKs], first_bucket[kNumSwKs + 1];
        first_bucket[0] = 0;
        for (int ki = 0; ki < kNumSwKs; ki++) {
            const bool clamp = !sw_clamp_inactive(nm_max[ki], params->match_value, params->mismatch_penalty, params->gap_open_penalty, params->gap_extend_penalty);
            duo[ki] = duo_weights && !clamp && sw_duo_has_k(kSwKs[ki]) && count[ki] > 1;
            nb[ki] = duo[ki] ? (32 * kSwKs[ki] + 2) * 16 : 256;
            first_bucket[ki + 1] = first_bucket[ki] + nb[ki];
        }
        std::vector<int32_t> slot((size_t)first_bucket[kNumSwKs] + 1, 0);
        auto bucket = [&](int64_t i) {
            const int ki = cls[(size_t)i];
            if (!duo[ki]) return first_bucket[ki] + (255 - (int)(cells[(size_t)i] * 255 / cells_max[ki]));
            const int64_t p = p0 + i;
            const int64_t n = ref_off[p + 1] - ref_off[p], m = alt_off[p + 1] - alt_off[p];
            const int rows = 32 * kSwKs[ki];
            const int n_slot = n > rows ? 0 : (int)(rows + 1 - n);  // banded references first, then longest first
            return first_bucket[ki] + n_slot * 16 + (15 - (int)(m * 15 / std::max<int64_t>(1, m_max[ki])));
        };
        for (int64_t i = 0; i < n_pairs; i++) slot[(size_t)bucket(i) + 1]++;
        for (size_t k = 1; k < slot.size(); k++) slot[k] += slot[k - 1];
        for (int64_t i = 0; i < n_pairs; i++) order[(size_t)slot[(size_t)bucket(i)]++] = (int32_t)i;
        int32_t first = 0;
        for (int ki = 0; ki < kNumSwKs; ki++) {
            if (count[ki] > 0) {
                const bool clamp = !sw_clamp_inactive(nm_max[ki], params->match_value, params->mismatch_penalty, params->gap_open_penalty, params->gap_extend_penalty);
                if (!duo[ki]) {
                    R->sw_launches.push_back(lk_resident::SwLaunch{ kSwKs[ki], first, (int32_t)count[ki], clamp, 0 });
                } else {
                    // The class is sorted by reference length.  Pairs that form duos of EQUAL length go first and run on the
                    // kernel built for that (the faster one); what is left -- the odd one of every length, and the banded
                    // references of the last class -- follows and runs on the any-lengths kernel.
                    const int rows = 32 * kSwKs[ki];
                    int32_t *run = order.data() + first;
                    std::vector<int32_t> rest;
                    int64_t w = 0;
                    for (int64_t a = 0; a < count[ki];) {
                        const int64_t pa = p0 + run[a], na = ref_off[pa + 1] - ref_off[pa];
                        int64_t b = a + 1;
                        while (b < count[ki] && ref_off[p0 + run[b] + 1] - ref_off[p0 + run[b]] == na) b++;
                        const int64_t len = b - a, take = (na <= rows && sw_duo_mode() == 1) ? (len & ~(int64_t)1) : 0;
                        for (int64_t q = 0; q < take; q++) run[w++] = run[a + q];  // w <= a: in place
                        for (int64_t q = take; q < len; q++) rest.push_back(run[a + q]);
                        a = b;
                    }
                    std::copy(rest.begin(), rest.end(), run + w);
                    // A short tail is not worth a launch of its own (it would start after the main one has drained and run on a
                    // handful of warps): up to eight left-over pairs ride along in the equal-lengths launch, where unequal neighbours
                    // take the one-pair routine; a few equal-length duos ride along in the any-lengths launch.
                    const int64_t few = std::max<int64_t>(64, count[ki] / 50);
                    if ((int64_t)rest.size() <= 8 && w > 0) R->sw_launches.push_back(lk_resident::SwLaunch{ kSwKs[ki], first, (int32_t)count[ki], clamp, 1 });
                    else if (w <= few) R->sw_launches.push_back(lk_resident::SwLaunch{ kSwKs[ki], first, (int32_t)count[ki], clamp, 2 });
                    else {
                        R->sw_launches.push_back(lk_resident::SwLaunch{ kSwKs[ki], first, (int32_t)w, clamp, 1 });
                        R->sw_launches.push_back(lk_resident::SwLaunch{ kSwKs[ki], first + (int32_t)w, (int32_t)rest.size(), clamp, 2 });
                    }
                    max_words = std::max(max_words, sw16_scratch_words((int)m_max[ki], kSwKs[ki]));
                }
            }
            first += (int32_t)count[ki];
        }
    }
    // persistent grid: up to 16 warps per SM, fewer if the per-warp backtrack scratch would exceed ~12 GB
    const int warps_per_block = LK_SW_THREADS / 32;
    int64_t warps = (int64_t)ctx->n_sm * 16;
    const int64_t budget_words = (12LL << 30) / 4;
    max_words = (max_words + 3) & ~(int64_t)3;  // uint4 plane stores need 16-byte aligned per-warp slices
    warps = std::max<int64_t>(warps_per_block, std::min(warps, budget_words / max_words));
    warps = std::min<int64_t>(warps, std::max<int64_t>(warps_per_block, ((n_pairs + warps_per_block - 1) / warps_per_block) * warps_per_block));
    R->sw_blocks = (int)(warps / warps_per_block);
    R->sw_scratch_words = max_words;
    const int64_t ref_bytes = n_pairs ? ref_off[p1] - R->sw_ref_byte0 : 0, alt_bytes = n_pairs ? alt_off[p1] - R->sw_alt_byte0 : 0;
    cudaError_t e;
    bool ok = (e = up(R->sw_refs, refs + R->sw_ref_byte0, (size_t)ref_bytes, st)) == cudaSuccess &&
              (e = up(R->sw_alts, alts + R->sw_alt_byte0, (size_t)alt_bytes, st)) == cudaSuccess &&
              (e = up(R->sw_ref_off, ref_off + p0, (size_t)n_pairs + 1, st)) == cudaSuccess &&
              (e = up(R->sw_alt_off, alt_off + p0, (size_t)n_pairs + 1, st)) == cudaSuccess &&
              (e = up(R->sw_order, order.data(), order.size(), st)) == cudaSuccess &&
              (e = R->sw_ops.reserve((size_t)std::max<int64_t>(n_pairs * max_ops * 4, 16))) == cudaSuccess &&
              (e = R->sw_nops.reserve((size_t)std::max<int64_t>(n_pairs * 4, 16))) == cudaSuccess &&
              (e = R->sw_offsets.reserve((size_t)std::max<int64_t>(n_pairs * 4, 16))) == cudaSuccess &&
              (e = R->counters.reserve(256)) == cudaSuccess &&
              (e = R->sw_bt.reserve((size_t)((int64_t)R->sw_blocks * warps_per_block * max_words * 4))) == cudaSuccess;
    if (!ok) return lk_cuda_fail(ctx, e, "SW upload");
    return LK_OK;
}

lk_status sw_validate(lk_ctx *ctx, int64_t n_pairs, const int64_t *ref_off, const uint8_t *refs, const int64_t *alt_off,
                      const uint8_t *alts, const lk_sw_params *params, lk_sw_overhang strategy, int32_t max_ops) {
    if (n_pairs < 0 || !params || max_ops < 1) return lk_fail(ctx, LK_ERR_INVALID, "bad SW arguments");
    if ((int)strategy < 0 || (int)strategy > 3) return lk_fail(ctx, LK_ERR_INVALID, "bad overhang strategy");
    if (n_pairs > 0 && (!ref_off || !refs || !alt_off || !alts)) return lk_fail(ctx, LK_ERR_INVALID, "null SW arrays");
    if (n_pairs > 0x7fffffffLL / std::max(1, max_ops)) return lk_fail(ctx, LK_ERR_INVALID, "SW batch too large");
    for (int64_t p = 0; p < n_pairs; p++)
        if (ref_off[p + 1] < ref_off[p] || alt_off[p + 1] < alt_off[p]) return lk_fail(ctx, LK_ERR_INVALID, "decreasing SW offsets");
    return LK_OK;
}

lk_status sw_upload_locked(lk_ctx *ctx, int64_t n_pairs, const int64_t *ref_off, const uint8_t *refs,
                           const int64_t *alt_off, const uint8_t *alts, const lk_sw_params *params,
                           lk_sw_overhang strategy, int32_t max_ops, lk_resident **out_res) {
    lk_status vs = sw_validate(ctx, n_pairs, ref_off, refs, alt_off, alts, params, strategy, max_ops);
    if (vs != LK_OK) return vs;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "cudaSetDevice");
    lk_resident *R = new (std::nothrow) lk_resident();
    if (!R) return lk_fail(ctx, LK_ERR_NOMEM, "out of host memory");
    lk_status st = sw_stage_range(ctx, R, 0, n_pairs, ref_off, refs, alt_off, alts, params, strategy, max_ops, ctx->stream[0]);
    if (st == LK_OK && (e = cudaStreamSynchronize(ctx->stream[0])) != cudaSuccess) st = lk_cuda_fail(ctx, e, "SW upload");
    if (st != LK_OK) { lk_resident_free(ctx, R); return st; }
    *out_res = R;
    return LK_OK;
}

// queue R's kernels on st (one launch per K class); no sync
lk_status sw_enqueue(lk_ctx *ctx, lk_resident *R, cudaStream_t st) {
    cudaError_t e;
    R->launch_count = 0;
    SwArgs A;
    memset(&A, 0, sizeof A);
    A.refs = (const uint8_t *)R->sw_refs.p - R->sw_ref_byte0; A.alts = (const uint8_t *)R->sw_alts.p - R->sw_alt_byte0;
    A.ref_off = (const int64_t *)R->sw_ref_off.p; A.alt_off = (const int64_t *)R->sw_alt_off.p;
    A.w_match = R->sw_params.match_value; A.w_mismatch = R->sw_params.mismatch_penalty;
    A.w_open = R->sw_params.gap_open_penalty; A.w_extend = R->sw_params.gap_extend_penalty;
    A.strategy = R->sw_strategy; A.max_ops = R->sw_max_ops;
    A.cigar_ops = (uint32_t *)R->sw_ops.p; A.n_ops = (int32_t *)R->sw_nops.p; A.offsets = (int32_t *)R->sw_offsets.p;
    A.scratch = (uint32_t *)R->sw_bt.p; A.scratch_words = R->sw_scratch_words;
    A.one = 1u; A.mone = 0xffffffffu; A.two16 = 65536u;
    A.packed_count = (int32_t *)R->counters.p + 15;
    A.prof = (unsigned long long *)((char *)R->counters.p + 64);
    if ((e = cudaMemsetAsync(R->counters.p, 0, 256, st)) != cudaSuccess) return lk_cuda_fail(ctx, e, "memset");
    int li = 0;
    for (const auto &L : R->sw_launches) {
        A.order = (const int32_t *)R->sw_order.p + L.first;
        A.n_order = L.count;
        A.cursor = (int32_t *)R->counters.p + li++;
        // the packed kernel runs 12 warps per SM (its two weight tables per warp), the one-pair kernel 16
        e = L.duo ? launch_sw_duo(L.K, L.duo == 1, A, ctx->n_sm, (int64_t)R->sw_blocks * (LK_SW_THREADS / 32), st) : launch_sw(L.K, L.clamp, A, R->sw_blocks, st);
        if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "sw_kernel");
        R->launch_count++;
    }
    return LK_OK;
}

lk_status sw_run_locked(lk_ctx *ctx, lk_resident *R, float *kernel_ms) {
    if (!R || R->kind != 1) return lk_fail(ctx, LK_ERR_INVALID, "not a Smith-Waterman resident batch");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "cudaSetDevice");
    cudaStream_t st = ctx->stream[0];
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (kernel_ms) { cudaEventCreate(&ev0); cudaEventCreate(&ev1); cudaEventRecord(ev0, st); }
    lk_status rs = sw_enqueue(ctx, R, st);
    if (kernel_ms) cudaEventRecord(ev1, st);
    e = cudaStreamSynchronize(st);
    if (kernel_ms) {
        if (e == cudaSuccess) cudaEventElapsedTime(kernel_ms, ev0, ev1);
        cudaEventDestroy(ev0); cudaEventDestroy(ev1);
    }
    if (rs != LK_OK) return rs;
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "sw kernels");
    int32_t packed = 0;
    if (cudaMemcpy(&packed, (const int32_t *)R->counters.p + 15, 4, cudaMemcpyDeviceToHost) == cudaSuccess) R->sw_packed_pairs = packed;
#ifdef LK_SW16_PROFILE
    {
        unsigned long long t[8] = { 0 };
        cudaMemcpy(t, (const char *)R->counters.p + 64, sizeof t, cudaMemcpyDeviceToHost);
        double tot = 0;
        for (int i = 0; i < 8; i++) tot += (double)t[i];
        fprintf(stderr, "sw16 phases (%% of warp time): classes %.1f  tables %.1f  fill %.1f  column+bottom %.1f  traceback %.1f  prologue %.1f  fallback %.1f\n",
                100 * t[0] / tot, 100 * t[1] / tot, 100 * t[2] / tot, 100 * t[3] / tot, 100 * t[4] / tot, 100 * t[5] / tot, 100 * t[6] / tot);
    }
#endif
    return LK_OK;
}

lk_status sw_download_locked(lk_ctx *ctx, lk_resident *R, uint32_t *cigar_ops, int32_t *n_ops, int32_t *offsets) {
    if (!R || R->kind != 1) return lk_fail(ctx, LK_ERR_INVALID, "not a Smith-Waterman resident batch");
    if (R->sw_n_pairs == 0) return LK_OK;
    if (!cigar_ops || !n_ops || !offsets) return lk_fail(ctx, LK_ERR_INVALID, "null SW outputs");
    cudaError_t e = cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream[0];
    if (e == cudaSuccess) e = cudaMemcpyAsync(cigar_ops, R->sw_ops.p, (size_t)R->sw_n_pairs * R->sw_max_ops * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(n_ops, R->sw_nops.p, (size_t)R->sw_n_pairs * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(offsets, R->sw_offsets.p, (size_t)R->sw_n_pairs * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "SW download");
    return LK_OK;
}

}  // namespace

extern "C" int64_t lk_resident_sw_packed_pairs(const lk_resident *r) { return (r && r->kind == 1) ? r->sw_packed_pairs : -1; }

extern "C" lk_status lk_sw_upload(lk_ctx *ctx, int64_t n_pairs, const int64_t *ref_off, const uint8_t *refs,
                                  const int64_t *alt_off, const uint8_t *alts, const lk_sw_params *params,
                                  lk_sw_overhang strategy, int32_t max_ops, lk_resident **out_res) {
    LK_ABI_GUARD_BEGIN
    if (!ctx || !out_res) return LK_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    *out_res = nullptr;
    return sw_upload_locked(ctx, n_pairs, ref_off, refs, alt_off, alts, params, strategy, max_ops, out_res);
    LK_ABI_GUARD_END(return lk_fail(ctx, LK_ERR_NOMEM, "out of host memory"), return lk_fail(ctx, LK_ERR_INVALID, "internal error"))
}

extern "C" lk_status lk_sw_run_resident(lk_ctx *ctx, lk_resident *res, float *kernel_ms) {
    if (!ctx) return LK_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    return sw_run_locked(ctx, res, kernel_ms);
}

extern "C" lk_status lk_sw_download(lk_ctx *ctx, lk_resident *res, uint32_t *cigar_ops, int32_t *n_ops, int32_t *offsets) {
    if (!ctx) return LK_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    return sw_download_locked(ctx, res, cigar_ops, n_ops, offsets);
}

// Host-buffer entry point.  Pairs are cut into chunks of contiguous pair ranges (by cell count, with a short first
// chunk so the GPU starts early) that flow through LK_PIPE_DEPTH context-owned, grow-only workspaces on as many
// streams: host-side ordering and H2D of chunk i+1 and the D2H of chunk i-1 overlap the kernels of chunk i, and no
// device memory is allocated or freed per call once the workspaces have grown to the batch shape.
extern "C" lk_status lk_sw_batch_host(lk_ctx *ctx, int64_t n_pairs, const int64_t *ref_off, const uint8_t *refs,
                                      const int64_t *alt_off, const uint8_t *alts, const lk_sw_params *params,
                                      lk_sw_overhang strategy, int32_t max_ops, uint32_t *cigar_ops, int32_t *n_ops,
                                      int32_t *alignment_offset) {
    LK_ABI_GUARD_BEGIN
    if (!ctx) return LK_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    lk_status st = sw_validate(ctx, n_pairs, ref_off, refs, alt_off, alts, params, strategy, max_ops);
    if (st != LK_OK) return st;
    ctx->last_launch_count = 0;
    if (n_pairs == 0) return LK_OK;
    if (!cigar_ops || !n_ops || !alignment_offset) return lk_fail(ctx, LK_ERR_INVALID, "null SW outputs");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "cudaSetDevice");
    // chunk boundaries by cells: ramp 1e9, 3e9, ... up to the steady size max(total / 8, 5e9)
    std::vector<int64_t> cut(1, 0);
    {
        double total = 0;
        for (int64_t p = 0; p < n_pairs; p++) total += (double)(ref_off[p + 1] - ref_off[p]) * (double)(alt_off[p + 1] - alt_off[p]);
        const double steady = std::max(total / 8.0, 5e9);
        double size = 1e9, acc = 0, next = size;
        for (int64_t p = 0; p < n_pairs; p++) {
            acc += (double)(ref_off[p + 1] - ref_off[p]) * (double)(alt_off[p + 1] - alt_off[p]);
            if (p + 1 == n_pairs) break;
            if (acc >= next && total - acc > 0.5 * size) {
                cut.push_back(p + 1);
                size = std::min(size * 3.0, steady);
                next = acc + size;
            }
        }
        cut.push_back(n_pairs);
    }
    const int C = (int)cut.size() - 1;
    int32_t launches = 0;
    for (int i = 0; i < C && st == LK_OK; i++) {
        const int w = i % LK_PIPE_DEPTH;
        if (!ctx->sw_ws[w]) { ctx->sw_ws[w] = new (std::nothrow) lk_resident(); if (!ctx->sw_ws[w]) return lk_fail(ctx, LK_ERR_NOMEM, "out of host memory"); }
        lk_resident *R = ctx->sw_ws[w];
        cudaStream_t s = ctx->stream[w];
        if ((e = cudaStreamSynchronize(s)) != cudaSuccess) { st = lk_cuda_fail(ctx, e, "SW pipeline"); break; }  // chunk i-3 done: R is free
        const int64_t p0 = cut[(size_t)i], p1 = cut[(size_t)i + 1];
        st = sw_stage_range(ctx, R, p0, p1, ref_off, refs, alt_off, alts, params, strategy, max_ops, s);
        if (st == LK_OK) st = sw_enqueue(ctx, R, s);
        if (st != LK_OK) break;
        launches += R->launch_count;
        e = cudaMemcpyAsync(cigar_ops + p0 * max_ops, R->sw_ops.p, (size_t)(p1 - p0) * max_ops * 4, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(n_ops + p0, R->sw_nops.p, (size_t)(p1 - p0) * 4, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(alignment_offset + p0, R->sw_offsets.p, (size_t)(p1 - p0) * 4, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) st = lk_cuda_fail(ctx, e, "SW download");
    }
    for (int i = 0; i < LK_PIPE_DEPTH; i++) {
        e = cudaStreamSynchronize(ctx->stream[i]);
        if (e != cudaSuccess && st == LK_OK) st = lk_cuda_fail(ctx, e, "SW pipeline");
    }
    ctx->last_launch_count = launches;
    return st;
    LK_ABI_GUARD_END(return lk_fail(ctx, LK_ERR_NOMEM, "out of host memory"), return lk_fail(ctx, LK_ERR_INVALID, "internal error"))
}

extern "C" lk_status lk_sw_align(lk_ctx *ctx, const uint8_t *ref, int64_t ref_len, const uint8_t *alt, int64_t alt_len,
                                 const lk_sw_params *params, lk_sw_overhang strategy, char *cigar_text, size_t cigar_cap,
                                 int32_t *alignment_offset) {
    if (!ctx) return LK_ERR_INVALID;
    if (!ref || !alt || !params || !cigar_text || !alignment_offset || cigar_cap < 2) return lk_fail(ctx, LK_ERR_INVALID, "null argument");
    if (ref_len <= 0 || alt_len <= 0)  // smith_waterman_aligner.rs:65-68
        return lk_fail(ctx, LK_ERR_INVALID, "non-empty sequences are required for the Smith-Waterman calculation");
    const int64_t ro[2] = { 0, ref_len }, ao[2] = { 0, alt_len };
    const int32_t max_ops = (int32_t)std::min<int64_t>(2 * (ref_len + alt_len) + 8, 1 << 20);
    std::vector<uint32_t> ops((size_t)max_ops);
    int32_t n_ops = 0;
    lk_status st = lk_sw_batch_host(ctx, 1, ro, ref, ao, alt, params, strategy, max_ops, ops.data(), &n_ops, alignment_offset);
    if (st != LK_OK) return st;
    if (n_ops < 0) return lk_fail(ctx, n_ops == -2 ? LK_ERR_CAPACITY : LK_ERR_INVALID, "Smith-Waterman failed");
    std::string text;
    static const char codes[] = "MIDNSHP=X";
    for (int i = 0; i < n_ops; i++) {
        char buf[32];
        snprintf(buf, sizeof buf, "%u%c", ops[(size_t)i] >> 4, codes[ops[(size_t)i] & 0xF]);
        text += buf;
    }
    if (text.size() + 1 > cigar_cap) return lk_fail(ctx, LK_ERR_CAPACITY, "cigar_text too small");
    memcpy(cigar_text, text.c_str(), text.size() + 1);
    return LK_OK;
}

// ---- best haplotype per read on matrices that stayed in HBM (N2 without the round trip of `values` through the host) ----
#include "best_allele.h"
lk_status lk_best_alleles_device(lk_ctx *ctx, const double *d_values, const lk::BestAlleleBlock *blocks, int64_t n_blocks,
                                 const int32_t *priorities, int64_t n_priorities, double informative_threshold, int64_t n_reads,
                                 int32_t *best_allele, uint8_t *informative) {
    if (!ctx) return LK_ERR_INVALID;
    if (n_blocks == 0 || n_reads == 0) return LK_OK;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "cudaSetDevice");
    cudaStream_t st = ctx->stream[0];
    DevBuf &d_blocks = ctx->d_ba_blocks, &d_prio = ctx->d_ba_prio, &d_best = ctx->d_ba_best, &d_info = ctx->d_ba_info;  // kept across calls
    auto done = [&](lk_status r) { return r; };
    if ((e = d_blocks.reserve((size_t)n_blocks * sizeof(lk::BestAlleleBlock))) != cudaSuccess || (e = d_prio.reserve((size_t)std::max<int64_t>(n_priorities, 1) * 4)) != cudaSuccess ||
        (e = d_best.reserve((size_t)n_reads * 4)) != cudaSuccess || (e = d_info.reserve((size_t)n_reads)) != cudaSuccess)
        return done(lk_cuda_fail(ctx, e, "alloc best-allele buffers"));
    if ((e = cudaMemcpyAsync(d_blocks.p, blocks, (size_t)n_blocks * sizeof(lk::BestAlleleBlock), cudaMemcpyHostToDevice, st)) != cudaSuccess ||
        (n_priorities > 0 && (e = cudaMemcpyAsync(d_prio.p, priorities, (size_t)n_priorities * 4, cudaMemcpyHostToDevice, st)) != cudaSuccess) ||
        (e = cudaMemsetAsync(d_best.p, 0xff, (size_t)n_reads * 4, st)) != cudaSuccess || (e = cudaMemsetAsync(d_info.p, 0, (size_t)n_reads, st)) != cudaSuccess)
        return done(lk_cuda_fail(ctx, e, "H2D best-allele inputs"));
    if ((e = lk::launch_best_alleles(d_values, (const lk::BestAlleleBlock *)d_blocks.p, n_blocks, (const int32_t *)d_prio.p, informative_threshold,
                                     (int32_t *)d_best.p, (uint8_t *)d_info.p, st)) != cudaSuccess)
        return done(lk_cuda_fail(ctx, e, "best_alleles_kernel"));
    if ((e = cudaMemcpyAsync(best_allele, d_best.p, (size_t)n_reads * 4, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(informative, d_info.p, (size_t)n_reads, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(st)) != cudaSuccess)
        return done(lk_cuda_fail(ctx, e, "best-allele D2H"));
    return done(LK_OK);
}
