// lk_api.cu — the C ABI declared in include/lorikeet_b200.h: context, staging, task building,
// launches.  All compute is on the GPU; this file only moves bytes and builds work lists.
//
// Reference call sites this layer replaces:
//   PairHMM::compute_likelihoods            src/pair_hmm/pair_hmm.rs:345-375  (gkl forward per pair)
//   PairHMM::compute_log10_likelihoods      src/pair_hmm/pair_hmm.rs:217-341  (result layout r*A + h)
//   SmithWatermanAligner::align             src/smith_waterman/smith_waterman_aligner.rs:47-113
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lorikeet_b200.h"
#include "host_model.h"
#include "host_parallel.h"
#include "host_prep.h"
#include "lk_api_internal.h"
#include "lk_internal.h"
#include "pairhmm_core.cuh"

using namespace lk;

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
lk_status lk_fail(lk_ctx *ctx, lk_status st, const std::string &msg) {
    if (ctx) ctx->err = msg;
    return st;
}

lk_status lk_cuda_fail(lk_ctx *ctx, cudaError_t e, const char *what) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    (void)cudaGetLastError();
    return lk_fail(ctx, e == cudaErrorMemoryAllocation ? LK_ERR_NOMEM : LK_ERR_CUDA, buf);
}

cudaError_t DevBuf::reserve(size_t bytes) {
    if (view) { p = nullptr; cap = 0; view = false; }
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFree(p); p = nullptr; cap = 0; }
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) { p = nullptr; return e; }
    cap = want;
    return cudaSuccess;
}
void DevBuf::release() {
    if (p && !view) cudaFree(p);
    p = nullptr; cap = 0; view = false;
}

template <class T>
static cudaError_t upload(DevBuf &b, const T *src, size_t n, cudaStream_t st) {
    cudaError_t e = b.reserve(std::max<size_t>(n * sizeof(T), 16));
    if (e != cudaSuccess) return e;
    if (n == 0) return cudaSuccess;
    return cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" const char *lk_version(void) { return "lorikeet_b200 0.1.0 (sm_100a)"; }

extern "C" lk_status lk_create(int device, lk_ctx **out_ctx) {
    if (!out_ctx) return LK_ERR_INVALID;
    *out_ctx = nullptr;
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0) {
        (void)cudaGetLastError();
        return LK_ERR_CUDA;  // no CPU fallback by design
    }
    if (device < 0) {
        e = cudaGetDevice(&device);
        if (e != cudaSuccess) return LK_ERR_CUDA;
    }
    if (device >= n_dev) return LK_ERR_INVALID;
    lk_ctx *ctx = new (std::nothrow) lk_ctx();
    if (!ctx) return LK_ERR_NOMEM;
    ctx->device = device;
    if ((e = cudaSetDevice(device)) != cudaSuccess) { delete ctx; return LK_ERR_CUDA; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { delete ctx; return LK_ERR_CUDA; }
    ctx->n_sm = prop.multiProcessorCount;
    for (int i = 0; i < 3; i++)
        if ((e = cudaStreamCreateWithFlags(&ctx->stream[i], cudaStreamNonBlocking)) != cudaSuccess) { delete ctx; return LK_ERR_CUDA; }
    ModelTables t;
    build_model_tables(t);
    cudaStream_t st = ctx->stream[0];
    if ((e = upload(ctx->d_qerr_f, t.qerr_f.data(), t.qerr_f.size(), st)) != cudaSuccess ||
        (e = upload(ctx->d_qerr_d, t.qerr_d.data(), t.qerr_d.size(), st)) != cudaSuccess ||
        (e = upload(ctx->d_m2m, t.m2m.data(), t.m2m.size(), st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(st)) != cudaSuccess) {
        lk_destroy(ctx);
        return LK_ERR_CUDA;
    }
    const char *fk = getenv("LK_FORCE_K");
    ctx->force_k = fk ? atoi(fk) : 0;
    // launches with fewer warp tasks than two rounds of the resident warps (3 blocks x 4 warps per SM) are cut finer
    const char *sp = getenv("LK_SPLIT_TARGET");
    ctx->split_target_tasks = sp ? atoll(sp) : (int64_t)ctx->n_sm * 12 * 2;
    *out_ctx = ctx;
    return LK_OK;
}

extern "C" void lk_destroy(lk_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < 3; i++)
        if (ctx->stream[i]) { cudaStreamSynchronize(ctx->stream[i]); cudaStreamDestroy(ctx->stream[i]); }
    for (int i = 0; i < LK_PIPE_DEPTH; i++)
        if (ctx->ws[i]) { lk_resident_free(nullptr, ctx->ws[i]); ctx->ws[i] = nullptr; }
    for (int i = 0; i < LK_PIPE_DEPTH; i++)
        if (ctx->sw_ws[i]) { lk_resident_free(nullptr, ctx->sw_ws[i]); ctx->sw_ws[i] = nullptr; }
    ctx->d_qerr_f.release(); ctx->d_qerr_d.release(); ctx->d_m2m.release();
    ctx->d_ba_blocks.release(); ctx->d_ba_prio.release(); ctx->d_ba_best.release(); ctx->d_ba_info.release();
    ctx->stage_refs.release(); ctx->stage_alts.release(); ctx->stage_ops.release(); ctx->stage_n.release(); ctx->stage_off.release();
    delete ctx;
}

extern "C" const char *lk_last_error(const lk_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" lk_status lk_set_tristate_correction(lk_ctx *ctx, int use_tristate) {
    if (!ctx) return LK_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->tristate = use_tristate != 0;
    return LK_OK;
}

extern "C" lk_status lk_set_force_f64(lk_ctx *ctx, int force_f64) {
    if (!ctx) return LK_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    ctx->force_f64 = force_f64 != 0;
    return LK_OK;
}

// ------------------------------------------------------------------------------------------------
// PairHMM: batch validation, K choice, task building
// ------------------------------------------------------------------------------------------------
static lk_status validate_batch(lk_ctx *ctx, const lk_pairhmm_batch *b, int64_t *n_reads, int64_t *n_haps) {
    if (!b) return lk_fail(ctx, LK_ERR_INVALID, "null batch");
    if (b->n_regions < 0) return lk_fail(ctx, LK_ERR_INVALID, "negative n_regions");
    if (!b->region_read_start || !b->region_hap_start) return lk_fail(ctx, LK_ERR_INVALID, "null region offsets");
    if (b->region_read_start[0] != 0 || b->region_hap_start[0] != 0)
        return lk_fail(ctx, LK_ERR_INVALID, "region offsets must start at 0");
    for (int64_t g = 0; g < b->n_regions; g++)
        if (b->region_read_start[g + 1] < b->region_read_start[g] || b->region_hap_start[g + 1] < b->region_hap_start[g])
            return lk_fail(ctx, LK_ERR_INVALID, "region offsets must be non-decreasing");
    *n_reads = b->region_read_start[b->n_regions];
    *n_haps = b->region_hap_start[b->n_regions];
    if (*n_reads > 0x7fffffffLL || *n_haps > 0x7fffffffLL) return lk_fail(ctx, LK_ERR_INVALID, "batch too large (2^31 reads/haplotypes)");
    if (*n_reads > 0) {
        if (!b->read_off || !b->read_bases || !b->read_qual || !b->read_ins || !b->read_del)
            return lk_fail(ctx, LK_ERR_INVALID, "null read arrays");
        for (int64_t r = 0; r < *n_reads; r++)
            if (b->read_off[r + 1] <= b->read_off[r])  // pair_hmm.rs:417-440 asserts equal, non-degenerate lengths
                return lk_fail(ctx, LK_ERR_INVALID, "empty read or decreasing read_off");
    }
    if (*n_haps > 0) {
        if (!b->hap_off || !b->hap_bases) return lk_fail(ctx, LK_ERR_INVALID, "null haplotype arrays");
        for (int64_t h = 0; h < *n_haps; h++)
            if (b->hap_off[h + 1] <= b->hap_off[h]) return lk_fail(ctx, LK_ERR_INVALID, "empty haplotype or decreasing hap_off");
    }
    return LK_OK;
}

extern "C" lk_status lk_pairhmm_batch_dims(const lk_pairhmm_batch *b, int64_t *n_pairs, int64_t *n_cells) {
    int64_t nr, nh;
    lk_status st = validate_batch(nullptr, b, &nr, &nh);
    if (st != LK_OK) return st;
    int64_t pairs = 0, cells = 0;
    for (int64_t g = 0; g < b->n_regions; g++) {
        const int64_t r0 = b->region_read_start[g], r1 = b->region_read_start[g + 1];
        const int64_t h0 = b->region_hap_start[g], h1 = b->region_hap_start[g + 1];
        pairs += (r1 - r0) * (h1 - h0);
        if (r1 > r0 && h1 > h0) cells += (b->read_off[r1] - b->read_off[r0]) * (b->hap_off[h1] - b->hap_off[h0]);
    }
    if (n_pairs) *n_pairs = pairs;
    if (n_cells) *n_cells = cells;
    return LK_OK;
}

// ------------------------------------------------------------------------------------------------
// PairHMM: staging a region range into a workspace, launching, collecting
// ------------------------------------------------------------------------------------------------
static bool gcp_is_uniform(const uint8_t *gcp, int64_t n, uint8_t *value) {
    if (n == 0) { *value = 10; return true; }
    const uint8_t v = gcp[0];
    for (int64_t i = 1; i < n; i++)
        if (gcp[i] != v) return false;
    *value = v;
    return true;
}

cudaError_t PinnedBuf::reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
    const size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
    if (e != cudaSuccess) { p = nullptr; return e; }
    cap = want;
    return cudaSuccess;
}
void PinnedBuf::release() {
    if (p) cudaFreeHost(p);
    p = nullptr; cap = 0;
}

struct GcpMode { bool vargcp; uint8_t constant; const uint8_t *array; };

static GcpMode resolve_gcp(const lk_pairhmm_batch *b, int64_t n_reads) {
    GcpMode m{ false, b->gcp_constant, nullptr };
    if (b->read_gcp && n_reads > 0) {
        uint8_t v;
        const int64_t lo = b->read_off[0], hi = b->read_off[n_reads];
        if (gcp_is_uniform(b->read_gcp + lo, hi - lo, &v)) m.constant = v;
        else { m.vargcp = true; m.array = b->read_gcp; }
    }
    return m;
}

// Stage regions [g0, g1) of a validated batch into workspace R: build the work lists, copy the byte arenas and the
// metadata to the device on stream st.  Asynchronous: returns once the copies are queued (metadata goes through
// the workspace's pinned arena; `ev_staged` marks the point where that arena may be reused).
static lk_status stage_range(lk_ctx *ctx, const lk_pairhmm_batch *b, int64_t g0, int64_t g1, const GcpMode &gcp,
                             lk_resident *R, cudaStream_t st, const lk_engine_hook *hook = nullptr) {
    cudaError_t e;
    if (R->ev_staged) {
        if ((e = cudaEventSynchronize(R->ev_staged)) != cudaSuccess) return lk_cuda_fail(ctx, e, "wait staging");
    } else if ((e = cudaEventCreateWithFlags(&R->ev_staged, cudaEventDisableTiming)) != cudaSuccess) {
        return lk_cuda_fail(ctx, e, "event create");
    }
    HostPrep &P = R->prep;
    // the haplotype-axis split of small launches exists for the engine's default gcp only (launch_fast_k)
    build_prep(b, g0, g1, ctx->force_k, !ctx->force_f64, P, (!gcp.vargcp && gcp.constant == 10) ? ctx->split_target_tasks : 0);
    const int64_t rbase = b->region_read_start[g0], hbase = b->region_hap_start[g0];
    const int64_t n_reads = b->region_read_start[g1] - rbase, n_haps = b->region_hap_start[g1] - hbase;
    const int64_t n_regions = g1 - g0;
    if (P.n_pairs > 0x7fffffffLL) return lk_fail(ctx, LK_ERR_INVALID, "batch too large (2^31 pairs per chunk)");
    R->kind = 0;
    R->n_regions = n_regions; R->n_reads = n_reads; R->n_haps = n_haps;
    R->n_pairs = P.n_pairs; R->n_cells = P.n_cells; R->h_max = P.h_max;
    R->launches = P.launches;
    R->all_f64 = ctx->force_f64;
    R->vargcp = gcp.vargcp; R->gcp_constant = gcp.constant;
    // reads too long for the fast kernel: fp32 band kernel, or straight to the f64 pass when gcp varies per base
    const bool use_band = !ctx->force_f64 && !gcp.vargcp;
    R->n_band = use_band ? (int32_t)P.direct.size() : 0;
    R->n_direct = (ctx->force_f64 || use_band) ? 0 : (int32_t)P.direct.size();
    R->read_byte0 = n_reads ? b->read_off[rbase] : 0;
    R->hap_byte0 = n_haps ? b->hap_off[hbase] : 0;
    const int64_t read_bytes = n_reads ? b->read_off[rbase + n_reads] - R->read_byte0 : 0;
    const int64_t hap_bytes = n_haps ? b->hap_off[hbase + n_haps] - R->hap_byte0 : 0;
    R->read_bytes_staged = read_bytes;

    // ---- metadata through the pinned arena ----
    struct Piece { DevBuf *dst; const void *src; size_t bytes; };
    std::vector<int64_t> local_rhs((size_t)n_regions + 1);
    for (int64_t g = 0; g <= n_regions; g++) local_rhs[(size_t)g] = b->region_hap_start[g0 + g] - hbase;
    // (region, sample) blocks of the engine epilogue: one per region, or one per sample of every region
    std::vector<int64_t> local_rrs;
    std::vector<int32_t> local_breg;
    const bool sampled = hook && b->region_sample_start && b->sample_read_start;
    if (hook) {
        if (sampled) {
            for (int64_t g = g0; g < g1; g++)
                for (int64_t sm = b->region_sample_start[g]; sm < b->region_sample_start[g + 1]; sm++) {
                    local_rrs.push_back(b->sample_read_start[sm] - rbase);
                    local_breg.push_back((int32_t)(g - g0));
                }
            local_rrs.push_back(n_reads);
        } else {
            local_rrs.resize((size_t)n_regions + 1);
            for (int64_t g = 0; g <= n_regions; g++) local_rrs[(size_t)g] = b->region_read_start[g0 + g] - rbase;
        }
        R->n_blocks = (int64_t)local_rrs.size() - 1;
        R->sampled = sampled;
    }
    const Piece pieces[] = {
        { &R->region_read_start, hook ? local_rrs.data() : nullptr, hook ? local_rrs.size() * 8 : 0 },
        { &R->block_region, sampled ? local_breg.data() : nullptr, sampled ? local_breg.size() * 4 : 0 },
        { &R->read_mapq, hook && hook->read_mapq ? hook->read_mapq + rbase : nullptr, hook && hook->read_mapq ? (size_t)n_reads : 0 },
        { &R->ref_allele_index, hook && hook->ref_allele_index ? hook->ref_allele_index + g0 : nullptr,
          hook && hook->ref_allele_index ? (size_t)n_regions * 4 : 0 },
        { &R->read_off, b->read_off + rbase, (size_t)(n_reads + 1) * 8 },
        { &R->read_region, P.read_region.data(), (size_t)n_reads * 4 },
        { &R->read_out_base, P.read_out_base.data(), (size_t)n_reads * 8 },
        { &R->region_hap_start, local_rhs.data(), (size_t)(n_regions + 1) * 8 },
        { &R->hap_off, b->hap_off + hbase, (size_t)(n_haps + 1) * 8 },
        { &R->henc_off, P.henc_off.data(), (size_t)n_haps * 8 },
        { &R->tasks, P.tasks.data(), P.tasks.size() * sizeof(WarpTask) },
        { &R->task_haps, P.task_haps.data(), P.task_haps.size() * sizeof(TaskHaps) },
        { &R->groups, P.groups.data(), P.groups.size() * sizeof(TaskGroup) },
        { &R->rescue_pairs, P.direct.data(), (size_t)R->n_direct * sizeof(RescuePair) },
        { &R->band_pairs, P.direct.data(), (size_t)R->n_band * sizeof(RescuePair) },
    };
    size_t total = 0;
    for (const Piece &pc : pieces) total += (pc.bytes + 255) & ~(size_t)255;
    if ((e = R->pinned.reserve(total + 256)) != cudaSuccess) return lk_cuda_fail(ctx, e, "pinned staging");
    // device buffers (grow-only; reused across calls)
#define RS(buf, bytes) if ((e = R->buf.reserve((size_t)std::max<int64_t>((int64_t)(bytes), 16))) != cudaSuccess) return lk_cuda_fail(ctx, e, "alloc " #buf);
    RS(read_bases, read_bytes); RS(read_qual, read_bytes); RS(read_ins, read_bytes); RS(read_del, read_bytes);
    if (gcp.vargcp) { RS(read_gcp, read_bytes); }
    RS(hap_bases, hap_bytes);
    RS(read_flags, n_reads); RS(hap_flags, n_haps);
    RS(hidx, P.henc_words);
    RS(out, P.n_pairs * 8); RS(flags, P.n_pairs);
    if (hook) { RS(values, P.n_pairs * 8); RS(keep, n_reads); if (hook->args->dynamic_read_disqualification) { RS(qual_orig, read_bytes); } }
    RS(counters, 64);
    RS(rescue_pairs, std::max<int64_t>(P.n_pairs, 1) * (int64_t)sizeof(RescuePair));
    R->plain_blocks = ctx->n_sm * 2;
    R->scratch_stride = 6 * ((int64_t)P.h_max + 1);
    RS(scratch, (int64_t)R->plain_blocks * (LK_PLAIN_THREADS / 32) * R->scratch_stride * 8);
    R->band_blocks = ctx->n_sm * LK_BAND_MINB;
    R->band_stride = 2 * ((int64_t)P.h_max + 1);
    if (R->n_band > 0) { RS(band_scratch, (int64_t)R->band_blocks * (LK_BAND_THREADS / 32) * R->band_stride * 16); }
    // All metadata arrays travel as ONE H2D copy of the pinned arena into one device block with the same layout (256-byte
    // aligned pieces); the per-array DevBufs become views into it.  (One copy instead of fourteen: a small launch of the
    // batcher is latency-bound, and every cudaMemcpyAsync costs a few microseconds on both sides.)  rescue_pairs is the
    // exception: the kernels append to it, so it keeps its own, larger buffer.
    if ((e = R->meta.reserve(total + 256)) != cudaSuccess) return lk_cuda_fail(ctx, e, "alloc metadata");
    size_t off = 0;
    for (const Piece &pc : pieces) {
        if (pc.bytes) memcpy((char *)R->pinned.p + off, pc.src, pc.bytes);
        if (pc.dst == &R->rescue_pairs) {
            if (pc.bytes && (e = cudaMemcpyAsync(pc.dst->p, (char *)R->pinned.p + off, pc.bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess)
                return lk_cuda_fail(ctx, e, "metadata H2D");
        } else {
            pc.dst->set_view((char *)R->meta.p + off);
        }
        off += (pc.bytes + 255) & ~(size_t)255;
    }
    if (total && (e = cudaMemcpyAsync(R->meta.p, R->pinned.p, total, cudaMemcpyHostToDevice, st)) != cudaSuccess)
        return lk_cuda_fail(ctx, e, "metadata H2D");
#undef RS
    if ((e = cudaEventRecord(R->ev_staged, st)) != cudaSuccess) return lk_cuda_fail(ctx, e, "event record");
    // ---- byte arenas straight from the caller's memory ----
#define CP(buf, src) if (read_bytes && (e = cudaMemcpyAsync(R->buf.p, (src) + R->read_byte0, (size_t)read_bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) return lk_cuda_fail(ctx, e, "H2D " #buf);
    CP(read_bases, b->read_bases); CP(read_qual, b->read_qual); CP(read_ins, b->read_ins); CP(read_del, b->read_del);
    if (gcp.vargcp) { CP(read_gcp, gcp.array); }
#undef CP
    if (hap_bytes && (e = cudaMemcpyAsync(R->hap_bases.p, b->hap_bases + R->hap_byte0, (size_t)hap_bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess)
        return lk_cuda_fail(ctx, e, "H2D hap_bases");
    return LK_OK;
}

static void fill_args(lk_ctx *ctx, lk_resident *R, PairHmmArgs &A) {
    memset(&A, 0, sizeof A);
    // arenas are indexed with the caller's absolute byte offsets: bias the base pointers instead of rebasing offsets
    A.read_bases = (const uint8_t *)R->read_bases.p - R->read_byte0; A.read_qual = (const uint8_t *)R->read_qual.p - R->read_byte0;
    A.read_ins = (const uint8_t *)R->read_ins.p - R->read_byte0; A.read_del = (const uint8_t *)R->read_del.p - R->read_byte0;
    A.read_gcp = R->vargcp ? (const uint8_t *)R->read_gcp.p - R->read_byte0 : nullptr;
    A.read_off = (const int64_t *)R->read_off.p; A.read_region = (const int32_t *)R->read_region.p;
    A.read_out_base = (const int64_t *)R->read_out_base.p; A.read_flags = (const uint8_t *)R->read_flags.p;
    A.region_hap_start = (const int64_t *)R->region_hap_start.p; A.hap_off = (const int64_t *)R->hap_off.p;
    A.hap_bases = (const uint8_t *)R->hap_bases.p - R->hap_byte0; A.hidx = (const uint8_t *)R->hidx.p;
    A.henc_off = (const int64_t *)R->henc_off.p; A.hap_flags = (const uint8_t *)R->hap_flags.p;
    A.qerr_f = (const float *)ctx->d_qerr_f.p; A.qerr_d = (const double *)ctx->d_qerr_d.p; A.m2m = (const double *)ctx->d_m2m.p;
    A.groups = (const TaskGroup *)R->groups.p;
    A.out = (double *)R->out.p; A.flags = (uint8_t *)R->flags.p;
    A.rescue_count = (int32_t *)R->counters.p; A.rescue_cursor = (int32_t *)R->counters.p + 1;
    A.rescue_pairs = (RescuePair *)R->rescue_pairs.p; A.rescue_capacity = (int32_t)R->n_pairs;
    A.scratch = (double *)R->scratch.p; A.scratch_stride = R->scratch_stride;
    A.band_pairs = (const RescuePair *)R->band_pairs.p; A.n_band = R->n_band;
    A.band_cursor = (int32_t *)R->counters.p + 2;
    A.band_scratch = R->band_scratch.p; A.band_scratch_stride = R->band_stride;
    A.gcp_constant = R->gcp_constant;
    A.mismatch_div = ctx->tristate ? 3.0f : 1.0f;
    A.cf_uniform = (float)std::pow(10.0, (double)R->gcp_constant / -10.0);
}

// Queue every kernel of one staged workspace on st (no synchronisation).
static void fill_engine_args(const lk_engine_hook *hook, lk_resident *R, const PairHmmArgs &A, EngineDevArgs &E) {
    memset(&E, 0, sizeof E);
    const lk_engine_args *a = hook->args;
    E.read_bases = A.read_bases;
    E.read_qual = (uint8_t *)A.read_qual; E.read_ins = (uint8_t *)A.read_ins; E.read_del = (uint8_t *)A.read_del;
    E.qual_orig = a->dynamic_read_disqualification ? (const uint8_t *)R->qual_orig.p - R->read_byte0 : nullptr;
    E.read_off = A.read_off;
    E.read_mapq = hook->read_mapq ? (const uint8_t *)R->read_mapq.p : nullptr;
    E.read_out_base = A.read_out_base;
    E.block_read_start = (const int64_t *)R->region_read_start.p;
    E.block_region = R->sampled ? (const int32_t *)R->block_region.p : nullptr;
    E.region_hap_start = A.region_hap_start;
    E.ref_allele_index = hook->ref_allele_index ? (const int32_t *)R->ref_allele_index.p : nullptr;
    E.lk = A.out;
    E.values = (double *)R->values.p;
    E.keep = (uint8_t *)R->keep.p;
    E.cap_diff = a->log10_global_read_mismapping_rate;
    E.expected_error_rate = a->expected_error_rate_per_base;
    E.disq_scale = a->read_disqualification_scale;
    E.use_pcr = a->pcr_error_model != LK_PCR_NONE;
    E.cap_to_mapq = !a->disable_cap_read_qualities_to_mapq;
    E.symmetric = a->symmetrically_normalize_alleles_to_reference != 0;
    E.dynamic_disq = a->dynamic_read_disqualification != 0;
    E.bq_threshold = a->base_quality_score_threshold;
    memcpy(E.pcr_cache, hook->pcr_cache, sizeof E.pcr_cache);
}

static lk_status enqueue_kernels(lk_ctx *ctx, lk_resident *R, cudaStream_t st, bool timed, const lk_engine_hook *hook = nullptr) {
    cudaError_t e;
    PairHmmArgs A;
    fill_args(ctx, R, A);
    EngineDevArgs E;
    if (hook) {
        fill_engine_args(hook, R, A, E);
        if (R->n_reads > 0) {
            if (E.dynamic_disq) {  // thresholds use the ORIGINAL qualities: keep a copy before they are rewritten in place
                if ((e = cudaMemcpyAsync(R->qual_orig.p, R->read_qual.p, R->read_bytes_staged, cudaMemcpyDeviceToDevice, st)) != cudaSuccess)
                    return lk_cuda_fail(ctx, e, "copy original qualities");
            }
            if ((e = launch_engine_modify(E, R->n_reads, st)) != cudaSuccess) return lk_cuda_fail(ctx, e, "engine_modify_reads");
        }
    }
    R->counter_init[0] = R->all_f64 ? (int32_t)R->n_pairs : R->n_direct;
    R->counter_init[1] = 0;
    R->counter_init[2] = 0;
    if ((e = cudaMemcpyAsync(R->counters.p, R->counter_init, 12, cudaMemcpyHostToDevice, st)) != cudaSuccess)
        return lk_cuda_fail(ctx, e, "reset counters");
    R->launch_count = 0;
    if (R->n_pairs == 0) return LK_OK;
    if (R->all_f64) {
        if ((e = launch_list_all_pairs(A, R->n_reads, st)) != cudaSuccess) return lk_cuda_fail(ctx, e, "list_all_pairs");
        R->launch_count++;
    } else {
        if ((e = launch_encode_haps(A.hap_bases, A.hap_off, A.henc_off, R->n_haps, (uint8_t *)R->hidx.p, (uint8_t *)R->hap_flags.p, st)) != cudaSuccess)
            return lk_cuda_fail(ctx, e, "encode_haps");
        if ((e = launch_scan_reads(A.read_bases, A.read_qual, A.read_ins, A.read_del, A.read_off, R->n_reads, (uint8_t *)R->read_flags.p, st)) != cudaSuccess)
            return lk_cuda_fail(ctx, e, "scan_reads");
        R->launch_count += 2;
        if (timed) { cudaEventCreate(&R->ev_fast0); cudaEventCreate(&R->ev_fast1); cudaEventRecord(R->ev_fast0, st); }
        for (const auto &L : R->launches) {
            PairHmmArgs AL = A;
            AL.tasks = (const WarpTask *)R->tasks.p + L.first_task;
            AL.n_tasks = L.n_tasks;
            AL.task_haps = L.split ? (const TaskHaps *)R->task_haps.p + L.first_task : nullptr;
            if ((e = launch_fast(L.K, AL, R->vargcp, st)) != cudaSuccess) return lk_cuda_fail(ctx, e, "pairhmm_fast");
            R->launch_count++;
        }
        if (R->n_band > 0) {
            if ((e = launch_band(A, R->band_blocks, st)) != cudaSuccess) return lk_cuda_fail(ctx, e, "pairhmm_band");
            R->launch_count++;
        }
        if (timed) cudaEventRecord(R->ev_fast1, st);
    }
    if ((e = launch_plain(A, R->plain_blocks, st)) != cudaSuccess) return lk_cuda_fail(ctx, e, "pairhmm_plain");
    R->launch_count++;
    if (hook) {
        if ((e = launch_engine_epilogue(E, R->n_blocks, st)) != cudaSuccess) return lk_cuda_fail(ctx, e, "engine_epilogue");
        R->launch_count += 2;  // modify_reads + epilogue
    }
    return LK_OK;
}

static lk_status run_locked(lk_ctx *ctx, lk_resident *R, float *kernel_ms) {
    if (!R || R->kind != 0) return lk_fail(ctx, LK_ERR_INVALID, "not a PairHMM resident batch");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "cudaSetDevice");
    cudaStream_t st = ctx->stream[0];
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (kernel_ms) {
        cudaEventCreate(&ev0); cudaEventCreate(&ev1);
        cudaEventRecord(ev0, st);
    }
    lk_status ls = enqueue_kernels(ctx, R, st, kernel_ms != nullptr);
    if (kernel_ms) cudaEventRecord(ev1, st);
    e = cudaStreamSynchronize(st);
    if (kernel_ms) {
        if (e == cudaSuccess) cudaEventElapsedTime(kernel_ms, ev0, ev1);
        cudaEventDestroy(ev0); cudaEventDestroy(ev1);
        R->last_fast_ms = 0.f;
        if (R->ev_fast0) {
            if (e == cudaSuccess) cudaEventElapsedTime(&R->last_fast_ms, R->ev_fast0, R->ev_fast1);
            cudaEventDestroy(R->ev_fast0); cudaEventDestroy(R->ev_fast1);
            R->ev_fast0 = R->ev_fast1 = nullptr;
        }
    }
    if (ls != LK_OK) return ls;
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "pairhmm kernels");
    return LK_OK;
}

static lk_status download_locked(lk_ctx *ctx, lk_resident *R, double *out, uint8_t *flags) {
    if (!R || R->kind != 0) return lk_fail(ctx, LK_ERR_INVALID, "not a PairHMM resident batch");
    if (R->n_pairs == 0) return LK_OK;
    if (!out) return lk_fail(ctx, LK_ERR_INVALID, "null out");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "cudaSetDevice");
    cudaStream_t st = ctx->stream[0];
    e = cudaMemcpyAsync(out, R->out.p, (size_t)R->n_pairs * 8, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && flags) e = cudaMemcpyAsync(flags, R->flags.p, (size_t)R->n_pairs, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "download");
    return LK_OK;
}

extern "C" lk_status lk_pairhmm_upload(lk_ctx *ctx, const lk_pairhmm_batch *batch, lk_resident **out_res) {
    LK_ABI_GUARD_BEGIN
    if (!ctx || !out_res) return LK_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    *out_res = nullptr;
    int64_t n_reads = 0, n_haps = 0;
    lk_status vs = validate_batch(ctx, batch, &n_reads, &n_haps);
    if (vs != LK_OK) return vs;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "cudaSetDevice");
    lk_resident *R = new (std::nothrow) lk_resident();
    if (!R) return lk_fail(ctx, LK_ERR_NOMEM, "out of host memory");
    const GcpMode gcp = resolve_gcp(batch, n_reads);
    lk_status st = stage_range(ctx, batch, 0, batch->n_regions, gcp, R, ctx->stream[0]);
    if (st == LK_OK && (e = cudaStreamSynchronize(ctx->stream[0])) != cudaSuccess) st = lk_cuda_fail(ctx, e, "upload sync");
    if (st != LK_OK) { lk_resident_free(ctx, R); return st; }
    *out_res = R;
    return LK_OK;
    LK_ABI_GUARD_END(return lk_fail(ctx, LK_ERR_NOMEM, "out of host memory"), return lk_fail(ctx, LK_ERR_INVALID, "internal error"))
}

extern "C" lk_status lk_pairhmm_run_resident(lk_ctx *ctx, lk_resident *res, float *kernel_ms) {
    if (!ctx) return LK_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    return run_locked(ctx, res, kernel_ms);
}

extern "C" lk_status lk_pairhmm_download(lk_ctx *ctx, lk_resident *res, double *out, uint8_t *flags) {
    if (!ctx) return LK_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    return download_locked(ctx, res, out, flags);
}

static void release_buffers(lk_resident *R) {
    DevBuf *bufs[] = { &R->read_bases, &R->read_qual, &R->read_ins, &R->read_del, &R->read_gcp, &R->read_off,
                       &R->read_region, &R->read_out_base, &R->read_flags, &R->region_hap_start, &R->hap_off,
                       &R->region_read_start, &R->block_region, &R->read_mapq, &R->ref_allele_index, &R->meta, &R->qual_orig, &R->values, &R->keep,
                       &R->hap_bases, &R->hidx, &R->henc_off, &R->hap_flags, &R->tasks, &R->task_haps, &R->groups, &R->out,
                       &R->flags, &R->counters, &R->rescue_pairs, &R->scratch, &R->band_pairs, &R->band_scratch,
                       &R->sw_refs, &R->sw_alts, &R->sw_ref_off, &R->sw_alt_off, &R->sw_bt, &R->sw_bt_off,
                       &R->sw_ops, &R->sw_nops, &R->sw_offsets, &R->sw_order };
    for (DevBuf *b : bufs) b->release();
    R->pinned.release();
    if (R->ev_staged) { cudaEventDestroy(R->ev_staged); R->ev_staged = nullptr; }
}

extern "C" void lk_resident_free(lk_ctx *ctx, lk_resident *R) {
    if (!R) return;
    if (ctx) cudaSetDevice(ctx->device);
    release_buffers(R);
    delete R;
}

// Host-buffer entry point: regions are cut into chunks that flow through two workspaces / two streams, so the
// H2D copy and task building of chunk i+1 and the D2H copy of chunk i-1 overlap the kernels of chunk i.
extern "C" lk_status lk_pairhmm_batch_host(lk_ctx *ctx, const lk_pairhmm_batch *batch, double *out, uint8_t *flags) {
    return lk_pairhmm_pipeline(ctx, batch, out, flags, nullptr);
}

// With a hook the pipeline is the region driver: out / flags are ignored, hook->values / hook->keep are filled.
lk_status lk_pairhmm_pipeline(lk_ctx *ctx, const lk_pairhmm_batch *batch, double *out, uint8_t *flags, const lk_engine_hook *hook) {
    LK_ABI_GUARD_BEGIN
    if (!ctx) return LK_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    timespec ts_entry; clock_gettime(CLOCK_MONOTONIC, &ts_entry);
    const double t_entry = ts_entry.tv_sec * 1e3 + ts_entry.tv_nsec * 1e-6;
    if (hook) { out = hook->values; flags = nullptr; }
    int64_t n_reads = 0, n_haps = 0;
    lk_status vs = validate_batch(ctx, batch, &n_reads, &n_haps);
    if (vs != LK_OK) return vs;
    int64_t n_pairs = 0;
    for (int64_t r = 0; r < batch->n_regions; r++)
        n_pairs += (batch->region_read_start[r + 1] - batch->region_read_start[r]) * (batch->region_hap_start[r + 1] - batch->region_hap_start[r]);
    // reads of a region without haplotypes are all filtered (the maximum over no alleles is -inf): written here, on the host,
    // because a chunk that holds no pair at all launches nothing
    if (hook && hook->keep) {
        cudaPointerAttributes pa;
        const bool keep_on_host = cudaPointerGetAttributes(&pa, hook->keep) != cudaSuccess || pa.type != cudaMemoryTypeDevice;
        (void)cudaGetLastError();
        for (int64_t r = 0; r < batch->n_regions; r++)
            if (batch->region_hap_start[r + 1] == batch->region_hap_start[r] && batch->region_read_start[r + 1] > batch->region_read_start[r]) {
                const size_t nr = (size_t)(batch->region_read_start[r + 1] - batch->region_read_start[r]);
                if (keep_on_host) memset(hook->keep + batch->region_read_start[r], 0, nr);
                else cudaMemset(hook->keep + batch->region_read_start[r], 0, nr);
            }
    }
    if (n_pairs == 0) { ctx->last_launch_count = 0; return LK_OK; }
    if (!out) return lk_fail(ctx, LK_ERR_INVALID, "null out");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "cudaSetDevice");
    const GcpMode gcp = resolve_gcp(batch, n_reads);
    // chunk boundaries by pair count.  Steady state: ~1M pairs per chunk up to 8 chunks, then chunks of <= 4M pairs so the
    // two workspaces stay bounded however large the batch is.  Nothing overlaps the first chunk's task building and H2D,
    // so the batch opens with a ramp (128K pairs, then x3 per chunk): the GPU starts after ~0.4 ms instead of ~3 ms and
    // each chunk's kernels still cover the staging of the next, larger one.
    int64_t n_chunks = std::min<int64_t>(8, std::max<int64_t>(1, n_pairs / (1 << 20)));
    n_chunks = std::max<int64_t>(n_chunks, (n_pairs + (4 << 20) - 1) / (4 << 20));
    const int64_t steady = (n_pairs + n_chunks - 1) / n_chunks;
    std::vector<int64_t> target;  // cumulative pair counts at which a chunk closes
    {
        int64_t acc = 0, size = 128 << 10;
        while (size < steady && acc + size < n_pairs) { acc += size; target.push_back(acc); size *= 3; }
        const int64_t rest = n_pairs - acc, k = std::max<int64_t>(1, (rest + steady - 1) / steady);
        for (int64_t i = 1; i < k; i++) target.push_back(acc + rest * i / k);
    }
    std::vector<int64_t> cut(1, 0), pair_base(1, 0);
    {
        int64_t acc = 0;
        size_t next = 0;
        for (int64_t r = 0; r < batch->n_regions; r++) {
            acc += (batch->region_read_start[r + 1] - batch->region_read_start[r]) * (batch->region_hap_start[r + 1] - batch->region_hap_start[r]);
            if (r + 1 == batch->n_regions || (next < target.size() && acc >= target[next])) {
                cut.push_back(r + 1); pair_base.push_back(acc);
                while (next < target.size() && acc >= target[next]) next++;
            }
        }
    }
    const int C = (int)cut.size() - 1;
    for (int i = 0; i < LK_PIPE_DEPTH; i++)
        if (!ctx->ws[i]) { ctx->ws[i] = new (std::nothrow) lk_resident(); if (!ctx->ws[i]) return lk_fail(ctx, LK_ERR_NOMEM, "out of host memory"); }
    lk_status st = LK_OK;
    int32_t launches = 0;
    auto collect = [&](int i) -> lk_status {  // D2H of chunk i (queued behind its kernels on the same stream)
        lk_resident *R = ctx->ws[i % LK_PIPE_DEPTH];
        cudaStream_t s = ctx->stream[i % LK_PIPE_DEPTH];
        if (R->n_pairs == 0) return LK_OK;
        // cudaMemcpyDefault: the result pointers may be host memory or memory of this device (a caller that hands the
        // results to a collective keeps them in HBM, see include/lorikeet_b200.h)
        cudaError_t ce = cudaMemcpyAsync(out + pair_base[(size_t)i], hook ? R->values.p : R->out.p, (size_t)R->n_pairs * 8, cudaMemcpyDefault, s);
        if (ce == cudaSuccess && hook && hook->keep && R->n_reads > 0)
            ce = cudaMemcpyAsync(hook->keep + batch->region_read_start[cut[(size_t)i]], R->keep.p, (size_t)R->n_reads, cudaMemcpyDefault, s);
        if (ce == cudaSuccess && flags) ce = cudaMemcpyAsync(flags + pair_base[(size_t)i], R->flags.p, (size_t)R->n_pairs, cudaMemcpyDefault, s);
        if (ce != cudaSuccess) return lk_cuda_fail(ctx, ce, "D2H");
        return LK_OK;
    };
    const bool trace = getenv("LK_TRACE") != nullptr;
    auto now_ms = []() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; };
    const double t_begin = now_ms();
    for (int i = 0; i < C && st == LK_OK; i++) {
        lk_resident *R = ctx->ws[i % LK_PIPE_DEPTH];
        cudaStream_t s = ctx->stream[i % LK_PIPE_DEPTH];
        const double t0 = now_ms();
        st = stage_range(ctx, batch, cut[(size_t)i], cut[(size_t)i + 1], gcp, R, s, hook);
        const double t1 = now_ms();
        if (st == LK_OK) st = enqueue_kernels(ctx, R, s, false, hook);
        launches += R->launch_count;
        const double t2 = now_ms();
        if (st == LK_OK && i >= 1) st = collect(i - 1);  // may block on a pageable destination: chunk i is already queued
        if (trace) fprintf(stderr, "[lk] chunk %d/%d: stage %.2f ms, enqueue %.2f ms, collect %.2f ms (t=%.2f)\n", i, C, t1 - t0, t2 - t1, now_ms() - t2, now_ms() - t_begin);
    }
    if (st == LK_OK) st = collect(C - 1);
    if (trace) fprintf(stderr, "[lk] all %d chunks queued (t=%.2f)\n", C, now_ms() - t_begin);
    for (int i = 0; i < LK_PIPE_DEPTH; i++) {
        e = cudaStreamSynchronize(ctx->stream[i]);
        if (e != cudaSuccess && st == LK_OK) st = lk_cuda_fail(ctx, e, "pairhmm pipeline");
    }
    if (trace) fprintf(stderr, "[lk] drained (t=%.2f; entry -> first chunk %.2f ms)\n", now_ms() - t_begin, t_begin - t_entry);
    ctx->last_launch_count = launches;
    return st;
    LK_ABI_GUARD_END(return lk_fail(ctx, LK_ERR_NOMEM, "out of host memory"), return lk_fail(ctx, LK_ERR_INVALID, "internal error"))
}

extern "C" lk_status lk_pairhmm_forward(lk_ctx *ctx, const uint8_t *hap, int64_t hap_len, const uint8_t *read,
                                        const uint8_t *qual, const uint8_t *ins, const uint8_t *del, const uint8_t *gcp,
                                        int64_t read_len, double *out_log10) {
    if (!ctx) return LK_ERR_INVALID;
    if (!hap || !read || !qual || !ins || !del || !gcp || !out_log10) return lk_fail(ctx, LK_ERR_INVALID, "null argument");
    if (hap_len <= 0 || read_len <= 0) return lk_fail(ctx, LK_ERR_INVALID, "empty read or haplotype");
    const int64_t rs[2] = { 0, 1 }, ro[2] = { 0, read_len }, ho[2] = { 0, hap_len };
    lk_pairhmm_batch b;
    memset(&b, 0, sizeof b);
    b.n_regions = 1; b.region_read_start = rs; b.region_hap_start = rs; b.read_off = ro;
    b.read_bases = read; b.read_qual = qual; b.read_ins = ins; b.read_del = del; b.read_gcp = gcp;
    b.gcp_constant = gcp[0]; b.hap_off = ho; b.hap_bases = hap;
    return lk_pairhmm_batch_host(ctx, &b, out_log10, nullptr);
}

extern "C" int32_t lk_debug_last_launch_count(const lk_ctx *ctx) { return ctx ? ctx->last_launch_count : -1; }
extern "C" int32_t lk_resident_launch_count(const lk_resident *r) { return r ? r->launch_count : -1; }
extern "C" int64_t lk_resident_cells(const lk_resident *r) { return r ? r->n_cells : -1; }
extern "C" int64_t lk_resident_pairs(const lk_resident *r) { return r ? r->n_pairs : -1; }

// device pointer of the resident result block (f64, n_pairs) so a caller can hand it to a collective
extern "C" void *lk_resident_out_device_ptr(const lk_resident *r) { return r ? r->out.p : nullptr; }
// duration of the fast-path (dominant) kernel launches inside the last timed lk_pairhmm_run_resident
extern "C" float lk_resident_last_fast_ms(const lk_resident *r) { return r ? r->last_fast_ms : -1.f; }

// modify_read_qualities on the device, stand-alone: the kernel the region driver runs on its staged copies, applied in
// place to the caller's arrays (upload, one launch, download).  Same contract as the host utility
// lk_engine_modify_read_qualities; exists so the device pre-ops can be compared byte for byte with it.
extern "C" lk_status lk_engine_modify_read_qualities_device(lk_ctx *ctx, const lk_engine_args *args, int64_t n_reads,
                                                            const int64_t *read_off, const uint8_t *read_bases, uint8_t *read_qual,
                                                            uint8_t *read_ins, uint8_t *read_del, const uint8_t *read_mapq) {
    if (!ctx) return LK_ERR_INVALID;
    if (!args || n_reads < 0) return lk_fail(ctx, LK_ERR_INVALID, "bad arguments");
    if (n_reads == 0) return LK_OK;
    if (!read_off || !read_bases || !read_qual || !read_ins || !read_del) return lk_fail(ctx, LK_ERR_INVALID, "null read arrays");
    if (!args->disable_cap_read_qualities_to_mapq && !read_mapq) return lk_fail(ctx, LK_ERR_INVALID, "read_mapq required");
    if (args->pcr_error_model < 0 || args->pcr_error_model > 3) return lk_fail(ctx, LK_ERR_INVALID, "bad PCR error model");
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "cudaSetDevice");
    const int64_t b0 = read_off[0], nb = read_off[n_reads] - b0;
    DevBuf d_off, d_b, d_q, d_i, d_d, d_m;
    cudaStream_t st = ctx->stream[0];
    auto fail = [&](const char *what) { d_off.release(); d_b.release(); d_q.release(); d_i.release(); d_d.release(); d_m.release(); return lk_cuda_fail(ctx, e, what); };
    if ((e = d_off.reserve((size_t)(n_reads + 1) * 8)) != cudaSuccess || (e = d_b.reserve((size_t)std::max<int64_t>(nb, 16))) != cudaSuccess ||
        (e = d_q.reserve((size_t)std::max<int64_t>(nb, 16))) != cudaSuccess || (e = d_i.reserve((size_t)std::max<int64_t>(nb, 16))) != cudaSuccess ||
        (e = d_d.reserve((size_t)std::max<int64_t>(nb, 16))) != cudaSuccess || (e = d_m.reserve((size_t)n_reads)) != cudaSuccess)
        return fail("alloc");
    if ((e = cudaMemcpyAsync(d_off.p, read_off, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_b.p, read_bases + b0, (size_t)nb, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_q.p, read_qual + b0, (size_t)nb, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_i.p, read_ins + b0, (size_t)nb, cudaMemcpyHostToDevice, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(d_d.p, read_del + b0, (size_t)nb, cudaMemcpyHostToDevice, st)) != cudaSuccess)
        return fail("H2D");
    if (read_mapq && (e = cudaMemcpyAsync(d_m.p, read_mapq, (size_t)n_reads, cudaMemcpyHostToDevice, st)) != cudaSuccess) return fail("H2D mapq");
    EngineDevArgs E;
    memset(&E, 0, sizeof E);
    E.read_bases = (const uint8_t *)d_b.p - b0;
    E.read_qual = (uint8_t *)d_q.p - b0; E.read_ins = (uint8_t *)d_i.p - b0; E.read_del = (uint8_t *)d_d.p - b0;
    E.read_off = (const int64_t *)d_off.p;
    E.read_mapq = read_mapq ? (const uint8_t *)d_m.p : nullptr;
    E.use_pcr = args->pcr_error_model != LK_PCR_NONE;
    E.cap_to_mapq = !args->disable_cap_read_qualities_to_mapq;
    E.bq_threshold = args->base_quality_score_threshold;
    lk_engine_fill_pcr_cache(args, E.pcr_cache);
    if ((e = launch_engine_modify(E, n_reads, st)) != cudaSuccess) return fail("engine_modify_reads");
    if ((e = cudaMemcpyAsync(read_qual + b0, d_q.p, (size_t)nb, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(read_ins + b0, d_i.p, (size_t)nb, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaMemcpyAsync(read_del + b0, d_d.p, (size_t)nb, cudaMemcpyDeviceToHost, st)) != cudaSuccess ||
        (e = cudaStreamSynchronize(st)) != cudaSuccess)
        return fail("D2H");
    d_off.release(); d_b.release(); d_q.release(); d_i.release(); d_d.release(); d_m.release();
    return LK_OK;
}

// normalize_likelihoods + filter_poorly_modeled_evidence on the device, stand-alone: the epilogue kernel the region driver runs
// after the PairHMM kernels, applied to likelihoods the CALLER supplies (layout of lk_pairhmm_batch_host: per region [r*A + h]).
// Exists so that the post-ops can be checked on their own against the reference's AlleleLikelihoods tests.
extern "C" lk_status lk_engine_normalize_and_filter_device(lk_ctx *ctx, const lk_engine_args *args, const lk_pairhmm_batch *batch,
                                                           const int32_t *ref_allele_index, const double *likelihoods,
                                                           double *values, uint8_t *keep) {
    LK_ABI_GUARD_BEGIN
    if (!ctx) return LK_ERR_INVALID;
    if (!args || !batch) return lk_fail(ctx, LK_ERR_INVALID, "null arguments");
    int64_t n_reads = 0, n_haps = 0;
    {
        std::lock_guard<std::mutex> g(ctx->mu);
        const lk_status vs = validate_batch(ctx, batch, &n_reads, &n_haps);
        if (vs != LK_OK) return vs;
    }
    const int64_t G = batch->n_regions;
    std::vector<int64_t> out_base((size_t)std::max<int64_t>(n_reads, 1)), brs;
    std::vector<int32_t> breg;
    int64_t n_pairs = 0;
    const bool sampled = batch->region_sample_start && batch->sample_read_start;
    for (int64_t g = 0; g < G; g++) {
        const int64_t r0 = batch->region_read_start[g], r1 = batch->region_read_start[g + 1];
        const int64_t A = batch->region_hap_start[g + 1] - batch->region_hap_start[g];
        for (int64_t r = r0; r < r1; r++) out_base[(size_t)r] = n_pairs + (r - r0) * A;
        n_pairs += (r1 - r0) * A;
        if (sampled) {
            for (int64_t sm = batch->region_sample_start[g]; sm < batch->region_sample_start[g + 1]; sm++) {
                brs.push_back(batch->sample_read_start[sm]);
                breg.push_back((int32_t)g);
            }
        } else {
            brs.push_back(r0);
        }
    }
    brs.push_back(n_reads);
    if (n_reads == 0) return LK_OK;
    if (n_pairs > 0 && (!likelihoods || !values)) return lk_fail(ctx, LK_ERR_INVALID, "null likelihoods / values");
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "cudaSetDevice");
    cudaStream_t st = ctx->stream[0];
    const int64_t q0 = batch->read_off[0], qn = batch->read_off[n_reads] - q0;
    DevBuf d_off, d_qual, d_base, d_brs, d_breg, d_rhs, d_ref, d_lk, d_val, d_keep;
    DevBuf *all[] = { &d_off, &d_qual, &d_base, &d_brs, &d_breg, &d_rhs, &d_ref, &d_lk, &d_val, &d_keep };
    auto done = [&](lk_status r) { for (DevBuf *b : all) b->release(); return r; };
    if ((e = upload(d_off, batch->read_off, (size_t)n_reads + 1, st)) != cudaSuccess ||
        (e = upload(d_qual, batch->read_qual + q0, (size_t)qn, st)) != cudaSuccess ||
        (e = upload(d_base, out_base.data(), (size_t)n_reads, st)) != cudaSuccess ||
        (e = upload(d_brs, brs.data(), brs.size(), st)) != cudaSuccess ||
        (e = upload(d_breg, breg.data(), breg.size(), st)) != cudaSuccess ||
        (e = upload(d_rhs, batch->region_hap_start, (size_t)G + 1, st)) != cudaSuccess ||
        (e = upload(d_ref, ref_allele_index, ref_allele_index ? (size_t)G : 0, st)) != cudaSuccess ||
        (e = upload(d_lk, likelihoods, (size_t)n_pairs, st)) != cudaSuccess ||
        (e = d_val.reserve((size_t)std::max<int64_t>(n_pairs, 1) * 8)) != cudaSuccess || (e = d_keep.reserve((size_t)n_reads)) != cudaSuccess)
        return done(lk_cuda_fail(ctx, e, "stage post-ops inputs"));
    EngineDevArgs E;
    memset(&E, 0, sizeof E);
    E.qual_orig = (const uint8_t *)d_qual.p - q0;
    E.read_off = (const int64_t *)d_off.p;
    E.read_out_base = (const int64_t *)d_base.p;
    E.block_read_start = (const int64_t *)d_brs.p;
    E.block_region = sampled ? (const int32_t *)d_breg.p : nullptr;
    E.region_hap_start = (const int64_t *)d_rhs.p;
    E.ref_allele_index = ref_allele_index ? (const int32_t *)d_ref.p : nullptr;
    E.lk = (const double *)d_lk.p;
    E.values = (double *)d_val.p;
    E.keep = (uint8_t *)d_keep.p;
    E.cap_diff = args->log10_global_read_mismapping_rate;
    E.expected_error_rate = args->expected_error_rate_per_base;
    E.disq_scale = args->read_disqualification_scale;
    E.symmetric = args->symmetrically_normalize_alleles_to_reference != 0;
    E.dynamic_disq = args->dynamic_read_disqualification != 0;
    if ((e = cudaMemsetAsync(d_keep.p, 0, (size_t)n_reads, st)) != cudaSuccess ||
        (e = launch_engine_epilogue(E, (int64_t)brs.size() - 1, st)) != cudaSuccess)
        return done(lk_cuda_fail(ctx, e, "engine_epilogue"));
    if (n_pairs > 0 && (e = cudaMemcpyAsync(values, d_val.p, (size_t)n_pairs * 8, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
        return done(lk_cuda_fail(ctx, e, "D2H values"));
    if (keep && (e = cudaMemcpyAsync(keep, d_keep.p, (size_t)n_reads, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
        return done(lk_cuda_fail(ctx, e, "D2H keep"));
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return done(lk_cuda_fail(ctx, e, "post-ops"));
    return done(LK_OK);
    LK_ABI_GUARD_END(return lk_fail(ctx, LK_ERR_NOMEM, "out of host memory"), return lk_fail(ctx, LK_ERR_INVALID, "internal error"))
}

// Page-lock memory the CALLER owns (its read / haplotype arenas, its result arrays) so that the copies of the host-buffer entry
// points are true asynchronous DMA instead of being staged through the driver's bounce buffer; portable across the devices of
// a pool.  Optional: every entry point works with pageable memory too.
extern "C" lk_status lk_host_register(void *ptr, size_t bytes) {
    if (!ptr || bytes == 0) return LK_ERR_INVALID;
    const cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return e == cudaErrorMemoryAllocation ? LK_ERR_NOMEM : LK_ERR_CUDA; }
    return LK_OK;
}
extern "C" lk_status lk_host_unregister(void *ptr) {
    if (!ptr) return LK_ERR_INVALID;
    const cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { (void)cudaGetLastError(); return LK_ERR_CUDA; }
    return LK_OK;
}
