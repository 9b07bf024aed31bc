// host_batcher.cpp — cross-thread batching front-end (SURVEY.md 8(b) "Threading", 8(f) N4).
//
// Lorikeet calls the likelihood engine from Rayon workers, ONE region per task
// (src/assembly/assembly_region_walker.rs:140-277: `par_iter` over assembly regions, an engine clone per region), and a
// single region of 64 x 16 pairs keeps well under 1 % of a B200 busy.  The batcher lets those call sites stay as they
// are: every worker thread submits its region through a blocking call; a dispatcher thread concatenates whatever is
// pending (up to max_regions, waiting at most max_wait_us for stragglers) into one batch, runs the region driver
// (lk_engine_compute_read_likelihoods, or the bare lk_pairhmm_batch_host when no engine arguments were given) once, and
// hands every caller its own slice of the result.
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <deque>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/lorikeet_b200.h"
#include "lk_api_internal.h"

namespace {

struct Request {
    const lk_pairhmm_batch *region;  // n_regions == 1, the caller's arrays (valid until the call returns)
    const uint8_t *read_mapq;
    int32_t ref_allele_index;
    double *values;
    uint8_t *keep_or_flags;
    lk_status status = LK_OK;
    bool done = false;
    // every caller sleeps on ITS OWN condition variable: with one shared one, each finished launch woke every blocked worker
    // of every device (hundreds of threads, thousands of launches per second) just to find that it was not their turn
    std::condition_variable cv;
};

}  // namespace

struct lk_batcher {
    std::vector<lk_ctx *> ctxs;  // one per device (lk_batcher_create: one; lk_pool_batcher_create: the pool's)
    bool engine = false;
    lk_engine_args args;
    int32_t max_regions = 256;
    int64_t max_wait_us = 200;
    std::mutex mu;
    std::condition_variable cv_work;
    std::deque<Request *> queue;
    bool stop = false;
    // statistics
    int64_t n_batches = 0, n_regions = 0;
    int idle_primaries = 0;  // primary dispatchers (one per device) waiting for work: a waking one leaves them their share
    // gather buffers, one set per dispatcher thread, kept across batches (fresh vectors cost more in page faults than the
    // copies)
    // The byte arenas a launch copies to the device, and the result buffers it copies back into, live in PINNED host memory
    // (grow-only, portable across the pool's devices): a cudaMemcpyAsync from / to pageable memory is staged through the
    // driver's own bounce buffer and blocks the dispatcher for the duration of the copy.
    template <class T> struct PinVec {
        T *p = nullptr;
        size_t n = 0, cap = 0;
        PinVec() = default;
        PinVec(const PinVec &) = delete;
        PinVec &operator=(const PinVec &) = delete;
        ~PinVec() { if (p) cudaFreeHost(p); }
        void reserve(size_t want) {
            if (want <= cap) return;
            const size_t c = want + want / 2 + 4096;
            T *q = nullptr;
            if (cudaHostAlloc((void **)&q, c * sizeof(T), cudaHostAllocPortable) != cudaSuccess) { (void)cudaGetLastError(); throw std::bad_alloc(); }
            if (n) memcpy(q, p, n * sizeof(T));
            if (p) cudaFreeHost(p);
            p = q; cap = c;
        }
        void clear() { n = 0; }
        void resize(size_t m) { reserve(m); n = m; }
        void append(const T *src, size_t m) { reserve(n + m); memcpy(p + n, src, m * sizeof(T)); n += m; }
        void append_fill(size_t m, T v) { reserve(n + m); for (size_t i = 0; i < m; i++) p[n + i] = v; n += m; }
        T *data() { return p; }
        size_t size() const { return n; }
    };
    struct Scratch {
        std::vector<int64_t> rrs, rhs, read_off, hap_off, rss, srs;
        PinVec<uint8_t> bases, qual, ins, del, gcp, haps, mapq, aux;
        std::vector<int32_t> ref_idx;
        PinVec<double> out;
    };
    // two dispatcher threads PER DEVICE: the device calls serialise on the context, but one thread gathers / scatters its
    // batch while the other's is on the GPU.  All dispatchers of all devices pull from the one queue, so the deal over
    // devices is by time: whichever device is free takes the next share.
    static constexpr int kWorkers = 2;
    static constexpr int kBacklogForSecond = 48;
    struct Worker {
        Scratch scratch;
        std::thread thread;
        size_t last_batch = 1;  // callers released by this dispatcher's previous batch: once that many are back there is nobody left to wait for
    };
    std::vector<Worker> workers;  // [device * kWorkers + w]

    void run(int dev, int w);
    void process(lk_ctx *ctx, std::vector<Request *> &reqs, Scratch &S);
};

static double bt_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

void lk_batcher::process(lk_ctx *ctx, std::vector<Request *> &reqs, Scratch &S) {
    auto &rrs = S.rrs; auto &rhs = S.rhs; auto &read_off = S.read_off; auto &hap_off = S.hap_off;
    auto &bases = S.bases; auto &qual = S.qual; auto &ins = S.ins; auto &del = S.del; auto &gcp = S.gcp; auto &haps = S.haps;
    auto &mapq = S.mapq; auto &aux = S.aux; auto &ref_idx = S.ref_idx; auto &out = S.out;
    static const bool trace = getenv("LK_TRACE") != nullptr;
    const double t0 = bt_now_ms();
    // concatenate the regions into one batch (host copies of a few KB per region)
    const size_t G = reqs.size();
    rrs.assign(G + 1, 0); rhs.assign(G + 1, 0); read_off.assign(1, 0); hap_off.assign(1, 0);
    bases.clear(); qual.clear(); ins.clear(); del.clear(); gcp.clear(); haps.clear(); mapq.clear();
    ref_idx.assign(G, 0);
    bool any_gcp = false, need_mapq = engine && !args.disable_cap_read_qualities_to_mapq;
    // callers are independent: a region with a per-base gcp array, or two regions with different constants, turn the merged
    // batch into the per-base form (each region keeps exactly the penalties its caller gave)
    for (Request *q : reqs) any_gcp |= (q->region->read_gcp != nullptr) || (q->region->gcp_constant != reqs[0]->region->gcp_constant);
    lk_status bad = LK_OK;
    // several samples per region (engine mode): as soon as one caller's region has them, the merged batch carries sample
    // arrays for all (a region without them is one sample)
    bool any_samples = false;
    for (Request *q : reqs) any_samples |= engine && q->region->region_sample_start && q->region->sample_read_start;
    auto &rss = S.rss; auto &srs = S.srs;
    rss.assign(1, 0); srs.assign(1, 0);
    for (size_t g = 0; g < G; g++) {
        const lk_pairhmm_batch *b = reqs[g]->region;
        const int64_t nr = b->region_read_start[1], nh = b->region_hap_start[1];
        if (any_samples) {
            if (b->region_sample_start && b->sample_read_start) {
                const int64_t ns = b->region_sample_start[1];
                if (ns < 0 || b->region_sample_start[0] != 0 || (ns == 0 && nr > 0) || (ns > 0 && (b->sample_read_start[0] != 0 || b->sample_read_start[ns] != nr))) { bad = LK_ERR_INVALID; break; }
                for (int64_t sm = 1; sm <= ns; sm++) srs.push_back(rrs[g] + b->sample_read_start[sm]);
                rss.push_back(rss.back() + ns);
            } else {
                srs.push_back(rrs[g] + nr);
                rss.push_back(rss.back() + 1);
            }
        }
        rrs[g + 1] = rrs[g] + nr;
        rhs[g + 1] = rhs[g] + nh;
        const int64_t rb0 = nr ? b->read_off[0] : 0, rb1 = nr ? b->read_off[nr] : 0;
        const int64_t hb0 = nh ? b->hap_off[0] : 0, hb1 = nh ? b->hap_off[nh] : 0;
        for (int64_t r = 0; r < nr; r++) read_off.push_back(read_off.back() + (b->read_off[r + 1] - b->read_off[r]));
        for (int64_t h = 0; h < nh; h++) hap_off.push_back(hap_off.back() + (b->hap_off[h + 1] - b->hap_off[h]));
        bases.append(b->read_bases + rb0, (size_t)(rb1 - rb0));
        qual.append(b->read_qual + rb0, (size_t)(rb1 - rb0));
        ins.append(b->read_ins + rb0, (size_t)(rb1 - rb0));
        del.append(b->read_del + rb0, (size_t)(rb1 - rb0));
        if (any_gcp) {
            if (b->read_gcp) gcp.append(b->read_gcp + rb0, (size_t)(rb1 - rb0));
            else gcp.append_fill((size_t)(rb1 - rb0), b->gcp_constant);
        }
        haps.append(b->hap_bases + hb0, (size_t)(hb1 - hb0));
        if (need_mapq && nr > 0) {
            if (!reqs[g]->read_mapq) { bad = LK_ERR_INVALID; break; }  // (rejected at submit; kept as a guard)
            mapq.append(reqs[g]->read_mapq, (size_t)nr);
        }
        ref_idx[g] = reqs[g]->ref_allele_index;
    }
    lk_pairhmm_batch big;
    memset(&big, 0, sizeof big);
    big.n_regions = (int64_t)G;
    big.region_read_start = rrs.data(); big.region_hap_start = rhs.data();
    big.read_off = read_off.data();
    big.read_bases = bases.data(); big.read_qual = qual.data(); big.read_ins = ins.data(); big.read_del = del.data();
    big.read_gcp = any_gcp ? gcp.data() : nullptr;
    big.gcp_constant = reqs[0]->region->gcp_constant;
    big.hap_off = hap_off.data(); big.hap_bases = haps.data();
    if (any_samples && bad == LK_OK) { big.region_sample_start = rss.data(); big.sample_read_start = srs.data(); }
    int64_t n_pairs = 0;
    for (size_t g = 0; g < G; g++) n_pairs += (rrs[g + 1] - rrs[g]) * (rhs[g + 1] - rhs[g]);
    out.resize((size_t)n_pairs);
    aux.resize((size_t)(engine ? rrs[G] : n_pairs));
    lk_status st = bad;
    const double t1 = bt_now_ms();
    if (st == LK_OK) {
        if (engine) st = lk_engine_compute_read_likelihoods(ctx, &args, &big, need_mapq ? mapq.data() : nullptr, ref_idx.data(), out.data(), aux.data());
        else st = lk_pairhmm_batch_host(ctx, &big, out.data(), aux.data());
    }
    const double t2 = bt_now_ms();
    int64_t pbase = 0;
    for (size_t g = 0; g < G; g++) {
        Request *q = reqs[g];
        const int64_t nr = rrs[g + 1] - rrs[g], np = nr * (rhs[g + 1] - rhs[g]);
        if (st == LK_OK) {
            if (np && q->values) memcpy(q->values, out.data() + pbase, (size_t)np * 8);
            if (q->keep_or_flags) {
                if (engine) { if (nr) memcpy(q->keep_or_flags, aux.data() + rrs[g], (size_t)nr); }
                else if (np) memcpy(q->keep_or_flags, aux.data() + pbase, (size_t)np);
            }
        }
        q->status = st;
        pbase += np;
    }
    if (trace) fprintf(stderr, "[lk] batcher: %zu regions, gather %.2f ms, device call %.2f ms, scatter %.2f ms\n", G, t1 - t0, t2 - t1, bt_now_ms() - t2);
}

void lk_batcher::run(int dev, int w) {
    lk_ctx *ctx = ctxs[(size_t)dev];
    Worker &me = workers[(size_t)(dev * kWorkers + w)];
    const size_t n_dev = ctxs.size();
    std::unique_lock<std::mutex> lock(mu);
    for (;;) {
        // The second dispatcher of a device only steps in for a backlog (callers that keep many regions in flight): with
        // blocked callers it would split them into two half-sized groups that alternate, each launch paying the fixed
        // latency for half the regions (16 threads: 925 instead of 1 170 GCUPS).
        const size_t need = (w == 0) ? 1 : (size_t)kBacklogForSecond * n_dev;
        if (w == 0) idle_primaries++;
        cv_work.wait(lock, [&] { return stop || queue.size() >= need; });
        const int idle_now = idle_primaries;  // primaries (me included) that could share what is queued
        if (w == 0) idle_primaries--;
        if (stop && queue.empty()) return;
        if (queue.size() < need) { cv_work.wait_for(lock, std::chrono::microseconds(50)); continue; }  // stopping: leave the rest to a primary
        // give stragglers a moment: a fuller batch uses the GPU better than a prompt one
        if ((int32_t)queue.size() < max_regions && max_wait_us > 0 && !stop)
            cv_work.wait_for(lock, std::chrono::microseconds(max_wait_us),
                             [&] { return stop || (int32_t)queue.size() >= max_regions || queue.size() >= me.last_batch * (size_t)(n_dev > 1 ? idle_now : 1); });
        // several devices: leave the other idle devices their share of what is queued (a share below ~8 regions is not
        // worth a launch of its own).  Measured alternative (round 2, 4 GPUs): making every dispatcher of a pool wait for and
        // take at least 32 regions -- fewer, fuller launches -- left 256 blocked threads and the asynchronous callers where
        // they were (5.6 / 6.1 / 9.4 TCUPS) and cost 64 blocked threads a third (4.0 -> 2.7 TCUPS): dropped.
        size_t take = (size_t)max_regions;
        if (n_dev > 1 && w == 0 && idle_now > 1) {
            const size_t share = (queue.size() + (size_t)idle_now - 1) / (size_t)idle_now;
            take = std::min(take, std::max<size_t>(share, 8));
        }
        std::vector<Request *> reqs;
        while (!queue.empty() && reqs.size() < take) { reqs.push_back(queue.front()); queue.pop_front(); }
        if (reqs.empty()) continue;  // another dispatcher took everything while this one waited for stragglers
        const bool more = !queue.empty();
        lock.unlock();
        if (more) cv_work.notify_all();
        try {
            process(ctx, reqs, me.scratch);
        } catch (...) {  // gather buffers could not grow: fail this batch's callers, keep the dispatcher alive
            for (Request *q : reqs) q->status = LK_ERR_NOMEM;
        }
        lock.lock();
        n_batches++;
        n_regions += (int64_t)reqs.size();
        me.last_batch = reqs.size();
        for (Request *q : reqs) { q->done = true; q->cv.notify_one(); }
    }
}

static lk_status batcher_create(lk_ctx *const *ctxs, int n_ctx, const lk_engine_args *args, int32_t max_regions, int32_t max_wait_us,
                                lk_batcher **out) {
    if (!ctxs || n_ctx < 1 || !out) return LK_ERR_INVALID;
    *out = nullptr;
    for (int i = 0; i < n_ctx; i++)
        if (!ctxs[i]) return LK_ERR_INVALID;
    if (max_regions < 1 || max_wait_us < 0) return lk_fail(ctxs[0], LK_ERR_INVALID, "bad batcher arguments");
    if (args && (args->pcr_error_model < 0 || args->pcr_error_model > 3)) return lk_fail(ctxs[0], LK_ERR_INVALID, "bad PCR error model");
    try {
        lk_batcher *b = new lk_batcher();
        b->ctxs.assign(ctxs, ctxs + n_ctx);
        b->engine = args != nullptr;
        if (args) b->args = *args;
        b->max_regions = max_regions;
        b->max_wait_us = max_wait_us;
        b->workers = std::vector<lk_batcher::Worker>((size_t)(n_ctx * lk_batcher::kWorkers));
        for (int d = 0; d < n_ctx; d++)
            for (int w = 0; w < lk_batcher::kWorkers; w++)
                b->workers[(size_t)(d * lk_batcher::kWorkers + w)].thread = std::thread([b, d, w] {
                    try { b->run(d, w); } catch (...) { /* a dispatcher must never unwind into std::terminate */ }
                });
        *out = b;
        return LK_OK;
    } catch (...) {
        return lk_fail(ctxs[0], LK_ERR_NOMEM, "out of host memory (batcher)");
    }
}

extern "C" lk_status lk_batcher_create(lk_ctx *ctx, const lk_engine_args *args, int32_t max_regions, int32_t max_wait_us,
                                       lk_batcher **out) {
    return batcher_create(&ctx, 1, args, max_regions, max_wait_us, out);
}

lk_status lk_batcher_create_multi(lk_ctx *const *ctxs, int n_ctx, const lk_engine_args *args, int32_t max_regions,
                                  int32_t max_wait_us, lk_batcher **out) {
    return batcher_create(ctxs, n_ctx, args, max_regions, max_wait_us, out);
}

extern "C" void lk_batcher_destroy(lk_batcher *b) {
    if (!b) return;
    {
        std::lock_guard<std::mutex> g(b->mu);
        b->stop = true;
    }
    b->cv_work.notify_all();
    for (auto &w : b->workers)
        if (w.thread.joinable()) w.thread.join();
    delete b;
}

struct lk_ticket { Request q; };

extern "C" lk_status lk_batcher_submit_async(lk_batcher *b, const lk_pairhmm_batch *region, const uint8_t *read_mapq,
                                             int32_t ref_allele_index, double *values, uint8_t *keep_or_flags,
                                             lk_ticket **out_ticket) {
    if (!b || !region || !out_ticket) return LK_ERR_INVALID;
    *out_ticket = nullptr;
    if (region->n_regions != 1 || !region->region_read_start || !region->region_hap_start || region->region_read_start[0] != 0 ||
        region->region_hap_start[0] != 0)
        return LK_ERR_INVALID;
    int64_t np = 0;
    if (lk_pairhmm_batch_dims(region, &np, nullptr) != LK_OK) return LK_ERR_INVALID;
    if (np > 0 && !values) return LK_ERR_INVALID;
    // a request that could only fail inside a merged batch is rejected here, so it cannot take co-batched callers down
    if (b->engine && !b->args.disable_cap_read_qualities_to_mapq && !read_mapq && region->region_read_start[1] > 0) return LK_ERR_INVALID;
    lk_ticket *t = new (std::nothrow) lk_ticket();
    if (!t) return LK_ERR_NOMEM;
    Request &q = t->q;
    q.region = region; q.read_mapq = read_mapq; q.ref_allele_index = ref_allele_index; q.values = values; q.keep_or_flags = keep_or_flags;
    {
        std::unique_lock<std::mutex> lock(b->mu);
        if (b->stop) { delete t; return LK_ERR_INVALID; }
        b->queue.push_back(&q);
    }
    b->cv_work.notify_all();
    *out_ticket = t;
    return LK_OK;
}

extern "C" lk_status lk_batcher_wait(lk_batcher *b, lk_ticket *ticket) {
    if (!b || !ticket) return LK_ERR_INVALID;
    lk_status st;
    {
        std::unique_lock<std::mutex> lock(b->mu);
        ticket->q.cv.wait(lock, [&] { return ticket->q.done; });
        st = ticket->q.status;
    }
    delete ticket;
    return st;
}

extern "C" lk_status lk_batcher_submit(lk_batcher *b, const lk_pairhmm_batch *region, const uint8_t *read_mapq,
                                       int32_t ref_allele_index, double *values, uint8_t *keep_or_flags) {
    lk_ticket *t = nullptr;
    const lk_status st = lk_batcher_submit_async(b, region, read_mapq, ref_allele_index, values, keep_or_flags, &t);
    if (st != LK_OK) return st;
    return lk_batcher_wait(b, t);
}

extern "C" void lk_batcher_stats(lk_batcher *b, int64_t *n_batches, int64_t *n_regions) {
    if (!b) return;
    std::lock_guard<std::mutex> g(b->mu);
    if (n_batches) *n_batches = b->n_batches;
    if (n_regions) *n_regions = b->n_regions;
}
