// host_pool.cpp — several devices behind ONE handle, in ONE process, no torch ("one host thread + lk_ctx
// + streams per device, dispatcher above").
//
// Lorikeet is a single process whose Rayon workers call the likelihood engine
// (src/assembly/assembly_region_walker.rs:140-277); a build that links this library reaches GPUs 1..N-1 through a pool:
//   * lk_pool_compute_read_likelihoods / lk_pool_pairhmm_batch_host: one batch of regions cut into contiguous chunks of
//     about equal CELL counts (cost is proportional to sum R x sum H per region, which varies), dealt dynamically: one host
//     thread per device pulls the next chunk when its device is free and runs the ordinary single-device call (three
//     streams, chunk pipeline) on a view of the caller's arrays -- regions are independent, so there is no data-path
//     exchange and no collective; every device writes its slice of the caller's result arrays;
//   * lk_pool_batcher_create: the cross-thread batcher (host_batcher.cpp) with the dispatcher threads of ALL devices
//     pulling from one queue.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/lorikeet_b200.h"
#include "lk_api_internal.h"

lk_status lk_batcher_create_multi(lk_ctx *const *ctxs, int n_ctx, const lk_engine_args *args, int32_t max_regions,
                                  int32_t max_wait_us, lk_batcher **out);

struct lk_pool {
    std::vector<lk_ctx *> ctxs;
    std::mutex mu;  // guards err
    std::string err;
    int32_t chunks_per_device = 4;
};

extern "C" lk_status lk_pool_create(const int *devices, int32_t n_devices, lk_pool **out) {
    if (!out) return LK_ERR_INVALID;
    *out = nullptr;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0) {
        (void)cudaGetLastError();
        return LK_ERR_CUDA;  // no CPU fallback by design
    }
    std::vector<int> devs;
    try {
        if (!devices || n_devices <= 0) {
            for (int d = 0; d < visible; d++) devs.push_back(d);
        } else {
            for (int i = 0; i < n_devices; i++) {
                if (devices[i] < 0 || devices[i] >= visible) return LK_ERR_INVALID;
                if (std::find(devs.begin(), devs.end(), devices[i]) != devs.end()) return LK_ERR_INVALID;  // a device once
                devs.push_back(devices[i]);
            }
        }
        lk_pool *p = new lk_pool();
        for (int d : devs) {
            lk_ctx *c = nullptr;
            const lk_status st = lk_create(d, &c);
            if (st != LK_OK) {
                for (lk_ctx *x : p->ctxs) lk_destroy(x);
                delete p;
                return st;
            }
            p->ctxs.push_back(c);
        }
        *out = p;
        return LK_OK;
    } catch (...) {
        return LK_ERR_NOMEM;
    }
}

extern "C" void lk_pool_destroy(lk_pool *pool) {
    if (!pool) return;
    for (lk_ctx *c : pool->ctxs) lk_destroy(c);
    delete pool;
}

extern "C" int32_t lk_pool_device_count(const lk_pool *pool) { return pool ? (int32_t)pool->ctxs.size() : 0; }

extern "C" lk_ctx *lk_pool_ctx(lk_pool *pool, int32_t i) {
    if (!pool || i < 0 || i >= (int32_t)pool->ctxs.size()) return nullptr;
    return pool->ctxs[(size_t)i];
}

extern "C" const char *lk_pool_last_error(const lk_pool *pool) { return pool ? pool->err.c_str() : "null pool"; }

extern "C" lk_status lk_pool_set_tristate_correction(lk_pool *pool, int use_tristate) {
    if (!pool) return LK_ERR_INVALID;
    for (lk_ctx *c : pool->ctxs) {
        const lk_status st = lk_set_tristate_correction(c, use_tristate);
        if (st != LK_OK) return st;
    }
    return LK_OK;
}

namespace {

struct Chunk { int64_t g0, g1; };

// Contiguous region ranges of about equal cell counts, chunks_per_device per device (dynamic deal evens out what is left).
lk_status cut_chunks(const lk_pairhmm_batch *b, int n_dev, int chunks_per_device, std::vector<Chunk> &chunks,
                     std::vector<int64_t> &pair_base) {
    int64_t n_pairs = 0, n_cells = 0;
    const lk_status st = lk_pairhmm_batch_dims(b, &n_pairs, &n_cells);  // validates offsets
    if (st != LK_OK) return st;
    const int64_t G = b->n_regions;
    pair_base.assign((size_t)G + 1, 0);
    std::vector<int64_t> cells((size_t)G, 0);
    for (int64_t g = 0; g < G; g++) {
        const int64_t r0 = b->region_read_start[g], r1 = b->region_read_start[g + 1];
        const int64_t h0 = b->region_hap_start[g], h1 = b->region_hap_start[g + 1];
        pair_base[(size_t)g + 1] = pair_base[(size_t)g] + (r1 - r0) * (h1 - h0);
        if (r1 > r0 && h1 > h0) cells[(size_t)g] = (b->read_off[r1] - b->read_off[r0]) * (b->hap_off[h1] - b->hap_off[h0]);
    }
    const int64_t want = std::max<int64_t>(1, (int64_t)n_dev * chunks_per_device);
    // a chunk below ~1e9 cells (a quarter of a millisecond of one B200) costs more in launch latency than it balances
    // (LK_POOL_MIN_CHUNK_CELLS: test switch, so that small batches exercise the chunk views too)
    const char *env_floor = getenv("LK_POOL_MIN_CHUNK_CELLS");
    const int64_t floor_cells = env_floor ? std::max<int64_t>(1, atoll(env_floor)) : 1000000000LL;
    const int64_t target = std::max<int64_t>((n_cells + want - 1) / want, floor_cells);
    chunks.clear();
    int64_t acc = 0, g0 = 0;
    for (int64_t g = 0; g < G; g++) {
        acc += cells[(size_t)g];
        if (acc >= target || g + 1 == G) { chunks.push_back(Chunk{ g0, g + 1 }); g0 = g + 1; acc = 0; }
    }
    return LK_OK;
}

// Run `call(ctx, sub-batch view, g0)` for every chunk, one host thread per device pulling chunks.
template <class F>
lk_status deal(lk_pool *pool, const lk_pairhmm_batch *b, F call) {
    std::vector<Chunk> chunks;
    std::vector<int64_t> pair_base;
    try {
        lk_status st = cut_chunks(b, (int)pool->ctxs.size(), pool->chunks_per_device, chunks, pair_base);
        if (st != LK_OK) return st;
        if (chunks.empty()) return LK_OK;
        std::atomic<size_t> next{ 0 };
        std::atomic<int> first_error{ (int)LK_OK };
        std::atomic<int> error_dev{ -1 };
        auto worker = [&](size_t d) {
            try {
                std::vector<int64_t> rrs, rhs, rss, srs;
                for (;;) {
                    const size_t i = next.fetch_add(1);
                    if (i >= chunks.size() || first_error.load() != (int)LK_OK) break;
                    const Chunk c = chunks[i];
                    const int64_t r0 = b->region_read_start[c.g0], h0 = b->region_hap_start[c.g0];
                    rrs.resize((size_t)(c.g1 - c.g0) + 1); rhs.resize(rrs.size());
                    for (int64_t g = c.g0; g <= c.g1; g++) {
                        rrs[(size_t)(g - c.g0)] = b->region_read_start[g] - r0;
                        rhs[(size_t)(g - c.g0)] = b->region_hap_start[g] - h0;
                    }
                    lk_pairhmm_batch v = *b;  // a view: arenas are indexed by absolute byte offsets, offsets arrays by read / haplotype
                    v.n_regions = c.g1 - c.g0;
                    v.region_read_start = rrs.data(); v.region_hap_start = rhs.data();
                    v.read_off = b->read_off ? b->read_off + r0 : nullptr;
                    v.hap_off = b->hap_off ? b->hap_off + h0 : nullptr;
                    if (b->region_sample_start && b->sample_read_start) {  // the chunk's samples, re-based like its regions
                        const int64_t s0 = b->region_sample_start[c.g0], s1 = b->region_sample_start[c.g1];
                        rss.resize(rrs.size()); srs.resize((size_t)(s1 - s0) + 1);
                        for (int64_t g = c.g0; g <= c.g1; g++) rss[(size_t)(g - c.g0)] = b->region_sample_start[g] - s0;
                        for (int64_t sm = s0; sm <= s1; sm++) srs[(size_t)(sm - s0)] = b->sample_read_start[sm] - r0;
                        v.region_sample_start = rss.data(); v.sample_read_start = srs.data();
                    }
                    const lk_status cs = call(pool->ctxs[d], &v, c.g0, r0, pair_base[(size_t)c.g0]);
                    if (cs != LK_OK) {
                        int expected = (int)LK_OK;
                        if (first_error.compare_exchange_strong(expected, (int)cs)) error_dev.store((int)d);
                        break;
                    }
                }
            } catch (...) {
                int expected = (int)LK_OK;
                first_error.compare_exchange_strong(expected, (int)LK_ERR_NOMEM);
            }
        };
        const size_t n_threads = std::min(pool->ctxs.size(), chunks.size());
        std::vector<std::thread> th;
        for (size_t d = 1; d < n_threads; d++) th.emplace_back(worker, d);
        worker(0);  // the calling thread drives device 0
        for (auto &t : th) t.join();
        const lk_status fe = (lk_status)first_error.load();
        if (fe != LK_OK) {
            std::lock_guard<std::mutex> g(pool->mu);
            const int d = error_dev.load();
            pool->err = d >= 0 ? std::string("device ") + std::to_string(pool->ctxs[(size_t)d]->device) + ": " + lk_last_error(pool->ctxs[(size_t)d])
                               : std::string("out of host memory");
        }
        return fe;
    } catch (...) {
        return LK_ERR_NOMEM;
    }
}

}  // namespace

extern "C" lk_status lk_pool_compute_read_likelihoods(lk_pool *pool, const lk_engine_args *args, const lk_pairhmm_batch *batch,
                                                      const uint8_t *read_mapq, const int32_t *ref_allele_index, double *values,
                                                      uint8_t *keep) {
    if (!pool || !args || !batch) return LK_ERR_INVALID;
    return deal(pool, batch, [&](lk_ctx *ctx, const lk_pairhmm_batch *v, int64_t g0, int64_t r0, int64_t p0) {
        return lk_engine_compute_read_likelihoods(ctx, args, v, read_mapq ? read_mapq + r0 : nullptr,
                                                  ref_allele_index ? ref_allele_index + g0 : nullptr, values ? values + p0 : nullptr,
                                                  keep ? keep + r0 : nullptr);
    });
}

extern "C" lk_status lk_pool_pairhmm_batch_host(lk_pool *pool, const lk_pairhmm_batch *batch, double *out, uint8_t *flags) {
    if (!pool || !batch) return LK_ERR_INVALID;
    return deal(pool, batch, [&](lk_ctx *ctx, const lk_pairhmm_batch *v, int64_t, int64_t, int64_t p0) {
        return lk_pairhmm_batch_host(ctx, v, out ? out + p0 : nullptr, flags ? flags + p0 : nullptr);
    });
}

extern "C" lk_status lk_pool_batcher_create(lk_pool *pool, const lk_engine_args *args, int32_t max_regions, int32_t max_wait_us,
                                            lk_batcher **out) {
    if (!pool || !out) return LK_ERR_INVALID;
    return lk_batcher_create_multi(pool->ctxs.data(), (int)pool->ctxs.size(), args, max_regions, max_wait_us, out);
}
