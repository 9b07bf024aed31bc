// host_parallel.h — a plain fork/join loop for the host-side per-pair work around the GPU batches (padding, CIGAR
// surgery).  The reference runs these per haplotype / per read inside Rayon workers; a batch call does them for the whole
// batch, so they are spread over the host cores here.  Small batches run inline on the calling thread.
#pragma once
#include <algorithm>
#include <cstdint>
#include <exception>
#include <mutex>
#include <thread>
#include <vector>

namespace lk {

// fn(begin, end) over [0, n) in contiguous slices; at least `grain` items per thread, at most 16 threads.
// Nothing unwinds out of a worker thread (that would be std::terminate) and the pool is always joined: the first exception
// of any slice -- in practice std::bad_alloc from a growing container -- is re-thrown on the calling thread after the join,
// where the extern "C" entry points turn it into LK_ERR_NOMEM (LK_ABI_GUARD).
template <class F>
inline void parallel_for(int64_t n, int64_t grain, F &&fn) {
    if (n <= 0) return;
    const int64_t hw = (int64_t)std::max(1u, std::thread::hardware_concurrency());
    const int64_t n_threads = std::min<int64_t>({ hw, 16, (n + grain - 1) / grain });
    if (n_threads <= 1) { fn((int64_t)0, n); return; }
    std::exception_ptr first;
    std::mutex mu;
    auto guarded = [&](int64_t b, int64_t e) {
        try { fn(b, e); } catch (...) {
            std::lock_guard<std::mutex> g(mu);
            if (!first) first = std::current_exception();
        }
    };
    std::vector<std::thread> pool;
    struct Joiner {  // joins whatever was started, also when emplace_back itself throws
        std::vector<std::thread> &p;
        ~Joiner() { for (auto &th : p) if (th.joinable()) th.join(); }
    } joiner{ pool };
    pool.reserve((size_t)n_threads - 1);
    for (int64_t t = 1; t < n_threads; t++) pool.emplace_back([&, t]() { guarded(n * t / n_threads, n * (t + 1) / n_threads); });
    guarded((int64_t)0, n / n_threads);
    for (auto &th : pool) th.join();
    if (first) std::rethrow_exception(first);
}

}  // namespace lk

// Body guard of an extern "C" entry point: no C++ exception crosses the ABI.  `fail` is evaluated with the
// status to return, e.g. `return lk_fail(ctx, st, "...")`.
#define LK_ABI_GUARD_BEGIN try {
#define LK_ABI_GUARD_END(on_nomem, on_other)                   \
    } catch (const std::bad_alloc &) { on_nomem; }             \
    catch (...) { on_other; }
