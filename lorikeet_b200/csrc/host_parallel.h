// host_parallel.h — a plain fork/join loop for the host-side per-pair work around the GPU batches (padding, CIGAR
// surgery).  The reference runs these per haplotype / per read inside Rayon workers; a batch call does them for the whole
// batch, so they are spread over the host cores here.  Small batches run inline on the calling thread.
#pragma once
#include <algorithm>
#include <cstdint>
#include <thread>
#include <vector>

namespace lk {

// fn(begin, end) over [0, n) in contiguous slices; at least `grain` items per thread, at most 16 threads
template <class F>
inline void parallel_for(int64_t n, int64_t grain, F &&fn) {
    if (n <= 0) return;
    const int64_t hw = (int64_t)std::max(1u, std::thread::hardware_concurrency());
    const int64_t n_threads = std::min<int64_t>({ hw, 16, (n + grain - 1) / grain });
    if (n_threads <= 1) { fn((int64_t)0, n); return; }
    std::vector<std::thread> pool;
    pool.reserve((size_t)n_threads - 1);
    for (int64_t t = 1; t < n_threads; t++) pool.emplace_back([&, t]() { fn(n * t / n_threads, n * (t + 1) / n_threads); });
    fn((int64_t)0, n / n_threads);
    for (auto &th : pool) th.join();
}

}  // namespace lk
