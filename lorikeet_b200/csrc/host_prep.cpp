// host_prep.cpp — K selection and warp-task packing for the PairHMM fast path (host side).
//
// A warp task = reads of ONE region packed side by side into the 32 lanes (each read takes
// ceil(R/K) lanes); the kernel then sweeps every haplotype of that region, i.e. the same
// "for read { for haplotype }" order as the reference driver (src/pair_hmm/pair_hmm.rs:349-369).
#include "host_prep.h"

#include <algorithm>

namespace lk {

const int *fast_k_values(int *n) {
    static const int ks[] = { LK_FAST_KS };
    *n = (int)(sizeof(ks) / sizeof(ks[0]));
    return ks;
}

// Pick rows-per-lane for a read of length R: maximise useful cell-updates per issue slot.
//   lanes per read G = ceil(R/K); reads per warp p = floor(32/G);
//   row efficiency = p*R / (32*K); skew efficiency ~ H/(H+G-1) with H ~ 2R;
//   issued instructions per cell ~ 5.25 (FMUL + 4 FFMA + a quarter LDS.128) + 14/K of per-step bookkeeping
//   (measured on B200 for R = 150: K = 19 -> 4 450 GCUPS, K = 20 -> 4 220, K = 15 -> 3 890, K = 16 -> 3 630).
int choose_k(int R, int force_k) {
    int n = 0;
    const int *ks = fast_k_values(&n);
    if (force_k < 0) return 0;  // developer switch: route every read to the band kernel
    if (force_k > 0) {
        for (int i = 0; i < n; i++)
            if (ks[i] == force_k && R <= 32 * force_k) return force_k;
    }
    double best = -1.0;
    int best_k = 0;
    for (int i = 0; i < n; i++) {
        const int K = ks[i];
        if (R > 32 * K) continue;
        const int G = (R + K - 1) / K;
        const int p = 32 / G;
        const double row_eff = (double)(p * R) / (32.0 * K);
        const double H = std::max(2.0 * R, 64.0);
        const double skew = H / (H + G - 1);
        const double score = row_eff * skew / (5.25 + 14.0 / K);
        if (score > best) { best = score; best_k = K; }
    }
    return best_k;  // 0: too long for the fast path
}


void build_prep(const lk_pairhmm_batch *b, int64_t g0, int64_t g1, int force_k, bool fast_enabled, HostPrep &P,
                int64_t split_target_tasks) {
    // Everything produced here is LOCAL to the region range [g0, g1): read / haplotype / region indices and
    // result offsets start at 0, so a chunk can be staged and run on its own.
    const int64_t rbase = b->region_read_start[g0], hbase = b->region_hap_start[g0];
    const int64_t n_reads = b->region_read_start[g1] - rbase, n_haps = b->region_hap_start[g1] - hbase;
    P = HostPrep();
    P.read_region.resize((size_t)n_reads);
    P.read_out_base.resize((size_t)n_reads);
    P.henc_off.resize((size_t)n_haps);
    int64_t w = 0;
    for (int64_t h = 0; h < n_haps; h++) {
        const int H = (int)(b->hap_off[hbase + h + 1] - b->hap_off[hbase + h]);
        w = (w + 15) & ~(int64_t)15;  // every encoded haplotype (left guard band included) starts 16-byte aligned
        P.henc_off[(size_t)h] = w + HENC_PADL;
        w += H + HENC_PADL + HENC_PADR;
        P.h_max = std::max(P.h_max, H);
    }
    P.henc_words = w;
    int nk = 0;
    const int *ks = fast_k_values(&nk);
    std::vector<int8_t> read_k((size_t)n_reads, 0);
    std::vector<int> k_by_len;  // memo: read length -> K
    bool k_present[64] = { false };
    int64_t base = 0;
    for (int64_t g = g0; g < g1; g++) {
        const int64_t r0 = b->region_read_start[g], r1 = b->region_read_start[g + 1];
        const int64_t h0 = b->region_hap_start[g], h1 = b->region_hap_start[g + 1];
        const int64_t A = h1 - h0;
        for (int64_t r = r0; r < r1; r++) {
            P.read_region[(size_t)(r - rbase)] = (int32_t)(g - g0);
            P.read_out_base[(size_t)(r - rbase)] = base + (r - r0) * A;
            const int R = (int)(b->read_off[r + 1] - b->read_off[r]);
            int K = 0;
            if (fast_enabled) {
                if ((size_t)R >= k_by_len.size()) k_by_len.resize((size_t)R + 1, -1);
                if (k_by_len[(size_t)R] < 0) k_by_len[(size_t)R] = choose_k(R, force_k);
                K = k_by_len[(size_t)R];
            }
            read_k[(size_t)(r - rbase)] = (int8_t)K;
            k_present[K] = true;
            if (K == 0 && fast_enabled)
                for (int64_t h = h0; h < h1; h++) P.direct.push_back(RescuePair{ (int32_t)(r - rbase), (int32_t)(h - hbase) });
        }
        base += (r1 - r0) * A;
        if (r1 > r0 && h1 > h0) P.n_cells += (b->read_off[r1] - b->read_off[r0]) * (b->hap_off[h1] - b->hap_off[h0]);
    }
    P.n_pairs = base;
    if (!fast_enabled) return;
    P.groups.reserve((size_t)n_reads);
    // one launch per K; inside a launch tasks follow region order so neighbouring warps share haplotypes.
    // Reads of a region (same K) are packed longest first so that lanes of a warp finish together.
    std::vector<int64_t> bucket;
    for (int ki = 0; ki < nk; ki++) {
        const int K = ks[ki];
        if (!k_present[K]) continue;
        FastLaunch L;
        L.K = K;
        L.split = false;
        L.first_task = (int32_t)P.tasks.size();
        for (int64_t g = g0; g < g1; g++) {
            const int64_t r0 = b->region_read_start[g], r1 = b->region_read_start[g + 1];
            if (b->region_hap_start[g + 1] == b->region_hap_start[g]) continue;
            bucket.clear();
            for (int64_t r = r0; r < r1; r++)
                if (read_k[(size_t)(r - rbase)] == K) bucket.push_back(r);
            if (bucket.empty()) continue;
            auto len = [&](int64_t r) { return (int)(b->read_off[r + 1] - b->read_off[r]); };
            std::stable_sort(bucket.begin(), bucket.end(), [&](int64_t x, int64_t y) { return len(x) > len(y); });
            int lane = 0;
            const TaskHaps whole{ 0, (int32_t)(b->region_hap_start[g + 1] - b->region_hap_start[g]) };
            WarpTask t{ (int32_t)P.groups.size(), 0 };
            for (size_t i = 0; i < bucket.size(); i++) {
                const int64_t rx = bucket[i];
                const int G = (len(rx) + K - 1) / K;
                if (lane + G > 32) {
                    P.tasks.push_back(t); P.task_haps.push_back(whole);
                    t = WarpTask{ (int32_t)P.groups.size(), 0 };
                    lane = 0;
                }
                P.groups.push_back(TaskGroup{ (int32_t)(rx - rbase), (int16_t)lane, (int16_t)G });
                t.n_groups++;
                lane += G;
            }
            if (t.n_groups > 0) { P.tasks.push_back(t); P.task_haps.push_back(whole); }
        }
        L.n_tasks = (int32_t)P.tasks.size() - L.first_task;
        // A small launch (a few regions from the cross-thread batcher, say) leaves most warp slots empty while each task
        // walks all haplotypes of its region: cut the tasks along the haplotype axis.  The per-row constants and the prior
        // table are rebuilt per slice (a few % of a 4-haplotype sweep), so large launches stay whole.
        if (split_target_tasks > 0 && L.n_tasks > 0 && L.n_tasks < split_target_tasks) {
            const int64_t parts_wanted = (split_target_tasks + L.n_tasks - 1) / L.n_tasks;
            std::vector<WarpTask> cut_t;
            std::vector<TaskHaps> cut_h;
            for (int32_t i = L.first_task; i < L.first_task + L.n_tasks; i++) {
                const WarpTask t = P.tasks[(size_t)i];
                const int32_t A = P.task_haps[(size_t)i].count;
                const int64_t parts = std::max<int64_t>(1, std::min<int64_t>(parts_wanted, A / 4));
                for (int64_t q = 0; q < parts; q++) {
                    const int32_t lo = (int32_t)(A * q / parts), hi = (int32_t)(A * (q + 1) / parts);
                    cut_t.push_back(t);
                    cut_h.push_back(TaskHaps{ lo, hi - lo });
                }
            }
            if ((int32_t)cut_t.size() > L.n_tasks) {
                P.tasks.resize((size_t)L.first_task); P.task_haps.resize((size_t)L.first_task);
                P.tasks.insert(P.tasks.end(), cut_t.begin(), cut_t.end());
                P.task_haps.insert(P.task_haps.end(), cut_h.begin(), cut_h.end());
                L.n_tasks = (int32_t)cut_t.size();
                L.split = true;
            }
        }
        if (L.n_tasks > 0) P.launches.push_back(L);
    }
}

}  // namespace lk
