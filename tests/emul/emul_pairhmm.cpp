// emul_pairhmm.cpp — TEST TOOLING ONLY (never linked into the product library).
//
// A 32-thread lock-step emulator of one CUDA warp: it runs the *same* templated warp routines the
// GPU kernels run (lorikeet_b200/csrc/pairhmm_core.cuh) with shuffles implemented through a shared
// slot array and barriers.  It exists because the build container has no GPU: it lets the CPU test
// suite exercise the lane/strip/skew index logic of the kernels before a gpurun call.
#include <cstdlib>
#include <atomic>
#include <barrier>
#include <cstring>
#include <thread>
#include <vector>

#include "../../lorikeet_b200/csrc/host_model.h"
#include "../../lorikeet_b200/csrc/host_prep.h"
#include "../../lorikeet_b200/csrc/pairhmm_core.cuh"

using namespace lk;

struct WarpShared {
    std::barrier<> bar{32};
    uint64_t slot[32];
};

struct EmulWarp {
    int lane_;
    WarpShared *sh;
    int lane() const { return lane_; }
    template <class T> T exchange(T v, int src) {
        uint64_t bits = 0;
        memcpy(&bits, &v, sizeof(T));
        sh->slot[lane_] = bits;
        sh->bar.arrive_and_wait();
        T r = v;
        if (src >= 0 && src < 32) memcpy(&r, &sh->slot[src], sizeof(T));
        sh->bar.arrive_and_wait();
        return r;
    }
    template <class T> T shfl_up1(T v) { return exchange(v, lane_ - 1); }
    template <class T> T shfl(T v, int src) { return exchange(v, src & 31); }
    template <class T> T shfl_xor(T v, int m) { return exchange(v, lane_ ^ m); }
    unsigned reduce_or(unsigned v) {
        sh->slot[lane_] = v;
        sh->bar.arrive_and_wait();
        unsigned r = 0;
        for (int i = 0; i < 32; i++) r |= (unsigned)sh->slot[i];
        sh->bar.arrive_and_wait();
        return r;
    }
    unsigned reduce_max(unsigned v) {
        sh->slot[lane_] = v;
        sh->bar.arrive_and_wait();
        unsigned r = 0;
        for (int i = 0; i < 32; i++) r = std::max(r, (unsigned)sh->slot[i]);
        sh->bar.arrive_and_wait();
        return r;
    }
    int reduce_min_i(int v) {
        sh->slot[lane_] = (uint64_t)(int64_t)v;
        sh->bar.arrive_and_wait();
        int r = v;
        for (int i = 0; i < 32; i++) r = std::min(r, (int)(int64_t)sh->slot[i]);
        sh->bar.arrive_and_wait();
        return r;
    }
    template <class T> T ldg(const T *p) const { return *p; }
    void lds4(float *dst, const float *src) const { memcpy(dst, src, 16); }
    int atomic_add(int32_t *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
    void sync() { sh->bar.arrive_and_wait(); }
    void prefetch(const void *) const {}
};

template <int K>
static void run_fast_tasks(const PairHmmArgs &A, bool vargcp, const WarpTask *tasks, int n_tasks, const TaskHaps *task_haps) {
    PairHmmArgs AL = A;
    AL.tasks = tasks;
    AL.n_tasks = n_tasks;
    AL.task_haps = task_haps;
    WarpShared sh;
    std::vector<float> tab((size_t)fast_table_floats(K));  // stands in for the warp's shared-memory prior table
    std::vector<std::thread> th;
    for (int l = 0; l < 32; l++)
        th.emplace_back([&, l]() {
            EmulWarp w{ l, &sh };
            for (int t = 0; t < n_tasks; t++) {
                if (vargcp) pairhmm_fast_task<EmulWarp, K, GM_VAR>(w, AL, t, tab.data());
                else if (AL.gcp_constant == 10 && AL.cf_uniform == 0.1f) {
                    if (AL.task_haps) pairhmm_fast_task<EmulWarp, K, GM_Q10, true>(w, AL, t, tab.data());
                    else pairhmm_fast_task<EmulWarp, K, GM_Q10>(w, AL, t, tab.data());
                }
                else pairhmm_fast_task<EmulWarp, K, GM_UNIFORM>(w, AL, t, tab.data());
            }
        });
    for (auto &t : th) t.join();
}

static void run_band(const PairHmmArgs &A, int k_band) {
    WarpShared sh;
    std::vector<float> tab((size_t)band_smem_floats(LK_BAND_K > 4 ? LK_BAND_K : 4));
    Bnd4 ring[BAND_RING];
    std::vector<std::thread> th;
    for (int l = 0; l < 32; l++)
        th.emplace_back([&, l]() {
            EmulWarp w{ l, &sh };
            for (int i = 0; i < A.n_band; i++) {
                if (k_band == 4) pairhmm_band_pair<EmulWarp, 4>(w, A, A.band_pairs[i].read, (int64_t)A.band_pairs[i].hap, (Bnd4 *)A.band_scratch, tab.data(), ring);
                else pairhmm_band_pair<EmulWarp, LK_BAND_K>(w, A, A.band_pairs[i].read, (int64_t)A.band_pairs[i].hap, (Bnd4 *)A.band_scratch, tab.data(), ring);
            }
        });
    for (auto &t : th) t.join();
}

static void run_plain(const PairHmmArgs &A, int n) {
    WarpShared sh;
    std::vector<std::thread> th;
    for (int l = 0; l < 32; l++)
        th.emplace_back([&, l]() {
            EmulWarp w{ l, &sh };
            for (int i = 0; i < n; i++)
                pairhmm_plain_pair<EmulWarp, LK_PLAIN_K>(w, A, A.rescue_pairs[i].read, (int64_t)A.rescue_pairs[i].hap, A.scratch);
        });
    for (auto &t : th) t.join();
}

// number of warp tasks build_prep makes for a batch (split_target_tasks as in build_prep); lets the tests see the split
extern "C" int64_t emul_count_fast_tasks(const lk_pairhmm_batch *b, int64_t split_target_tasks) {
    HostPrep P;
    build_prep(b, 0, b->n_regions, 0, true, P, split_target_tasks);
    return (int64_t)P.tasks.size();
}

// Same contract as lk_pairhmm_batch_host, executed by the emulator.  force_k > 0 pins K; force_f64 sends
// every pair through the plain-form pass.  n_rescued (may be null) receives the f64-pass pair count.
extern "C" int emul_pairhmm_batch(const lk_pairhmm_batch *b, int use_tristate, int force_k, int force_f64,
                                  double *out, uint8_t *flags, int32_t *n_rescued) {
    const int64_t n_reads = b->region_read_start[b->n_regions], n_haps = b->region_hap_start[b->n_regions];
    HostPrep P;
    const char *split = getenv("LK_EMUL_SPLIT_TASKS");  // exercise the haplotype-axis task split of small launches
    build_prep(b, 0, b->n_regions, force_k, !force_f64, P, split ? atoll(split) : 0);
    if (force_k == -4) { HostPrep Q; build_prep(b, 0, b->n_regions, -1, !force_f64, Q); P = Q; }
    ModelTables T;
    build_model_tables(T);
    std::vector<uint8_t> hidx((size_t)P.henc_words, 0);
    std::vector<uint8_t> hap_flags((size_t)n_haps, 0), read_flags((size_t)n_reads, 0);
    for (int64_t h = 0; h < n_haps; h++) {  // = encode_haps_kernel
        bool ex = false;
        for (int64_t j = b->hap_off[h]; j < b->hap_off[h + 1]; j++)
        {
            (void)encode_base(b->hap_bases[j], ex);
            hidx[(size_t)(P.henc_off[(size_t)h] + (j - b->hap_off[h]))] = base_slice(b->hap_bases[j]);
        }
        hap_flags[(size_t)h] = ex ? SEQ_EXOTIC : 0;
    }
    bool vargcp = false;
    uint8_t gcp_const = b->gcp_constant;
    if (b->read_gcp) {
        gcp_const = b->read_gcp[0];
        for (int64_t i = 0; i < b->read_off[n_reads]; i++) if (b->read_gcp[i] != gcp_const) vargcp = true;
    }
    for (int64_t r = 0; r < n_reads; r++) {  // = scan_reads_kernel
        bool ex = false, qz = false;
        int imin = 255, imax = 0;
        for (int64_t i = b->read_off[r]; i < b->read_off[r + 1]; i++) {
            (void)encode_base(b->read_bases[i], ex);
            qz |= b->read_qual[i] == 0 || b->read_ins[i] < 4 || b->read_del[i] < 4;
            imin = std::min<int>(imin, b->read_ins[i]);
            imax = std::max<int>(imax, b->read_ins[i]);
        }
        read_flags[(size_t)r] = (ex ? SEQ_EXOTIC : 0) | ((qz || imax - imin > FAST_MAX_INS_RANGE) ? SEQ_QUAL_DOMAIN : 0);
    }
    std::vector<RescuePair> rescue((size_t)std::max<int64_t>(P.n_pairs, 1));
    int32_t counters[2] = { 0, 0 };
    std::vector<double> scratch((size_t)(6 * (P.h_max + 1)));
    PairHmmArgs A;
    memset(&A, 0, sizeof A);
    A.read_bases = b->read_bases; A.read_qual = b->read_qual; A.read_ins = b->read_ins; A.read_del = b->read_del;
    A.read_gcp = vargcp ? b->read_gcp : nullptr;
    A.read_off = b->read_off; A.read_region = P.read_region.data(); A.read_out_base = P.read_out_base.data();
    A.read_flags = read_flags.data(); A.region_hap_start = b->region_hap_start; A.hap_off = b->hap_off;
    A.hap_bases = b->hap_bases; A.hidx = hidx.data(); A.henc_off = P.henc_off.data(); A.hap_flags = hap_flags.data();
    A.qerr_f = T.qerr_f.data(); A.qerr_d = T.qerr_d.data(); A.m2m = T.m2m.data();
    A.groups = P.groups.data(); A.out = out; A.flags = flags;
    A.rescue_count = &counters[0]; A.rescue_cursor = &counters[1];
    A.rescue_pairs = rescue.data(); A.rescue_capacity = (int32_t)P.n_pairs;
    A.scratch = scratch.data(); A.scratch_stride = (int64_t)scratch.size();
    A.gcp_constant = gcp_const;
    A.mismatch_div = use_tristate ? 3.0f : 1.0f;
    A.cf_uniform = T.qerr_f[gcp_const];
    if (force_f64) {
        for (int64_t r = 0; r < n_reads; r++) {
            const int32_t g = P.read_region[(size_t)r];
            for (int64_t h = b->region_hap_start[g]; h < b->region_hap_start[g + 1]; h++)
                rescue[(size_t)counters[0]++] = RescuePair{ (int32_t)r, (int32_t)h };
        }
    } else {
        // long reads: fp32 band routine (force_k == -4 runs it with K = 4 so that short test reads span many bands)
        std::vector<Bnd4> band_scratch((size_t)(2 * (P.h_max + 1)));
        int32_t band_cursor = 0;
        if (!vargcp) {
            A.band_pairs = P.direct.data(); A.n_band = (int32_t)P.direct.size();
            A.band_cursor = &band_cursor; A.band_scratch = band_scratch.data(); A.band_scratch_stride = (int64_t)band_scratch.size();
            run_band(A, force_k == -4 ? 4 : LK_BAND_K);
        } else {
            for (const auto &d : P.direct) rescue[(size_t)counters[0]++] = d;
        }
        for (const auto &L : P.launches) {
            const WarpTask *t = P.tasks.data() + L.first_task;
            switch (L.K) {
#define LK_CASE(KV, MINB) case KV: run_fast_tasks<KV>(A, vargcp, t, L.n_tasks, L.split ? P.task_haps.data() + L.first_task : nullptr); break;
                LK_FAST_K_TABLE(LK_CASE)
#undef LK_CASE
                default: return -1;
            }
        }
    }
    if (n_rescued) *n_rescued = counters[0];
    run_plain(A, counters[0]);
    return 0;
}

// Test hook: the six transition probabilities of one read position as the KERNELS see them, i.e. read out of the tables
// host_model.cpp builds and lk_create uploads (qerr_d, triangular m2m) -- order of pair_hmm_model.rs:88-113:
// match->match, indel->match, match->insertion, insertion->insertion, match->deletion, deletion->deletion.
extern "C" void emul_model_trans_probs(int ins, int del, int gcp, double out[6]) {
    static ModelTables T;
    static bool built = false;
    if (!built) { build_model_tables(T); built = true; }
    out[0] = m2m_lookup(T.m2m.data(), ins, del);
    out[1] = 1.0 - T.qerr_d[(size_t)gcp];
    out[2] = T.qerr_d[(size_t)ins];
    out[3] = T.qerr_d[(size_t)gcp];
    out[4] = T.qerr_d[(size_t)del];
    out[5] = T.qerr_d[(size_t)gcp];
}
