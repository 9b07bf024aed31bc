// emul_sw.cpp — TEST TOOLING ONLY: runs the Smith-Waterman warp routine (lorikeet_b200/csrc/sw_core.cuh) in the
// 32-thread lock-step emulator (see emul_pairhmm.cpp).
#include <algorithm>
#include <barrier>
#include <cstring>
#include <thread>
#include <vector>

#include "../../lorikeet_b200/csrc/lk_config.h"
#include "../../lorikeet_b200/csrc/sw16_core.cuh"

using namespace lk;

struct SwShared {
    std::barrier<> bar{32};
    uint64_t slot[32];
};

struct SwEmulWarp {
    int lane_;
    SwShared *sh;
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
    unsigned ballot(bool p) {
        sh->slot[lane_] = p ? 1 : 0;
        sh->bar.arrive_and_wait();
        unsigned r = 0;
        for (int i = 0; i < 32; i++) r |= (unsigned)sh->slot[i] << i;
        sh->bar.arrive_and_wait();
        return r;
    }
    int ffs(unsigned v) const { return __builtin_ffs((int)v); }
    template <class T> T ldg(const T *p) const { return *p; }
    void sync() { sh->bar.arrive_and_wait(); }
    void store4(uint32_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const { p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
    void store2(uint32_t *p, uint32_t a, uint32_t b) const { p[0] = a; p[1] = b; }
    void load2(const uint32_t *p, uint32_t &a, uint32_t &b) const { a = p[0]; b = p[1]; }
    void lds4i(int32_t *dst, const int32_t *src) const { memcpy(dst, src, 16); }
    void lds4u(uint32_t *dst, const uint32_t *src) const { memcpy(dst, src, 16); }
    void load4(const uint32_t *p, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) const { a = p[0]; b = p[1]; c = p[2]; d = p[3]; }
};

static int choose_k(int n, int force_k) {
    if (force_k > 0) return force_k;
    static const int ks[] = { LK_SW_KS };
    for (int k : ks) if (n <= 32 * k) return k;
    return 17;
}

extern "C" int emul_sw_batch(int64_t n_pairs, const int64_t *ref_off, const uint8_t *refs, const int64_t *alt_off,
                             const uint8_t *alts, const int32_t *params, int strategy, int max_ops, int force_k,
                             uint32_t *cigar_ops, int32_t *n_ops, int32_t *offsets) {
    SwArgs A;
    memset(&A, 0, sizeof A);
    A.refs = refs; A.alts = alts; A.ref_off = ref_off; A.alt_off = alt_off;
    A.w_match = params[0]; A.w_mismatch = params[1]; A.w_open = params[2]; A.w_extend = params[3];
    A.strategy = strategy; A.max_ops = max_ops; A.cigar_ops = cigar_ops; A.n_ops = n_ops; A.offsets = offsets;
    A.one = 1u; A.mone = 0xffffffffu; A.two16 = 65536u;
    SwShared sh;
    for (int64_t p = 0; p < n_pairs; p++) {
        const int n = (int)(ref_off[p + 1] - ref_off[p]), m = (int)(alt_off[p + 1] - alt_off[p]);
        const int K = choose_k(n, force_k);
        // same rule as the host launcher; odd pairs take the clamped form so both instantiations are exercised
        const bool clamp = !sw_clamp_inactive((int64_t)n + m, params[0], params[1], params[2], params[3]) || (p & 1);
        std::vector<int32_t> wtab((size_t)SwTable<17>::words);
        std::vector<uint32_t> scratch((size_t)(n > 0 && m > 0 ? sw_scratch_words(n, m, K) : 64));
        std::vector<std::thread> th;
        for (int l = 0; l < 32; l++)
            th.emplace_back([&, l]() {
                SwEmulWarp w{ l, &sh };
                switch (K) {
#define LK_CASE(KV) case KV: if (clamp) sw_align_pair<SwEmulWarp, KV, true>(w, A, (int)p, scratch.data(), wtab.data()); \
                             else sw_align_pair<SwEmulWarp, KV, false>(w, A, (int)p, scratch.data(), wtab.data()); break;
                    LK_SW_K_TABLE(LK_CASE)
#undef LK_CASE
                }
            });
        for (auto &t : th) t.join();
    }
    return 0;
}

// The packed two-pairs-per-warp routine (sw16_core.cuh) with the dispatch of sw_duo_kernel: consecutive pairs (2i, 2i+1) go
// through sw16_align_duo when they qualify, one at a time through sw_dp_pair otherwise.  duo_taken[i] reports which.
extern "C" int emul_sw_duo_batch(int64_t n_pairs, const int64_t *ref_off, const uint8_t *refs, const int64_t *alt_off,
                                 const uint8_t *alts, const int32_t *params, int strategy, int max_ops, int force_k,
                                 uint32_t *cigar_ops, int32_t *n_ops, int32_t *offsets, uint8_t *duo_taken) {
    SwArgs A;
    memset(&A, 0, sizeof A);
    A.refs = refs; A.alts = alts; A.ref_off = ref_off; A.alt_off = alt_off;
    A.w_match = params[0]; A.w_mismatch = params[1]; A.w_open = params[2]; A.w_extend = params[3];
    A.strategy = strategy; A.max_ops = max_ops; A.cigar_ops = cigar_ops; A.n_ops = n_ops; A.offsets = offsets;
    A.one = 1u; A.mone = 0xffffffffu; A.two16 = 65536u;
    if (!sw16_weights_ok(params[0], params[1], params[2], params[3])) return -1;
    SwShared sh;
    for (int64_t p = 0; p < n_pairs; p += 2) {
        const int pa = (int)p, pb = (p + 1 < n_pairs) ? (int)p + 1 : -1;
        int64_t nmax = 0, mmax = 0;
        for (int q : { pa, pb })
            if (q >= 0) { nmax = std::max(nmax, ref_off[q + 1] - ref_off[q]); mmax = std::max(mmax, alt_off[q + 1] - alt_off[q]); }
        const int K = force_k > 0 ? force_k : (nmax <= 256 ? 8 : (nmax <= 384 ? 12 : (nmax <= 512 ? 16 : 17)));
        if (!sw_clamp_inactive(nmax + mmax, params[0], params[1], params[2], params[3])) return -2;
        std::vector<int32_t> wtab((size_t)std::max(SwTable<17>::words, Sw16Table<17>::words));
        std::vector<uint32_t> scratch((size_t)std::max<int64_t>(64, std::max(nmax > 0 && mmax > 0 ? sw_scratch_words((int)nmax, (int)mmax, K) : 64,
                                                                           sw16_scratch_words((int)mmax, K))));
        uint8_t took[32] = { 0 };
        std::vector<std::thread> th;
        for (int l = 0; l < 32; l++)
            th.emplace_back([&, l]() {
                SwEmulWarp w{ l, &sh };
                auto run = [&](auto ktag) {
                    constexpr int KV = decltype(ktag)::value;
                    const bool done_a = sw_prologue(w, A, pa), done_b = (pb < 0) || sw_prologue(w, A, pb);
                    if (!done_a && !done_b) {
                        const int64_t na = ref_off[pa + 1] - ref_off[pa], nb = ref_off[pb + 1] - ref_off[pb];
                        if (na <= 32 * KV && nb <= 32 * KV &&
                            (na == nb ? sw16_align_duo<SwEmulWarp, KV, true>(w, A, pa, pb, scratch.data(), wtab.data())
                                      : sw16_align_duo<SwEmulWarp, KV, false>(w, A, pa, pb, scratch.data(), wtab.data()))) { took[l] = 1; return; }
                    }
                    if (!done_a) { sw_dp_pair<SwEmulWarp, KV, false>(w, A, pa, scratch.data(), wtab.data()); w.sync(); }
                    if (!done_b) { sw_dp_pair<SwEmulWarp, KV, false>(w, A, pb, scratch.data(), wtab.data()); w.sync(); }
                };
                if (K == 8) run(std::integral_constant<int, 8>());
                else if (K == 12) run(std::integral_constant<int, 12>());
                else if (K == 16) run(std::integral_constant<int, 16>());
                else run(std::integral_constant<int, 17>());
            });
        for (auto &t : th) t.join();
        if (duo_taken) { duo_taken[p] = took[0]; if (pb >= 0) duo_taken[pb] = took[0]; }
    }
    return 0;
}
