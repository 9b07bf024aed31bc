// pairhmm_core.cuh — PairHMM forward wavefront, warp-level routines.
//
// Computes what the reference's forward recurrence computes
// (src/pair_hmm/pair_hmm.rs:503-615 `sub_compute_read_likelihood_given_haplotype_log10`,
//  priors :626-654, transitions src/pair_hmm/pair_hmm_model.rs:142-156), re-designed for a
// 32-lane warp:
//
//   * each lane owns a strip of K consecutive READ rows in registers; the warp sweeps the
//     haplotype columns as a skewed wavefront (lane l works on column s-l at step s), handing the
//     strip's last row (M, I, D) to the next lane with one __shfl_up_sync per value;
//   * per-row constants (transition x prior products) are built once per read and reused for every
//     haplotype of the region;
//   * "scaled hat form" (fp32 fast path): I^ = I / tMI[i], D^ = D / tMD[i] turn the I and D updates into
//     one FMA each, the match prior is folded into the row constants and each row's state carries the
//     next row's match->match factor, so a cell costs 7 issue slots (compare, select, FMUL, 4 FFMA)
//     for the 12 flops the reference spends (pair_hmm.rs:575-591);
//   * "plain form" (f64 pass): the reference's own operation order, initial condition 2^1020,
//     un-fused multiplies/adds, any read / haplotype length (row bands with power-of-two
//     renormalisation between bands).
//
// The routines are templated on a Warp policy so that the identical code runs (a) on the GPU
// (DevWarp: shuffles) and (b) in tests/emul, a 32-thread lock-step host emulator used to debug
// index logic where no GPU is present.  (b) is test tooling only and is never linked into the product.
#pragma once
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LK_DEV __device__ __forceinline__
#else
#define LK_DEV inline
#endif

namespace lk {

// ---- encoded sequences ----------------------------------------------------------------------
// Haplotypes are re-coded by a prologue kernel to one byte per column: the slice of the per-warp prior
// table its base selects (base_slice), with zero guard bands either side so lanes can read columns <1
// and >H unconditionally.  Read bases become a one-hot mask (encode_base: A=01 C=02 G=04 T=08, N=FF so
// that it matches anything) when the prior table is built.  Any other byte flags the sequence
// SEQ_EXOTIC, which routes its pairs to the byte-exact f64 pass.
constexpr int HENC_PADL = 48;  // a multiple of 16: every encoded haplotype starts 16-byte aligned (host_prep.cpp), so a bulk copy can stage it
constexpr int HENC_PADR = 44;

// Slice of the per-warp prior table (fast kernel) a haplotype base selects.
constexpr int N_PRIOR_SLICES = 5;
LK_DEV uint8_t base_slice(uint8_t b) {
    switch (b) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        case 'N': return 4;
        default: return 0;  // exotic bytes flag the haplotype; its pairs are redone in f64
    }
}
// floats of per-warp shared memory pairhmm_fast_task<K> needs
// LK_FAST_TMA = 1 (experiment, see DESIGN.md "why no TMA staging"): the fast kernel stages every encoded haplotype into
// per-warp shared memory with cp.async.bulk + an mbarrier (double-buffered: haplotype h+1 is in flight while h is swept)
// and reads its per-step table slice with LDS.U8 instead of LDG.U8.
#ifndef LK_FAST_TMA
#define LK_FAST_TMA 0
#endif
constexpr int FAST_TMA_BUF = 512;  // bytes per staged haplotype (pads included): H <= 420, enough for cfg#2
constexpr int FAST_TMA_FLOATS = LK_FAST_TMA ? (2 * FAST_TMA_BUF + 16) / 4 : 0;
template <int K> struct FastTable { static constexpr int floats = N_PRIOR_SLICES * ((K + 3) / 4) * 32 * 4 + FAST_TMA_FLOATS; };
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr int fast_table_floats(int K) { return N_PRIOR_SLICES * ((K + 3) / 4) * 32 * 4 + FAST_TMA_FLOATS; }

LK_DEV uint32_t encode_base(uint8_t b, bool &exotic) {
    uint32_t m;
    switch (b) {
        case 'A': m = 0x01u; break;
        case 'C': m = 0x02u; break;
        case 'G': m = 0x04u; break;
        case 'T': m = 0x08u; break;
        case 'N': m = 0xFFu; break;
        default: m = 0u; exotic = true; break;
    }
    return m * 0x01010101u;
}

// fp32 fast path: linear space pre-scaled by 2^100 (the reference uses 2^1020 in f64, pair_hmm.rs:16);
// 2^28 of headroom is left for the hat variables (bounded by the spread of insertion qualities).
constexpr float FAST_INIT = 1.2676506002282294e30f;          // 2^100
constexpr double FAST_INIT_LOG10 = 30.102999566398120;        // 100*log10(2)
constexpr float FAST_MIN_ACCEPTED = 7.7037197775489434e-34f;  // 2^-110; smaller sums are redone in f64
constexpr int FAST_MAX_INS_RANGE = 60;                        // phred spread of ins quals the hat form tolerates

enum : uint8_t { SEQ_EXOTIC = 1, SEQ_QUAL_DOMAIN = 2 };
// per-pair result flags (same values as LK_FLAG_* in include/lorikeet_b200.h)
enum : uint8_t { PAIR_RESCUED = 1, PAIR_NEG_INF = 2 };

struct TaskGroup {       // one read inside a warp task
    int32_t read;        // read index (local to the staged chunk)
    int16_t lane_start;  // first lane of the group
    int16_t lanes;       // ceil(R / K)
};
struct WarpTask {        // all groups of a task belong to one region
    int32_t first_group;
    int32_t n_groups;
};
struct TaskHaps {        // haplotype slice of a split task, relative to its region's first haplotype
    int32_t begin;
    int32_t count;
};
struct RescuePair {
    int32_t read;
    int32_t hap;         // batch-global haplotype index
};

struct PairHmmArgs {
    // reads (arenas as given by the caller)
    const uint8_t *read_bases, *read_qual, *read_ins, *read_del, *read_gcp;  // read_gcp may be null
    const int64_t *read_off;
    const int32_t *read_region;
    const int64_t *read_out_base;  // index in out[] of (read, first haplotype of its region)
    const uint8_t *read_flags;
    // haplotypes
    const int64_t *region_hap_start;
    const int64_t *hap_off;
    const uint8_t *hap_bases;
    const uint8_t *hidx;           // per haplotype column: prior-table slice of its base (A C G T N = 0..4), guard bands = 0
    const int64_t *henc_off;       // index of column 1 of haplotype h inside hidx
    const uint8_t *hap_flags;
    // tables
    const float *qerr_f;           // 10^(-q/10), q = 0..255
    const double *qerr_d;
    const double *m2m;             // triangular match->match table, (max*(max+1)/2 + min)
    // fast-path tasks
    const WarpTask *tasks;
    const TaskGroup *groups;
    int32_t n_tasks;
    // results
    double *out;
    uint8_t *flags;
    int32_t *rescue_count;
    RescuePair *rescue_pairs;
    int32_t rescue_capacity;
    // band path (long reads, fp32)
    const RescuePair *band_pairs;
    int32_t n_band;
    int32_t *band_cursor;
    void *band_scratch;            // per-warp boundary rows (Bnd4)
    int64_t band_scratch_stride;   // elements per warp
    // f64 pass
    double *scratch;               // per-warp band boundaries
    int64_t scratch_stride;        // doubles per warp
    int32_t *rescue_cursor;
    uint8_t gcp_constant;
    float mismatch_div;            // 3 (tristate, pair_hmm.rs:53) or 1
    float cf_uniform;              // 10^(-gcp_constant/10) = tII = tDD when gcp is uniform.  Lives in the kernel-parameter
                                   // constant bank so FFMAs can take it as a c[][] operand: a 3-register-source FFMA issues at
                                   // only ~2/3 rate on B200 (tools/microbench/ffma3.cu), a 2-register one at full rate.
    // small launches only (SPLIT instantiations): per task, the slice of its region's haplotypes it sweeps.  Kept out
    // of WarpTask and at the end of the struct: the whole-region kernels must compile exactly as without it (two more
    // task fields cost the cfg#2 kernel 5 % through register allocation).
    const TaskHaps *task_haps;
};

LK_DEV int lk_popc(unsigned v) {
#if defined(__CUDA_ARCH__)
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}


#ifndef LK_FAST_UNROLL
#define LK_FAST_UNROLL 8
#endif
#ifndef LK_BAND_UNROLL
#define LK_BAND_UNROLL 4
#endif
#ifndef LK_BAND_BLOCK
#define LK_BAND_BLOCK 16  // steps per lean block of the band kernel (= block-floating-point upkeep cadence)
#endif
#ifndef LK_BAND_ORDER
#define LK_BAND_ORDER 0  // statement order of the band kernel's cell (ptxas schedules them differently)
#endif
#define LK_PRAGMA_(x) _Pragma(#x)
#define LK_PRAGMA(x) LK_PRAGMA_(x)
#define LK_UNROLL_N(n) LK_PRAGMA(unroll n)

// match->match, pair_hmm_model.rs:442-461 (table) — quals are bytes so max <= 255; 255 > MAX_QUAL uses
// the closed form of :453-458 which the host folded into row 255 of the table.
LK_DEV double m2m_lookup(const double *m2m, int ins, int del) {
    int mn = ins < del ? ins : del, mx = ins < del ? del : ins;
    return m2m[((mx * (mx + 1)) >> 1) + mn];
}

// =============================================================================================
// Fast path: fp32 hat form, one warp task = up to 32 lanes of read groups, every haplotype of the
// region in turn.  Requires R <= 32*K for every read of the task.
// =============================================================================================
//
// Match/mismatch prior WITHOUT per-cell compare + select: LOP3->predicate and FSEL issue on the half-rate ALU pipe
// and cost ~1.8 cycles each next to ~1.05 for an FMA-pipe instruction (tools/microbench/pipes.cu: 5 FMA + LOP3.P +
// FSEL = 8.4 cycles, 5 FMA = 5.3).  Instead each lane keeps, in per-warp shared memory, the finished multiplier
// s[i] * (match ? 1 : r[i]) of each of its K rows against each possible haplotype base (A C G T N), built once per
// task and reused for every haplotype and column:  tab[slice][k/4][lane] is a float4, so a step fetches its K
// multipliers with ceil(K/4) conflict-free LDS.128 at an address picked by the column's base, and the cell is
// FMUL + 4 FFMA.
// GM (gap-continuation mode): GM_UNIFORM = one gcp for the whole batch (factor in a register / the parameter
// bank), GM_VAR = per-base gcp array, GM_Q10 = uniform gcp of 10, the engine's default (`vec![gcp; R]`,
// pair_hmm_likelihood_calculation_engine.rs:649-651): the factor 0.1f is then an FFMA immediate, which saves
// one register-file read per cell in a kernel that is register-read bound.
enum { GM_UNIFORM = 0, GM_VAR = 1, GM_Q10 = 2 };
template <class W, int K, int GM, bool SPLIT = false>
LK_DEV void pairhmm_fast_task(W &w, const PairHmmArgs &A, int task_id, float *tab) {
    constexpr bool VARGCP = (GM == GM_VAR);
    const int lane = w.lane();
    const WarpTask t = A.tasks[task_id];
    TaskGroup g;
    g.read = -1; g.lane_start = 0; g.lanes = 0;
    if (lane < t.n_groups) g = A.groups[t.first_group + lane];
    const unsigned start_mask = w.reduce_or(lane < t.n_groups ? (1u << g.lane_start) : 0u);
    const int maxlanes = (int)w.reduce_max((unsigned)g.lanes);
    const int my_g = lk_popc(start_mask & (0xFFFFFFFFu >> (31 - lane))) - 1;
    const int read = w.shfl(g.read, my_g);
    const int lstart = w.shfl((int)g.lane_start, my_g);
    const int glanes = w.shfl((int)g.lanes, my_g);
    const int lig = lane - lstart;            // lane in group; lanes >= glanes are idle padding
    const bool is_start = (lig == 0);

    const int64_t ro = A.read_off[read];
    const int R = (int)(A.read_off[read + 1] - ro);
    const uint8_t rflag = A.read_flags[read];

    // ---- per-row constants (scaled hat form), built in f64 and rounded once --------------------
    // True recurrences in hat variables (I^ = I/tMI[i], D^ = D/tMD[i], prior folded in, a = pm*tMM, b = pm*tGM*tMI[i-1],
    // c = pm*tGM*tMD[i-1], e = tII*tMI[i-1]/tMI[i], r = px/pm):
    //     mm[i]  = a[i] M[i-1] + b[i] I^[i-1] + c[i] D^[i-1]      (previous column)
    //     M[i]   = mm[i] * (match ? 1 : r[i])
    //     D^[i]  = f D^[i] + M[i]                                  (previous column)
    //     I^[i]  = e[i] I^[i-1] + M[i-1]                           (same column)
    // Row i's state is additionally stored scaled: M' = s[i] M, D' = s[i] D^ with s[i] = a~[i+1], and I' = a~[i] I^
    // (a~ = a for real rows >= 2, 1 for row 1 and for padding rows).  With that choice the M[i-1] term of mm and of
    // I^ has coefficient 1 and is absorbed as an FMA addend:
    //     mm    = fma(D'[i-1], c', fma(I'[i-1], b', M'[i-1]))      b' = b/a~[i-1], c' = c/a~[i]
    //     M'[i] = mm * (match ? s[i] : s[i] r[i])
    //     D'[i] = fma(D'[i], f, M'[i]),   I'[i] = fma(I'[i-1], e', M'[i-1]),   e' = a~[i] e / a~[i-1]
    // i.e. 7 issue slots per cell: compare (LOP3 -> predicate), select, FMUL, 4 FFMA.
    constexpr int KQ = (K + 3) / 4;
    w.sync();  // the table may still be in use by the warp's previous task
    float cb[K], cc[K], ce[K];
    float cf[VARGCP ? K : 1];
    float kcoef = 0.f, wM = 0.f, wI = 0.f;
    if (!VARGCP) cf[0] = A.qerr_f[A.gcp_constant];
    // a~ of a 0-based read position: pm*tMM for real rows >= 1, else 1
    auto a_tilde = [&](int i) -> double {
        if (i < 1 || i >= R) return 1.0;
        return (1.0 - A.qerr_d[A.read_qual[ro + i]]) * m2m_lookup(A.m2m, A.read_ins[ro + i], A.read_del[ro + i]);
    };
#pragma unroll
    for (int k = 0; k < K; k++) {
        const int i = lig * K + k;  // 0-based read position of this row
        float b = 0.f, c = 0.f, e = (i > R) ? 1.f : 0.f, sm = 0.f, sx = 0.f, f = 1.f;
        uint32_t mask = 0u;
        if (i < R) {
            const int q = A.read_qual[ro + i], qi = A.read_ins[ro + i];
            const int qg = A.read_gcp ? A.read_gcp[ro + i] : A.gcp_constant;
            const double err = A.qerr_d[q];
            const double pm = 1.0 - err;                         // pair_hmm.rs:644
            const double px = err / (double)A.mismatch_div;       // :646-651
            const double tGM = 1.0 - A.qerr_d[qg];                // pair_hmm_model.rs:152
            const double tII = A.qerr_d[qg];                      // :153-155
            const double s_i = a_tilde(i + 1);
            bool ex = false;
            mask = encode_base(A.read_bases[ro + i], ex) & 0xFFu;
            if (i == 0) {
                kcoef = (float)(pm * tGM);  // row 1 sees only D[0][j-1] = init/H (pair_hmm.rs:515-529)
            } else {
                const double tMIp = A.qerr_d[A.read_ins[ro + i - 1]];
                const double tMDp = A.qerr_d[A.read_del[ro + i - 1]];
                const double a_i = a_tilde(i), a_p = a_tilde(i - 1);
                b = (float)(pm * tGM * tMIp / a_p);
                c = (float)(pm * tGM * tMDp / a_i);
                e = (float)(a_i * tII * tMIp / (A.qerr_d[qi] * a_p));
            }
            sm = (float)s_i;
            sx = (pm > 0.0) ? (float)(s_i * px / pm) : 0.f;
            f = (float)tII;  // tDD == tII
        } else if (i == R) {
            // first padding row under the read: I' becomes the column's summand M[R][j] + I[R][j]
            e = (float)(A.qerr_d[A.read_ins[ro + R - 1]] / a_tilde(R - 1));
        }
        cb[k] = b; cc[k] = c; ce[k] = e;
        if (VARGCP) cf[k] = f;
        // prior table: multiplier of row k against haplotype base A, C, G, T (one-hot bit of the read base's mask; a read
        // 'N' has all bits set) and against a haplotype 'N' (matches every real base), pair_hmm.rs:638-654
#pragma unroll
        for (int sl = 0; sl < N_PRIOR_SLICES; sl++) {
            const bool match = sl < 4 ? ((mask >> sl) & 1u) != 0u : mask != 0u;
            tab[(((sl * KQ) + (k >> 2)) * 32 + lane) * 4 + (k & 3)] = match ? sm : sx;
        }
    }
#pragma unroll
    for (int k = K; k < KQ * 4; k++)
#pragma unroll
        for (int sl = 0; sl < N_PRIOR_SLICES; sl++) tab[(((sl * KQ) + (k >> 2)) * 32 + lane) * 4 + (k & 3)] = 0.f;
    w.sync();
    if (lig == glanes - 1) {
        const int kstar = (R - 1) - lig * K;
        if (kstar == K - 1) { wM = 1.f; wI = (float)(A.qerr_d[A.read_ins[ro + R - 1]] / a_tilde(R - 1)); }
        else { wM = 0.f; wI = 1.f; }
    }

    const int32_t region = A.read_region[read];
    const int64_t h0 = A.region_hap_start[region], h1 = A.region_hap_start[region + 1];
    const int64_t out_base = A.read_out_base[read];

    int64_t h_first = h0, h_end = h1;
    if constexpr (SPLIT) {
        const TaskHaps th = A.task_haps[task_id];
        h_first = h0 + th.begin;
        h_end = h_first + th.count;
    }
#if LK_FAST_TMA && defined(__CUDA_ARCH__)
    // ---- bulk-copy staging of the encoded haplotypes (experiment) ----
    uint8_t *tma_buf = reinterpret_cast<uint8_t *>(tab + N_PRIOR_SLICES * KQ * 128);
    const uint32_t tma_buf_s = (uint32_t)__cvta_generic_to_shared(tma_buf);
    const uint32_t tma_bar_s = tma_buf_s + 2 * FAST_TMA_BUF;
    auto tma_issue = [&](int64_t hh, int buf) {  // lane 0 only
        const int HH = (int)(A.hap_off[hh + 1] - A.hap_off[hh]);
        const uint32_t bytes = (uint32_t)((HENC_PADL + HH + HENC_PADR + 15) & ~15);
        const uint8_t *src = A.hidx + A.henc_off[hh] - HENC_PADL;
        const uint32_t bar = tma_bar_s + 8u * (uint32_t)buf;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tma_buf_s + (uint32_t)buf * FAST_TMA_BUF),
                     "l"(src), "r"(bytes), "r"(bar)
                     : "memory");
    };
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tma_bar_s) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tma_bar_s + 8u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (h_first < h_end) tma_issue(h_first, 0);
    }
    __syncwarp();
    int tma_it = 0;
#endif
    for (int64_t h = h_first; h < h_end; ++h) {
        const int H = (int)(A.hap_off[h + 1] - A.hap_off[h]);
#if LK_FAST_TMA && defined(__CUDA_ARCH__)
        const int tma_cur = tma_it & 1;
        {
            const uint32_t bar = tma_bar_s + 8u * (uint32_t)tma_cur, parity = (uint32_t)(tma_it >> 1) & 1u;
            uint32_t ok;
            do {
                asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
            } while (!ok);
        }
        __syncwarp();  // every lane is past its last read of the other buffer (haplotype h-1)
        if (lane == 0 && h + 1 < h_end) tma_issue(h + 1, tma_cur ^ 1);
        tma_it++;
        const uint8_t *hp = tma_buf + tma_cur * FAST_TMA_BUF + HENC_PADL - 1 - lig;  // hp[s] = table slice of column (s - lig), in shared memory
#else
        const uint8_t *hp = A.hidx + A.henc_off[h] - 1 - lig;  // hp[s] = table slice of column (s - lig)
#endif
        const float kinit = kcoef * (FAST_INIT / (float)H);
        float M[K], I[K], D[K];
#pragma unroll
        for (int k = 0; k < K; k++) { M[k] = 0.f; I[k] = 0.f; D[k] = 0.f; }
        float Mu = 0.f, Iu = 0.f, Du = 0.f, Md = 0.f, Id = 0.f, Dd = 0.f, acc = 0.f;
        const int n_steps = H + maxlanes - 1;
        const int s_last = H + lig;  // last step at which this lane's column is still <= H

LK_UNROLL_N(LK_FAST_UNROLL)
        for (int s = 1; s <= n_steps; ++s) {
#if LK_FAST_TMA && defined(__CUDA_ARCH__)
            const float *trow = tab + ((int)hp[s] * KQ * 32 + lane) * 4;
#else
            const float *trow = tab + ((int)w.ldg(hp + s) * KQ * 32 + lane) * 4;
#endif
            float sel[KQ * 4];
#pragma unroll
            for (int q = 0; q < KQ; q++) w.lds4(sel + 4 * q, trow + q * 128);
            // ---- one column of the strip, rows top to bottom.  Row k reads row k-1 of the previous column (pm, pi, pd)
            // and, for I', row k-1 of this column (im, ii).  Written in this order the serial I' chain (one FFMA per row)
            // is spread evenly between the independent M'/D' work of the rows instead of being bunched at the end. ----
            {
                float pm = Md + kinit, pi = Id, pd = Dd, im = Mu, ii = Iu;  // start lanes: Md = 0, b'0 = c'0 = 0 -> mm = kinit
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const float om = M[k], oi = I[k], od = D[k];
                    // (statement order matters to ptxas here: this one measured 4 452 GCUPS, the textbook order 4 408,
                    // two others 4 045)
                    D[k] = fmaf(od, VARGCP ? cf[k] : (GM == GM_Q10 ? 0.1f : A.cf_uniform), om);
                    float mm = fmaf(pd, cc[k], pm);
                    mm = fmaf(pi, cb[k], mm);
                    M[k] = mm * sel[k];
                    I[k] = fmaf(ii, ce[k], im);
                    pm = om; pi = oi; pd = od; im = M[k]; ii = I[k];
                }
            }
            if (s <= s_last) {
                acc = fmaf(M[K - 1], wM, acc);
                acc = fmaf(I[K - 1], wI, acc);
            }
            Md = Mu; Id = Iu; Dd = Du;
            Mu = w.shfl_up1(M[K - 1]);
            Iu = w.shfl_up1(I[K - 1]);
            Du = w.shfl_up1(D[K - 1]);
            if (is_start) Mu = 0.f;  // b'0 = c'0 = e'0 = 0 already mute the other inputs of a group's first lane
        }

        if (lig == glanes - 1) {
            const int64_t oi = out_base + (h - h0);
            const bool redo = !(acc >= FAST_MIN_ACCEPTED) || rflag != 0 || A.hap_flags[h] != 0;
            if (redo) {
                const int slot = w.atomic_add(A.rescue_count, 1);
                if (slot < A.rescue_capacity) { A.rescue_pairs[slot].read = read; A.rescue_pairs[slot].hap = (int32_t)h; }
            } else {
                A.out[oi] = log10((double)acc) - FAST_INIT_LOG10;
                if (A.flags) A.flags[oi] = 0;
            }
        }
    }
#if LK_FAST_TMA && defined(__CUDA_ARCH__)
    __syncwarp();
    if (lane == 0) {  // the warp's next task initialises the barriers afresh
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tma_bar_s) : "memory");
        asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tma_bar_s + 8u) : "memory");
    }
#endif
}

// =============================================================================================
// Band path: fp32 scaled hat form for reads longer than the fast path's 32*K rows (cfg#3: 10 kbp reads).
// One warp per (read, haplotype) pair; rows go in bands of 32*K, the band's last row travels through
// per-warp scratch as 16-byte (M', I', D', exponent) elements.
//
// fp32 cannot hold the decay of hundreds of noisy rows (a 10 kbp ONT-like read loses ~7000 binary orders),
// so every lane carries a private binary exponent xi ("stored = true * 2^xi", block floating point per
// lane): every 16 columns a lane looks at the magnitude of its strip and, if it drifted out of
// its range (BAND_RANGE_LO/HI below), multiplies its whole state by a power of two and adjusts xi.  The exponent travels with the
// data (a 4th shuffle / the 4th word of the boundary element) and the receiver re-aligns what it gets
// with one FMUL per value.  Fully adaptive: no a-priori decay estimate, no retries.
// =============================================================================================
// 16-byte aligned so a boundary element moves as ONE 128-bit load / store (per-warp scratch slices are 16-byte multiples)
#if defined(__CUDACC__)
struct __align__(16) Bnd4 { float m, i, d; int32_t xi; };
#else
struct alignas(16) Bnd4 { float m, i, d; int32_t xi; };
#endif

constexpr int BAND_INIT_EXP = 40;  // initial condition 2^40 / H
// A lane re-centres its strip to 2^LK_BAND_CENTRE when its largest value leaves [2^-LK_BAND_SPAN, 2^+LK_BAND_SPAN] around
// the centre.  Head-room budget above the centre: what a lane receives can grow between two checkpoints (16 columns) by
// the whole spread of the sender's strip -- the alignment path enters the sender's strip at its top row and leaves at the
// bottom row K columns later, and the bottom row sits 2^-(15 + 3.3 (K - 2)) below the path (one insertion opening, K - 2
// extensions at gcp 10) -- on top of the sender's own allowed drift: centre + span + 15 + 3.3 (K - 2) must stay under
// 127.  With centre 20 / span 40 (the first version) that holds for K <= 16 and K >= 19 overflowed on every pair; centre 0
// / span 30 leaves room up to K = 24.
#ifndef LK_BAND_CENTRE
#define LK_BAND_CENTRE 0
#endif
#ifndef LK_BAND_SPAN
#define LK_BAND_SPAN 30
#endif
constexpr float band_pow2(int e) { float r = 1.f; for (int i = 0; i < (e < 0 ? -e : e); i++) r = e < 0 ? r * 0.5f : r * 2.f; return r; }
constexpr float BAND_RANGE_LO = band_pow2(LK_BAND_CENTRE - LK_BAND_SPAN);
constexpr float BAND_RANGE_HI = band_pow2(LK_BAND_CENTRE + LK_BAND_SPAN);

LK_DEV int __float_as_int_hd(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int i;
    memcpy(&i, &f, 4);
    return i;
#endif
}

LK_DEV float pow2i(int e) {  // 2^e for e in [-126, 127]
    e = e < -126 ? -126 : (e > 127 ? 127 : e);
#if defined(__CUDA_ARCH__)
    return __int_as_float((e + 127) << 23);
#else
    return ldexpf(1.0f, e);
#endif
}

// Floats of per-warp shared memory pairhmm_band_pair<K> needs: the prior table plus a 32-element ring of boundary
// elements (the previous band's last row, staged 32 columns at a time).
constexpr int BAND_RING = 32;
struct LeanTag { static constexpr bool value = true; };
struct GenericTag { static constexpr bool value = false; };
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr int band_smem_floats(int K) { return fast_table_floats(K) + BAND_RING * 4; }

// Structure of one band (R rows of the read against all H columns):
//   * steps 1..32 ("start-up": lanes come alive one by one; upkeep at every step) and the last steps, from the first
//     column block in which lane 0 runs past the haplotype, go through the GENERIC step (column-range gates);
//   * everything in between goes in BLOCKS of 16 steps through the LEAN step, in which every lane's column is inside
//     [1, H] by construction, so nothing is gated, and all block-floating-point upkeep happens once per block:
//     the checkpoint (re-centre the strip), ONE exchange of exponents (the factor `al` that brings the row above
//     into my scale stays in a register for the 16 receipts of the block), and every second block the staging of the
//     next 32 boundary elements of the previous band into the warp's shared-memory ring (one coalesced 16-byte load per
//     lane; lane 0 then receives them like values from a lane above it, see `stage`).
//   A lean step is: 1 LDG.U8 (next column's table slice) + KQ LDS.128 + K x (FMUL + 4 FFMA) + 2 FFMA (running sum)
//   + 3 SHFL + 3 FMUL (re-alignment) + predicated LDS.128 (lane 0's feed) / STG.128 (last lane's boundary row).
// Invariant: at the start of a step everything a lane holds (M, I, D, Mu.., Md.., acc) is in its own scale xi,
// and `al` = 2^(xi - xi_up) with xi_up the current exponent of the lane above (lane 0: of the staged ring).
template <class W, int K, bool Q10 = false>
LK_DEV void pairhmm_band_pair(W &w, const PairHmmArgs &A, int read, int64_t h, Bnd4 *scratch, float *tab, Bnd4 *ring) {
    constexpr int KQ = (K + 3) / 4;
    const int lane = w.lane();
    const int64_t ro = A.read_off[read];
    const int R = (int)(A.read_off[read + 1] - ro);
    const int H = (int)(A.hap_off[h + 1] - A.hap_off[h]);
    const int32_t region = A.read_region[read];
    const int64_t oi = A.read_out_base[read] + (h - A.region_hap_start[region]);
    auto to_rescue = [&]() {
        if (lane == 0) {
            const int slot = w.atomic_add(A.rescue_count, 1);
            if (slot < A.rescue_capacity) { A.rescue_pairs[slot].read = read; A.rescue_pairs[slot].hap = (int32_t)h; }
        }
    };
    if (A.read_flags[read] != 0 || A.hap_flags[h] != 0 || A.read_gcp != nullptr) { to_rescue(); return; }

    const int BAND = 32 * K;
    const int n_bands = (R + BAND - 1) / BAND;
    const int64_t W1 = (int64_t)H + 1;
    Bnd4 *bnd[2] = { scratch, scratch + W1 };
    const uint8_t *hp = A.hidx + A.henc_off[h] - 1 - lane;  // hp[s] = prior-table slice of column (s - lane)
    auto a_tilde = [&](int i) -> double {
        if (i < 1 || i >= R) return 1.0;
        return (1.0 - A.qerr_d[A.read_qual[ro + i]]) * m2m_lookup(A.m2m, A.read_ins[ro + i], A.read_del[ro + i]);
    };
    const float initH = ldexpf(1.0f, BAND_INIT_EXP) / (float)H;
    float acc_final = 0.f;
    int xi_final = 0;
    const bool feeds = (lane == 0);

    for (int band = 0; band < n_bands; band++) {
        const int row0 = band * BAND;
        const int rows = (R - row0 < BAND) ? (R - row0) : BAND;
        const int lanes = (rows + K - 1) / K;
        const bool last_band = (band == n_bands - 1);
        const Bnd4 *bin = bnd[band & 1];
        Bnd4 *bout = bnd[(band + 1) & 1];

        // ---- row constants (same scaled hat form as the fast kernel) ----
        // the match / mismatch multipliers go to the warp's shared-memory prior table, as in the fast kernel
        w.sync();
        float cb[K], cc[K], ce[K];
        float kcoef = 0.f, wM = 0.f, wI = 0.f;
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int i = row0 + lane * K + k;  // 0-based read position
            float b = 0.f, c = 0.f, e = (i > R) ? 1.f : 0.f, sm = 0.f, sx = 0.f;
            uint32_t mask = 0u;
            if (i < R) {
                const int q = A.read_qual[ro + i], qi = A.read_ins[ro + i];
                const double err = A.qerr_d[q];
                const double pm = 1.0 - err, px = err / (double)A.mismatch_div;
                const double tGM = 1.0 - A.qerr_d[A.gcp_constant], tII = A.qerr_d[A.gcp_constant];
                const double s_i = a_tilde(i + 1);
                bool ex = false;
                mask = encode_base(A.read_bases[ro + i], ex) & 0xFFu;
                if (i == 0) {
                    kcoef = (float)(pm * tGM);
                } else {
                    const double tMIp = A.qerr_d[A.read_ins[ro + i - 1]], tMDp = A.qerr_d[A.read_del[ro + i - 1]];
                    const double a_i = a_tilde(i), a_p = a_tilde(i - 1);
                    b = (float)(pm * tGM * tMIp / a_p);
                    c = (float)(pm * tGM * tMDp / a_i);
                    e = (float)(a_i * tII * tMIp / (A.qerr_d[qi] * a_p));
                }
                sm = (float)s_i;
                sx = (pm > 0.0) ? (float)(s_i * px / pm) : 0.f;
            } else if (i == R) {
                e = (float)(A.qerr_d[A.read_ins[ro + R - 1]] / a_tilde(R - 1));
            }
            // read row 1 sees only D[0][j-1] = init / H of the virtual row 0 (pair_hmm.rs:515-529): that constant is what
            // the ring feeds lane 0 in band 0 (d = init / H), and c'(row 1) = pm * tGM picks it up
            if (i == 0) c = kcoef;
            cb[k] = b; cc[k] = c; ce[k] = e;
#pragma unroll
            for (int sl = 0; sl < N_PRIOR_SLICES; sl++) {
                const bool match = sl < 4 ? ((mask >> sl) & 1u) != 0u : mask != 0u;
                tab[(((sl * KQ) + (k >> 2)) * 32 + lane) * 4 + (k & 3)] = match ? sm : sx;
            }
        }
#pragma unroll
        for (int k = K; k < KQ * 4; k++)
#pragma unroll
            for (int sl = 0; sl < N_PRIOR_SLICES; sl++) tab[(((sl * KQ) + (k >> 2)) * 32 + lane) * 4 + (k & 3)] = 0.f;
        if (band == 0) { Bnd4 z; z.m = 0.f; z.i = 0.f; z.d = initH; z.xi = 0; ring[lane] = z; }  // band 0: the virtual row 0
        w.sync();
        if (last_band && lane == lanes - 1) {
            const int kstar = (rows - 1) - lane * K;
            if (kstar == K - 1) { wM = 1.f; wI = (float)(A.qerr_d[A.read_ins[ro + R - 1]] / a_tilde(R - 1)); }
            else { wM = 0.f; wI = 1.f; }
        }

        float M[K], I[K], D[K];
#pragma unroll
        for (int k = 0; k < K; k++) { M[k] = 0.f; I[k] = 0.f; D[k] = 0.f; }
        float Mu = 0.f, Iu = 0.f, Du = 0.f, Md = 0.f, Id = 0.f, Dd = 0.f, acc = 0.f;
        if (band == 0 && lane == 0) { Dd = initH; Du = initH; }  // D[0][0], D[0][1]
        float al = 1.f;
        int xi = 0, xi_up = 0;  // my exponent; exponent of the lane above (lane 0: of the ring, xc)
        const int n_steps = H + lanes - 1;
        const int s_last = H + lane;
        const bool writes = !last_band && (lane == lanes - 1);
        Bnd4 *bw = bout - lane;  // bw[s] = boundary element of my column at step s
        if (band > 0 && lane == 0) {  // column 1 of the row above: adopted as it is (al = 1)
            const Bnd4 v = bin[1];
            Mu = v.m; Iu = v.i; Du = v.d; xi = v.xi;
        }

        // ring <- boundary elements of columns c0 + 2 .. c0 + 33 (what lane 0 receives at the end of steps c0 + 1 .. c0 + 32):
        // a plain coalesced copy.  No re-scaling is needed: the lane that wrote them (lane 31 of the previous band, at step
        // column + 31) changes its exponent only at ITS upkeep steps (s = 1 mod 16 once past the start-up), i.e. between
        // columns 16j + 1 and 16j + 2 -- exactly where lane 0 of this band, which receives column s + 1 at the end of step
        // s, has its own upkeep.  The 16 elements one block receives therefore share one exponent, like 16 receipts
        // from a lane above.
        auto stage = [&](int c0) {
            const int c = c0 + 2 + lane;
            Bnd4 e = bin[c <= H ? c : H];
            if (c > H) { e.m = 0.f; e.i = 0.f; e.d = 0.f; }
            w.prefetch(bin + (c + 32 <= H ? c + 32 : H));
            w.sync();  // every lane is past its last read of the previous chunk
            ring[lane] = e;
            w.sync();
        };
        // checkpoint + exchange at the START of a step (every step while lanes are starting, s <= 32; then when s = 1 mod 16):
        // re-centre what I hold into [BAND_RANGE_LO, BAND_RANGE_HI] (looking at the running sum and at what has just arrived too), then learn
        // the exponent of the lane above and set `al`.
        auto upkeep = [&](int s) {
            if (band > 0 && ((s - 1) & 31) == 0) stage(s - 1);
            float m = fmaxf(acc, fmaxf(Mu, fmaxf(Iu, Du)));
#pragma unroll
            for (int k = 0; k < K; k++) m = fmaxf(m, fmaxf(M[k], I[k]));
            if (m > 0.f && (m < BAND_RANGE_LO || m > BAND_RANGE_HI)) {
                int sh = LK_BAND_CENTRE - ilogbf(m);
                sh = sh > 120 ? 120 : (sh < -120 ? -120 : sh);
                const float f = pow2i(sh);
#pragma unroll
                for (int k = 0; k < K; k++) { M[k] *= f; I[k] *= f; D[k] *= f; }
                Mu *= f; Iu *= f; Du *= f; Md *= f; Id *= f; Dd *= f; acc *= f;
                xi += sh;
            }
            xi_up = w.shfl_up1(xi);
            if (lane == 0) xi_up = ring[(s - 1) & 31].xi;  // exponent of what this block receives (band 0: 0)
            if (s <= lane) xi = xi_up;  // not started yet (holds zeros only): adopt the sender's scale
            al = pow2i(xi - xi_up);
        };
        // table slice of my column, fetched two steps ahead of the LDS address it forms
        unsigned sl_a = w.ldg(hp + 1), sl_b = w.ldg(hp + 2);
        const float *tab_lane = tab + lane * 4;
        // One step.  hq = hp + s, bq = bw + s, rp = ring + ((s - 1) & 31) (passed in so that the unrolled lean block addresses
        // them with immediate offsets).  LEAN = true: every lane's column is inside [1, H] (no gates); false: generic step.
        auto step = [&](int s, const uint8_t *hq, Bnd4 *bq, const Bnd4 *rp, auto lean_tag) {
            constexpr bool LEAN = decltype(lean_tag)::value;
            const float *trow = tab_lane + sl_a * (unsigned)(KQ * 128);
            sl_a = sl_b;
            sl_b = w.ldg(hq + 2);
            float sel[KQ * 4];
#pragma unroll
            for (int q = 0; q < KQ; q++) w.lds4(sel + 4 * q, trow + q * 128);
            // ---- one column of the strip, rows top to bottom.  Row k reads row k-1 of the previous column (pm, pi, pd)
            // and, for I', row k-1 of this column (im, ii).  Written in this order the serial I' chain (one FFMA per row)
            // is spread evenly between the independent M'/D' work of the rows instead of being bunched at the end. ----
            {
                float pm = Md, pi = Id, pd = Dd, im = Mu, ii = Iu;  // start lanes: b'0 = 0 and Md = 0 -> mm = c'0 * D[0][j-1]
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const float om = M[k], oi = I[k], od = D[k];
#if LK_BAND_ORDER == 1
                    D[k] = fmaf(od, Q10 ? 0.1f : A.cf_uniform, om);
                    float mm = fmaf(pd, cc[k], pm);
                    mm = fmaf(pi, cb[k], mm);
                    M[k] = mm * sel[k];
                    I[k] = fmaf(ii, ce[k], im);
#else
                    float mm = fmaf(pi, cb[k], pm);
                    mm = fmaf(pd, cc[k], mm);
                    D[k] = fmaf(od, Q10 ? 0.1f : A.cf_uniform, om);
                    M[k] = mm * sel[k];
                    I[k] = fmaf(ii, ce[k], im);
#endif
                    pm = om; pi = oi; pd = od; im = M[k]; ii = I[k];
                }
            }
            if (LEAN) {
                acc = fmaf(M[K - 1], wM, acc);
                acc = fmaf(I[K - 1], wI, acc);
                if (writes) { Bnd4 v; v.m = M[K - 1]; v.i = I[K - 1]; v.d = D[K - 1]; v.xi = xi; *bq = v; }
            } else {
                // columns past the haplotype read the guard band of the encoded haplotype (an ordinary table slice), so what
                // a lane holds there is finite: the sum is gated by zeroed weights and the boundary store by a predicate
                const bool in_cols = s <= s_last;
                acc = fmaf(M[K - 1], in_cols ? wM : 0.f, acc);
                acc = fmaf(I[K - 1], in_cols ? wI : 0.f, acc);
                if (writes && in_cols && s > lane) { Bnd4 v; v.m = M[K - 1]; v.i = I[K - 1]; v.d = D[K - 1]; v.xi = xi; *bq = v; }
            }
            Md = Mu; Id = Iu; Dd = Du;
            float rm = w.shfl_up1(M[K - 1]), ri = w.shfl_up1(I[K - 1]), rd = w.shfl_up1(D[K - 1]);
            if (feeds) { const Bnd4 v = *rp; rm = v.m; ri = v.i; rd = v.d; }  // row above lane 0: previous band's last row
            Mu = rm * al; Iu = ri * al; Du = rd * al;
        };

        constexpr int BLK = LK_BAND_BLOCK;  // steps per lean block = upkeep cadence (16 or 32; must divide the 32-element ring)
        static_assert(BLK == 16 || BLK == 32, "the ring holds 32 elements and the exponent groups follow the upkeep cadence");
        const int s_lean_end = (H / BLK) * BLK;  // last step of the last block in which lane 0's column is still <= H
        int s = 1, gen_end = n_steps < 32 ? n_steps : 32;
        for (;;) {  // generic start-up, lean blocks, generic tail (one copy of the generic loop in the code)
            for (; s <= gen_end; ++s) {
                if (s <= 32 || (s & (BLK - 1)) == 1) upkeep(s);
                step(s, hp + s, bw + s, ring + ((s - 1) & 31), GenericTag());
            }
            if (gen_end == n_steps) break;
            while (s + BLK - 1 <= s_lean_end) {  // s = 1 (mod BLK)
                upkeep(s);
                const Bnd4 *rp = ring + ((s - 1) & 31);
                const uint8_t *hq = hp + s;
                Bnd4 *bq = bw + s;
#pragma unroll 1
                for (int u4 = 0; u4 < BLK / LK_BAND_UNROLL; ++u4) {
#pragma unroll
                    for (int u = 0; u < LK_BAND_UNROLL; ++u) step(s + u, hq + u, bq + u, rp + u, LeanTag());
                    hq += LK_BAND_UNROLL; bq += LK_BAND_UNROLL; rp += LK_BAND_UNROLL; s += LK_BAND_UNROLL;
                }
            }
            gen_end = n_steps;
        }
        w.sync();  // the band's boundary row is visible to lane 0 of the next band
        if (last_band) { acc_final = w.shfl(acc, lanes - 1); xi_final = w.shfl(xi, lanes - 1); }
    }
#if !defined(__CUDACC__) && defined(LK_EMUL_DEBUG)
    if (lane == 0) fprintf(stderr, "band result acc %g xi %d\n", (double)acc_final, xi_final);
#endif
#if defined(__CUDA_ARCH__) && defined(LK_BAND_DEBUG)
    if (lane == 0 && (!(acc_final > 0.f) || !(acc_final < 3.0e38f))) printf("band rescue read %d hap %d R %d H %d acc %g xi %d\n", read, (int)h, R, H, (double)acc_final, xi_final);
#endif
    if (!(acc_final > 0.f) || !(acc_final < 3.0e38f)) { to_rescue(); return; }  // flushed to zero / overflowed: f64 pass
    if (lane == 0) {
        A.out[oi] = log10((double)acc_final) - (double)(xi_final + BAND_INIT_EXP) * 0.30102999566398120;
        if (A.flags) A.flags[oi] = 0;
    }
}

// =============================================================================================
// f64 pass ("plain form"): one warp per (read, haplotype) pair, any lengths.  Rows are processed
// in bands of 32*K; the band's last row travels through per-warp scratch; between bands the row is
// renormalised by a power of two (exponent carried in an int) so long reads cannot underflow.
// The arithmetic order is the reference's (pair_hmm.rs:575-591) with explicit un-fused mul/add.
// =============================================================================================
LK_DEV double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
LK_DEV double dadd(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}

constexpr double PLAIN_INIT_EXP = 1020.0;  // INITIAL_CONDITION = 2^1020, pair_hmm.rs:16

template <class W, int K>
LK_DEV void pairhmm_plain_pair(W &w, const PairHmmArgs &A, int read, int64_t h, double *scratch) {
    const int lane = w.lane();
    const int64_t ro = A.read_off[read];
    const int R = (int)(A.read_off[read + 1] - ro);
    const int64_t ho = A.hap_off[h];
    const int H = (int)(A.hap_off[h + 1] - ho);
    const int32_t region = A.read_region[read];
    const int64_t oi = A.read_out_base[read] + (h - A.region_hap_start[region]);

    const int BAND = 32 * K;
    const int n_bands = (R + BAND - 1) / BAND;
    const double initH = ldexp(1.0, (int)PLAIN_INIT_EXP) / (double)H;  // pair_hmm.rs:517

    // scratch: two boundary rows, each 3 arrays of (H+1)
    const int64_t W1 = (int64_t)H + 1;
    double *bnd[2] = { scratch, scratch + 3 * W1 };

    long shift_total = 0;     // power of two removed so far
    double in_scale = 1.0;    // 2^shift applied to the incoming boundary row of this band
    double result_sum = 0.0;
    bool underflow = false;

    for (int band = 0; band < n_bands; band++) {
        const int row0 = band * BAND;                        // first 0-based read position of the band
        const int rows = (R - row0 < BAND) ? (R - row0) : BAND;
        const int lanes = (rows + K - 1) / K;                // lanes holding real rows
        const bool last_band = (band == n_bands - 1);
        const double *bin = bnd[band & 1];
        double *bout = bnd[(band + 1) & 1];

        // ---- row constants ----
        double tMM[K], tGM[K], tMI[K], tII[K], tMD[K], tDD[K], pm[K], px[K];
        uint32_t rb[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int i = row0 + lane * K + k;
            tMM[k] = tGM[k] = tMD[k] = tDD[k] = pm[k] = px[k] = 0.0;
            tMI[k] = tII[k] = 1.0;  // padding rows: I copies M+I of the row above (column summand)
            rb[k] = 0x100u;         // never equals a byte
            if (i < R) {
                const int q = A.read_qual[ro + i], qi = A.read_ins[ro + i], qd = A.read_del[ro + i];
                const int qg = A.read_gcp ? A.read_gcp[ro + i] : A.gcp_constant;
                tMM[k] = m2m_lookup(A.m2m, qi, qd);
                tGM[k] = 1.0 - A.qerr_d[qg];
                tMI[k] = A.qerr_d[qi];
                tII[k] = A.qerr_d[qg];
                tMD[k] = A.qerr_d[qd];
                tDD[k] = A.qerr_d[qg];
                pm[k] = 1.0 - A.qerr_d[q];
                px[k] = A.qerr_d[q] / (double)A.mismatch_div;
                rb[k] = A.read_bases[ro + i];
            }
        }
        // which lane/row carries the band's last real row
        const int last_lane = lanes - 1;
        const int kstar = (rows - 1) - last_lane * K;
        const double wM = (lane == last_lane && kstar == K - 1) ? 1.0 : 0.0;
        const bool is_out_lane = (lane == last_lane);

        double M[K], I[K], D[K];
#pragma unroll
        for (int k = 0; k < K; k++) { M[k] = 0.0; I[k] = 0.0; D[k] = 0.0; }
        // column 0: row 0 has D = init (pair_hmm.rs:520), rows >= 1 are zero
        double Mu = 0.0, Iu = 0.0, Du = 0.0, Md = 0.0, Id = 0.0, Dd = 0.0;
        if (lane == 0 && band == 0) Dd = initH;
        double band_max = 0.0;
        const int n_steps = H + lanes - 1;
        if (lane == 0) { bout[0] = 0.0; bout[W1] = 0.0; bout[2 * W1] = 0.0; }

        for (int s = 1; s <= n_steps; s++) {
            const int c = s - lane;  // this lane's column
            if (lane == 0) {
                if (band == 0) { Mu = 0.0; Iu = 0.0; Du = initH; }
                else if (c <= H) { Mu = bin[c] * in_scale; Iu = bin[W1 + c] * in_scale; Du = bin[2 * W1 + c] * in_scale; }
                else { Mu = 0.0; Iu = 0.0; Du = 0.0; }
            }
            const uint32_t hb = (c >= 1 && c <= H) ? (uint32_t)A.hap_bases[ho + c - 1] : 0x200u;
            const bool hN = (hb == 'N');
#pragma unroll
            for (int k = K - 1; k >= 0; k--) {
                const double md = (k == 0) ? Md : M[k - 1];
                const double id = (k == 0) ? Id : I[k - 1];
                const double dd = (k == 0) ? Dd : D[k - 1];
                const bool match = (rb[k] == hb) || (rb[k] == 'N') || hN;  // pair_hmm.rs:643
                const double prior = match ? pm[k] : px[k];
                const double sum = dadd(dadd(dmul(md, tMM[k]), dmul(id, tGM[k])), dmul(dd, tGM[k]));
                const double dn = dadd(dmul(M[k], tMD[k]), dmul(D[k], tDD[k]));
                M[k] = dmul(prior, sum);
                D[k] = dn;
            }
            {
                double mu = Mu, iu = Iu;
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const double in = dadd(dmul(mu, tMI[k]), dmul(iu, tII[k]));
                    mu = M[k];
                    I[k] = in;
                    iu = I[k];
                }
            }
            if (is_out_lane && c >= 1 && c <= H) {
                if (last_band) {
                    result_sum = dadd(result_sum, dadd(dmul(M[K - 1], wM), I[K - 1]));  // pair_hmm.rs:600-603
                } else {
                    bout[c] = M[K - 1]; bout[W1 + c] = I[K - 1]; bout[2 * W1 + c] = D[K - 1];
                    band_max = fmax(band_max, fmax(M[K - 1], fmax(I[K - 1], D[K - 1])));
                }
            }
            Md = Mu; Id = Iu; Dd = Du;
            Mu = w.shfl_up1(M[K - 1]);
            Iu = w.shfl_up1(I[K - 1]);
            Du = w.shfl_up1(D[K - 1]);
        }
        if (!last_band) {
            band_max = w.shfl(band_max, last_lane);
            w.sync();  // boundary row visible to lane 0 of the next band
            if (!(band_max > 0.0)) { underflow = true; break; }
            int e2;
            (void)frexp(band_max, &e2);
            const int k2 = 1000 - e2;  // bring the row maximum back near 2^1000
            in_scale = ldexp(1.0, k2);
            shift_total += k2;
        }
    }

    if (lane == ((R - (n_bands - 1) * BAND) + K - 1) / K - 1) {
        double v;
        uint8_t fl = PAIR_RESCUED;
        if (underflow || !(result_sum > 0.0)) { v = -INFINITY; fl |= PAIR_NEG_INF; }
        else v = log10(result_sum) - log10(ldexp(1.0, (int)PLAIN_INIT_EXP)) - (double)shift_total * log10(2.0);
        A.out[oi] = v;
        if (A.flags) A.flags[oi] = fl;
    }
}

}  // namespace lk
