// sw16_core.cuh — TWO Smith-Waterman problems per warp, in the two 16-bit halves of every register ("duo").
//
// Same algorithm and results as sw_core.cuh (src/smith_waterman/smith_waterman_aligner.rs calculate_matrix :124-263,
// calculate_cigar :273-442), different arithmetic: the matrix is carried as DIFFERENCES between neighbouring cells, which
// are bounded by the weights (|.| <= 2 max|w| + 2 max|gap|), so a cell of two independent pairs fits the halves of one
// 32-bit word and sm_100a's packed DPX forms (VIMNMX3.U16x2, VIMNMX.U16x2) serve both at once.  With H the reference's
// `sw`, F / E its best vertical / horizontal gap score ending at a cell (step_down / step_right, :210-245):
//     u(i,j) = H(i,j) - H(i-1,j)        v(i,j) = H(i,j) - H(i,j-1)
//     nf(i,j) = H(i,j) - F(i,j) >= 0    ne(i,j) = H(i,j) - E(i,j) >= 0         D = extend - open >= 0
//     F(i+1,j) = H(i,j) + extend - min(nf(i,j), D),  opened iff nf(i,j) > D    (prev_gap > best_gap, strict, :212-218)
// and relative to the diagonal predecessor H(i-1,j-1) a cell is
//     a = v(i-1,j) + extend - nx(i-1,j)   b = u(i,j-1) + extend - ny(i,j-1)   z = max3(s, a, b)
//     u = z - v(i-1,j)   v = z - u(i,j-1)   nf = z - a   ne = z - b   nx = min(nf, D)   ny = min(ne, D)
// The choice bits are the same comparisons the reference makes on absolute scores (both sides of each comparison are
// shifted by the same H(i-1,j-1)): not diagonal iff z > s (:248-252), vertical iff b < a, i.e. ne >= 1 when not diagonal
// (:253-260).  Absolute scores are needed only along the last column and the bottom row (:301-329); they are rebuilt from
// u and v by prefix sums after the fill.
//
// Ranges (S = max|substitution weight|, G = max|gap weight|, both <= 1000 here).  u, v >= open >= -G because a cell is at
// least the gap opened from its upper / left neighbour.  u <= S + G: if the cell took the diagonal, u = s - v(i-1,j) <= S + G;
// if it took the vertical gap, u = extend - nx <= 0; if it took a horizontal gap of length k, the same gap is available to
// row i-1 (E(i-1,j) >= H(i-1,j-k) + open + (k-1) extend), so u(i,j) <= u(i,j-k), which ends at one of the first two cases or
// in column 0 (u in {0, open, extend}); v by symmetry.  Hence z = u + v(i-1,j) in [-2G, 2S + 2G], a, b in [-3G, z],
// nf = H - F <= u - open <= S + 2G, and with the biases below every half stays inside [1048, 8096] (values) or below
// 0x7fff + 5000 (flag words): no carry ever leaves a half.  The same holds for the cells nobody reads (rows past a reference,
// columns past an alternate): they are ordinary cells of a larger problem.
//
// Representation: every half is an UNSIGNED biased value and every sum below ends in [0, 65535] per half, so plain 32-bit
// adds and subtracts never carry between the halves and can issue on either integer pipe (IADD3 on the ALU pipe, IMAD on
// the FMA pipe): u, v carry bias B; s, a, b, z carry 2 B; nf, ne, nx, ny are plain.  A flag is the top bit of a half of
// (value + 0x7fff - threshold); one PRMT (sign-replicating byte select) gathers the four flags of two rows x two pairs as
// 0x00 / 0xff bytes and one IMAD files them (plane = 2 plane + bytes, i.e. 255 x the bit plane, divided out by the traceback
// with the inverse of 255 mod 2^32), so the four bit-planes cost 4 PRMT (ALU pipe) + 4 IMAD (FMA pipe) per two rows of two
// pairs.
// The gap flags are computed one cell EARLY (vo stored with (i,j) is the opening of the vertical gap ending at (i+1,j)),
// which the traceback accounts for (sw_finish<1>).
//
// Eligibility (sw16_weights_ok + the checks in sw16_align_duo; everything else takes sw_core.cuh): |weights| <= 1000,
// open <= extend < 0, the -1e8 clamp proven inactive, both alternates made of A C G T N only (the classes of the weight
// tables), both references within one band of the launch's K class (n <= 32 K; K = 8 / 12 / 16 / 17, references of 129..544 rows;
// the two lengths may differ).
#pragma once
#include "sw_core.cuh"

#ifndef LK_SW16_LEAN_UNROLL
#define LK_SW16_LEAN_UNROLL 1
#endif
// which of the cell's three-input sums issue as one IADD3 on the ALU pipe (the others as two IMADs on the FMA pipe):
// bit 0 a, bit 1 b, bit 2 the not-diagonal flag word
#ifndef LK_SW16_ALU_MASK
#define LK_SW16_ALU_MASK 7
#endif
// 1 (default): both pairs' weight tables hold 16-bit values (11.5 KB per warp, one PRMT per cell packs the two weights): the
// shared-memory footprint that lets 16 warps share an SM.  0: pair 0's table in 32-bit words (16.6 KB per warp, 12 warps per
// SM, the pack is an IMAD for even rows and a LOP3 for odd ones) -- measured 1 896 against 1 966 GCUPS on cfg#4.
#ifndef LK_SW16_COMPACT
#define LK_SW16_COMPACT 1
#endif
// latency-bound set-up phases: how many of a lane's reference rows / alternate columns are loaded per round trip (all the
// rows at once measured best: 2 011 GCUPS against 1 940 in batches of six)
#ifndef LK_SW16_ROW_BATCH
#define LK_SW16_ROW_BATCH 32
#endif
#ifndef LK_SW16_CLASS_BATCH
#define LK_SW16_CLASS_BATCH 4
#endif

// developer builds (-DLK_SW16_PROFILE): lane 0 adds the clock64 ticks of each phase to A.packed_count[1 + phase]
#if defined(LK_SW16_PROFILE) && defined(__CUDA_ARCH__)
#define LK_SW16_TICK(phase) do { const long long t_now_ = clock64(); if (w.lane() == 0) atomicAdd((unsigned long long *)A.prof + (phase), (unsigned long long)(t_now_ - t_prev_)); t_prev_ = t_now_; } while (0)
#define LK_SW16_TICK_INIT long long t_prev_ = clock64()
#else
#define LK_SW16_TICK(phase) do { } while (0)
#define LK_SW16_TICK_INIT do { } while (0)
#endif

namespace lk {

constexpr uint32_t SW16_B = 2048;   // bias of the u / v halves; s, a, b, z carry twice that
constexpr int32_t SW16_MAX_WEIGHT = 1000;
constexpr uint32_t SW16_INV255 = 0xfefefeffu;  // 255 * SW16_INV255 = 1 (mod 2^32): the stored plane words are 255 x the bit planes

LK_HD bool sw16_weights_ok(int32_t wm, int32_t wx, int32_t wo, int32_t we) {
    const int32_t L = SW16_MAX_WEIGHT;
    return wm >= -L && wm <= L && wx >= -L && wx <= L && wo >= -L && we >= -L && wo <= we && we < 0;
}

LK_DEV uint32_t sw16_both(uint32_t x) { return x | (x << 16); }  // the same 16-bit value in both halves
LK_DEV uint32_t sw16_max3(uint32_t a, uint32_t b, uint32_t c) {   // per-half unsigned max of three: one VIMNMX3.U16x2
#if defined(__CUDA_ARCH__)
    return __vimax3_u16x2(a, b, c);
#else
    uint32_t r = 0;
    for (int h = 0; h < 32; h += 16) {
        uint32_t x = (a >> h) & 0xffffu, y = (b >> h) & 0xffffu, z = (c >> h) & 0xffffu;
        x = x > y ? x : y;
        r |= (x > z ? x : z) << h;
    }
    return r;
#endif
}
LK_DEV uint32_t sw16_min(uint32_t a, uint32_t b) {                // per-half unsigned min: one VIMNMX.U16x2
#if defined(__CUDA_ARCH__)
    return __vminu2(a, b);
#else
    uint32_t r = 0;
    for (int h = 0; h < 32; h += 16) {
        const uint32_t x = (a >> h) & 0xffffu, y = (b >> h) & 0xffffu;
        r |= (x < y ? x : y) << h;
    }
    return r;
#endif
}
// bytes [a.lo, a.hi, b.lo, b.hi] -> 0xff where the half's top bit is set, else 0x00: one PRMT with sign replication
LK_DEV uint32_t sw16_flags(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, 0xfdb9;" : "=r"(d) : "r"(a), "r"(b));
    return d;
#else
    return ((a >> 15) & 1u ? 0xffu : 0u) | ((a >> 31) & 1u ? 0xff00u : 0u) | ((b >> 15) & 1u ? 0xff0000u : 0u) | ((b >> 31) & 1u ? 0xff000000u : 0u);
#endif
}
// the low (odd = 0) or high (odd = 1) halves of a and b as one word, a's in the low half: one PRMT
LK_DEV uint32_t sw16_halves(uint32_t a, uint32_t b, int odd) {
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, odd ? 0x7632 : 0x5410);
#else
    return odd ? ((a >> 16) | (b & 0xffff0000u)) : ((a & 0xffffu) | (b << 16));
#endif
}
// a + b and a - b issued as IMADs (FMA pipe) through multipliers the compiler cannot see through (SwArgs::one / mone):
// the ALU pipe is the bound of the cell, the FMA pipe has room
LK_DEV uint32_t sw16_fadd(uint32_t a, uint32_t b, uint32_t one) {
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b));
    return d;
#else
    return a * one + b;
#endif
}
LK_DEV uint32_t sw16_fsub(uint32_t a, uint32_t b, uint32_t mone) {
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(mone), "r"(a));
    return d;
#else
    return b * mone + a;
#endif
}

// per-warp substitution-weight tables of a duo: for each pair, each alternate class A C G T N and each of the lane's K rows
// the biased weight.  Pair 0's as 32-bit words (value in the low half): rows 0..15 of a lane as int4 quads
// [class][quad][lane], a 17th row as [class][lane] words.  Pair 1's as 16-bit values, rows (2r, 2r+1) sharing a word (even row
// low), int4 quads of eight rows, and a 17th row already in the HIGH half of its [class][lane] word -- the two 32-bit tables
// of a warp would not leave room for 12 warps per SM.  A cell's two weights are packed by one IMAD (even rows: w1 * 65536 + w0)
// or one LOP3 (odd rows: (w1 & 0xffff0000) | w0).
template <int K> struct Sw16Table {
    static_assert(K == 17 || (K % 4 == 0 && K <= 16), "rows come in groups of four, plus one odd row for K = 17");
    static constexpr int rows16 = K < 16 ? K : 16;
    static constexpr int quads_a = rows16 / 4, quads_b = (rows16 + 7) / 8;  // K = 12: the second 16-bit quad is half used
    static constexpr int a_quad_words = SW_N_CLASSES * (LK_SW16_COMPACT ? quads_b : quads_a) * 128, b_quad_words = SW_N_CLASSES * quads_b * 128;
    static constexpr int a_words = a_quad_words + (K > 16 ? SW_N_CLASSES * 32 : 0);
    static constexpr int b_words = b_quad_words + (K > 16 ? SW_N_CLASSES * 32 : 0);
    static constexpr int words = a_words + b_words;
};

// words of per-warp scratch a duo needs (M = the longer alternate): four plane words per (step, lane), one byte per
// (step, lane) for a 17th row, the bottom row's packed v, one rebuilt bottom row, one class byte per column
LK_HD int64_t sw16_scratch_words(int M, int K) {
    const int64_t steps = (int64_t)M + 32;
    const int64_t w = steps * 128 + (K > 16 ? steps * 8 : 0) + 2 * ((int64_t)M + 1) + ((int64_t)M + 3) / 4 + 8;
    return (w + 3) & ~(int64_t)3;
}

// Aligns pairs pa and pb (reference lengths <= 32 K, prologue already done for both).  false: not eligible (an
// alternate holds a byte outside A C G T N) -- nothing has been written for either pair and the caller takes sw_dp_pair.
// SAME_N: both references have the same length (one set of row indices; the case every haplotype of a region against that
// region's reference is).  The general form costs ~6 % on cfg#4, and both forms inlined into ONE kernel cost the faster one
// its register allocation, so each lives in a kernel of its own (sw_duo_kernel<K, SAME_N>) and the host gives each its pairs.
template <class Wp, int K, bool SAME_N>
LK_DEV bool sw16_align_duo(Wp &w, const SwArgs &A, int pa, int pb, uint32_t *scratch, int32_t *wtab_i) {
    typedef Sw16Table<K> T;
    constexpr int QA = T::quads_a, QB = T::quads_b;
    constexpr int R16 = T::rows16;
    constexpr bool HAS17 = K > 16;
    uint32_t *wtab = reinterpret_cast<uint32_t *>(wtab_i);
    const int lane = w.lane();
    const int pr[2] = { pa, pb };
    const uint8_t *ref[2] = { A.refs + A.ref_off[pa], A.refs + A.ref_off[pb] };
    const uint8_t *alt[2] = { A.alts + A.alt_off[pa], A.alts + A.alt_off[pb] };
    const int n0 = (int)(A.ref_off[pa + 1] - A.ref_off[pa]), n1 = SAME_N ? n0 : (int)(A.ref_off[pb + 1] - A.ref_off[pb]);
    const int N = n0 > n1 ? n0 : n1;
    const int mm[2] = { (int)(A.alt_off[pa + 1] - A.alt_off[pa]), (int)(A.alt_off[pb + 1] - A.alt_off[pb]) };
    const int M = mm[0] > mm[1] ? mm[0] : mm[1], m_min = mm[0] < mm[1] ? mm[0] : mm[1];
    const int strategy = A.strategy;
    const int32_t wm = A.w_match, wx = A.w_mismatch, wo = A.w_open, we = A.w_extend;
    const bool edge_gaps = (strategy == SW_INDEL || strategy == SW_LEADING_INDEL);
    const bool need_bottom = !edge_gaps;
    constexpr int BAND = 32 * K;
    // SoftClip / Ignore: each pair's rows are shifted down so that its matrix's last row is register K-1 of lane 31 (see
    // sw_core.cuh, bottom alignment); the pad_top rows above row 1 are virtual rows that reproduce the all-zero top row, as
    // many as that pair's reference is short of 32 K.  The edge-gap strategies cannot have virtual rows (their top row
    // carries gap penalties): both pairs start at row 1 and the shorter reference is followed by rows nobody reads.
    const int pad0 = need_bottom ? BAND - n0 : 0, pad1 = SAME_N ? pad0 : (need_bottom ? BAND - n1 : 0);
    const int lanes = need_bottom ? 32 : (N + K - 1) / K;
    LK_SW16_TICK_INIT;
    uint32_t *planes = scratch;                                               // [step][lane] x {pd, pr, vo, ho}
    uint8_t *nibs = reinterpret_cast<uint8_t *>(planes + ((int64_t)M + 32) * 128);  // [step][lane], row 16 of both pairs
    uint32_t *botV = reinterpret_cast<uint32_t *>(nibs) + (HAS17 ? ((int64_t)M + 32) * 8 : 0);  // packed v of the bottom row
    int32_t *botH = reinterpret_cast<int32_t *>(botV + (M + 1));              // one pair's bottom row, rebuilt
    uint8_t *acls = reinterpret_cast<uint8_t *>(botH + (M + 1));              // class of column j: pair 0 | pair 1 << 4

    {
        bool other = false;
        constexpr int CB = LK_SW16_CLASS_BATCH;
        for (int j0 = 0; j0 < M; j0 += 32 * CB) {  // 2 CB independent loads per lane and round trip
            uint32_t by[CB][2];
#pragma unroll
            for (int q = 0; q < CB; q++) {
                const int j = j0 + q * 32 + lane;
                by[q][0] = (j < mm[0]) ? (uint32_t)w.ldg(alt[0] + j) : (uint32_t)'A';  // columns past the shorter alternate: any class
                by[q][1] = (j < mm[1]) ? (uint32_t)w.ldg(alt[1] + j) : (uint32_t)'A';
            }
#pragma unroll
            for (int q = 0; q < CB; q++) {
                const int j = j0 + q * 32 + lane;
                const int c0 = sw_byte_class(by[q][0]), c1 = sw_byte_class(by[q][1]);
                other |= (c0 == SW_N_CLASSES) | (c1 == SW_N_CLASSES);
                if (j < M) acls[j] = (uint8_t)((c0 % SW_N_CLASSES) | ((c1 % SW_N_CLASSES) << 4));
            }
        }
        if (w.ballot(other) != 0u) return false;
    }
    LK_SW16_TICK(0);
    const uint32_t B = SW16_B;
    const uint32_t Dp = sw16_both((uint32_t)(we - wo)), Bp = sw16_both(B), Ca = sw16_both(B + (uint32_t)we);
    const uint32_t C7F = 0x7fff7fffu, CFL = sw16_both(0x7fffu - (uint32_t)(we - wo));
    const uint32_t one = A.one, mone = A.mone, two16 = A.two16, two = A.one + A.one;

    // ---- tables and column-0 state of this lane's rows ------------------------------------------------------------------
    uint32_t u[K], ny[K], even0[SW_N_CLASSES], even1[SW_N_CLASSES];
    w.sync();
    constexpr int RB = LK_SW16_ROW_BATCH < K ? LK_SW16_ROW_BATCH : K;  // rows per batch of reference loads: that many x 2 loads in flight, then their compares and table stores
#pragma unroll
    for (int k0 = 0; k0 < K; k0 += RB) {
        int rcls0[RB], rcls1[RB];  // class of a row in either reference (SW_N_CLASSES: no base, never matches)
#pragma unroll
        for (int kk = 0; kk < RB; kk++) {
            const int i0 = lane * K + k0 + kk + 1 - pad0, i1 = SAME_N ? i0 : lane * K + k0 + kk + 1 - pad1;
            const bool real0 = (k0 + kk < K) && (i0 >= 1 && i0 <= n0), real1 = SAME_N ? real0 : ((k0 + kk < K) && (i1 >= 1 && i1 <= n1));
            rcls0[kk] = real0 ? (int)w.ldg(ref[0] + (real0 ? i0 - 1 : 0)) : 0;
            rcls1[kk] = real1 ? (int)w.ldg(ref[1] + (real1 ? i1 - 1 : 0)) : 0;
        }
#pragma unroll
        for (int kk = 0; kk < RB; kk++) {
            const int k = k0 + kk;
            if (k < K) {
                // 1-based matrix row of either pair; < 1: virtual row (SoftClip / Ignore only)
                const int i0 = lane * K + k + 1 - pad0, i1 = SAME_N ? i0 : lane * K + k + 1 - pad1;
                const int32_t u0 = edge_gaps ? (i0 == 1 ? wo : we) : 0;  // sw[i][0] - sw[i-1][0] (:146-179); no padding when edge_gaps
                u[k < K ? k : 0] = sw16_both(B + (uint32_t)u0);
                ny[k < K ? k : 0] = Dp;                                  // best_gap_h starts at i32::MIN / 2: the first gap opens
                const int cls0 = sw_byte_class((uint32_t)rcls0[kk]), cls1 = sw_byte_class((uint32_t)rcls1[kk]);
#pragma unroll
                for (int c = 0; c < SW_N_CLASSES; c++) {
                    // a virtual row scores 0 against every column
                    const uint32_t s0 = (uint32_t)(((i0 < 1) ? 0 : ((cls0 == c) ? wm : wx)) + (int32_t)(2 * B));
                    const uint32_t s1 = (uint32_t)(((i1 < 1) ? 0 : ((cls1 == c) ? wm : wx)) + (int32_t)(2 * B));
                    uint32_t *tb1 = wtab + T::a_words;
                    if (k < 16) {
                        if (!LK_SW16_COMPACT) wtab[((c * QA + (k >> 2)) * 32 + lane) * 4 + (k & 3)] = s0;
                        if ((k & 1) == 0) { even0[c] = s0; even1[c] = s1; }
                        else {
                            if (LK_SW16_COMPACT) wtab[((c * QB + (k >> 3)) * 32 + lane) * 4 + ((k >> 1) & 3)] = even0[c] | (s0 << 16);
                            tb1[((c * QB + (k >> 3)) * 32 + lane) * 4 + ((k >> 1) & 3)] = even1[c] | (s1 << 16);
                        }
                    } else {
                        wtab[T::a_quad_words + c * 32 + lane] = s0;
                        tb1[T::b_quad_words + c * 32 + lane] = s1 << 16;
                    }
                }
            }
        }
    }
    w.sync();

    LK_SW16_TICK(1);
    // (Vu, Xu) = (v, nx) of the row above at this lane's column of the coming step.  Lane 0's come from the matrix top row:
    // v = 0, or open then extend for the edge-gap strategies (:146-179); nx = D (best_gap_v starts at i32::MIN / 2)
    const uint32_t vtop_rest = edge_gaps ? sw16_both(B + (uint32_t)we) : Bp;
    uint32_t Vu = (lane == 0 && edge_gaps) ? sw16_both(B + (uint32_t)wo) : Bp, Xu = Dp;
    // last column of each pair: this lane's rows, relative to the cell above its first row
    int32_t col_tot[2] = { 0, 0 }, col_best[2] = { (int32_t)(-2147483647 - 1), (int32_t)(-2147483647 - 1) };
    int col_row[2] = { 0, 0 };

    int cls_ahead = (lane == 0) ? (int)acls[0] : 0;  // column of step 1 (only lane 0 is active then)
    auto step = [&](int s, auto lean_tag) {
        constexpr bool LEAN = decltype(lean_tag)::value;
        const int c = s - lane;
        const bool active = LEAN || ((c >= 1) && (c <= M) && (lane < lanes));
        const int cls_now = cls_ahead;
        cls_ahead = (LEAN || (c >= 0 && c < M)) ? (int)acls[c] : 0;
        uint32_t v_last = Bp, x_last = Dp;
        if (active) {
            const int c0 = cls_now & 15, c1 = cls_now >> 4;
            constexpr int QA_ = LK_SW16_COMPACT ? QB : QA;
            uint32_t ta[QA * 4 + 1], tb[QB * 4 + 1];
            const uint32_t *t0 = wtab, *t1 = wtab + T::a_words;
#pragma unroll
            for (int q = 0; q < QA_; q++) w.lds4u(ta + 4 * q, t0 + ((c0 * QA_ + q) * 32 + lane) * 4);
#pragma unroll
            for (int q = 0; q < QB; q++) w.lds4u(tb + 4 * q, t1 + ((c1 * QB + q) * 32 + lane) * 4);
            if (HAS17) { ta[QA * 4] = t0[T::a_quad_words + c0 * 32 + lane]; tb[QB * 4] = t1[T::b_quad_words + c1 * 32 + lane]; }
            uint32_t vup = Vu, nxup = Xu;
            uint32_t acc_pd = 0u, acc_pr = 0u, acc_vo = 0u, acc_ho = 0u, nib = 0u;
            uint32_t e_pd = 0u, e_pr = 0u, e_vo = 0u, e_ho = 0u;  // flag words of the even row of a row pair
#pragma unroll
            for (int k = 0; k < K; k++) {
                // the cell's two weights in one word: pair 1's comes from the half of tb[k / 2] that row k owns
                const uint32_t sb = (k >= 16) ? sw16_fadd(ta[QA * 4], tb[QB * 4], one)
                                   : LK_SW16_COMPACT ? sw16_halves(ta[k >> 1], tb[k >> 1], k & 1)
                                              : ((k & 1) ? ((tb[k >> 1] & 0xffff0000u) | ta[k]) : sw16_fadd(tb[k >> 1], ta[k], two16));
                const uint32_t ab = (LK_SW16_ALU_MASK & 1) ? vup - nxup + Ca : sw16_fadd(sw16_fsub(vup, nxup, mone), Ca, one);
                const uint32_t bb = (LK_SW16_ALU_MASK & 2) ? u[k] - ny[k] + Ca : sw16_fadd(sw16_fsub(u[k], ny[k], mone), Ca, one);
                const uint32_t zb = sw16_max3(sb, ab, bb);
                const uint32_t un = sw16_fsub(zb, vup, mone), vn = sw16_fsub(zb, u[k], mone);
                const uint32_t nf = sw16_fsub(zb, ab, mone), ne = sw16_fsub(zb, bb, mone);
                const uint32_t d_pd = (LK_SW16_ALU_MASK & 4) ? zb - sb + C7F : sw16_fadd(sw16_fsub(zb, sb, mone), C7F, one);  // top bit: z > s, not diagonal (:248-252)
                const uint32_t d_pr = sw16_fadd(ne, C7F, one);    // top bit: b < z, i.e. vertical when not diagonal (:253-260)
                const uint32_t d_vo = sw16_fadd(nf, CFL, one);    // top bit: nf > D, the vertical gap of the row below opens here
                const uint32_t d_ho = sw16_fadd(ne, CFL, one);    // top bit: ne > D, the horizontal gap of the next column opens here
                nxup = sw16_min(nf, Dp);
                ny[k] = sw16_min(ne, Dp);
                u[k] = un;
                vup = vn;
                if (k < R16) {
                    if ((k & 1) == 0) { e_pd = d_pd; e_pr = d_pr; e_vo = d_vo; e_ho = d_ho; }
                    else {  // acc = 2 acc + (0xff per set flag): 255 x the plane word, undone by the traceback (SW16_INV255)
                        acc_pd = (k == 1) ? sw16_flags(e_pd, d_pd) : sw16_fadd(acc_pd, sw16_flags(e_pd, d_pd), two);
                        acc_pr = (k == 1) ? sw16_flags(e_pr, d_pr) : sw16_fadd(acc_pr, sw16_flags(e_pr, d_pr), two);
                        acc_vo = (k == 1) ? sw16_flags(e_vo, d_vo) : sw16_fadd(acc_vo, sw16_flags(e_vo, d_vo), two);
                        acc_ho = (k == 1) ? sw16_flags(e_ho, d_ho) : sw16_fadd(acc_ho, sw16_flags(e_ho, d_ho), two);
                    }
                } else {  // row 16: bits pd.0 pd.1 pr.0 pr.1 vo.0 vo.1 ho.0 ho.1 of one byte
                    const uint32_t t = (sw16_flags(d_pd, d_pr) & 0x08040201u) | (sw16_flags(d_vo, d_ho) & 0x80402010u);
                    nib = (t * 0x01010101u) >> 24;
                }
            }
            v_last = vup; x_last = nxup;
            w.store4(planes + ((int64_t)s * 32 + lane) * 4, acc_pd, acc_pr, acc_vo, acc_ho);
            if (HAS17) nibs[(int64_t)s * 32 + lane] = (uint8_t)nib;
            if (!LEAN) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    if (c == mm[h]) {  // this lane's rows of pair h's last column, top -> bottom with >= (:301-307)
                        // (a second reference length is re-read here rather than kept in a register across the lean loop)
                        const int n_h = SAME_N ? n0 : (int)(A.ref_off[pr[h] + 1] - A.ref_off[pr[h]]);
                        const int pad_h = SAME_N ? pad0 : (need_bottom ? BAND - n_h : 0);
                        int32_t run = 0;
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            const int i = lane * K + k + 1 - pad_h;
                            run += (int32_t)((u[k] >> (16 * h)) & 0xffffu) - (int32_t)B;
                            if (i >= 1 && i <= n_h && run >= col_best[h]) { col_best[h] = run; col_row[h] = i; }
                        }
                        col_tot[h] = run;
                    }
                }
            }
            if (need_bottom && lane == 31) botV[c] = v_last;  // the matrix's bottom row, as v; rebuilt after the fill (:316-329)
        }
        const uint32_t vs = w.shfl_up1(v_last), xs = w.shfl_up1(x_last);
        Vu = (lane != 0) ? vs : vtop_rest;
        Xu = (lane != 0) ? xs : Dp;
    };
    {
        const int n_steps = M + lanes - 1;
        const int s_lean_end = m_min - 1;  // last step at which no lane has reached either pair's last column
        int s = 1, gen_end = n_steps < 31 ? n_steps : 31;
        for (;;) {  // generic start-up, lean middle, generic tail
            for (; s <= gen_end; ++s) step(s, SwGenericTag());
            if (gen_end == n_steps) break;
LK_SW_PRAGMA(unroll LK_SW16_LEAN_UNROLL)
            for (; s <= s_lean_end; ++s) step(s, SwLeanTag());
            gen_end = n_steps;
        }
    }
    w.sync();
    LK_SW16_TICK(2);

    // ---- per pair: absolute last column and bottom row, then start cell, traceback and output as in sw_core.cuh ---------
    for (int h = 0; h < 2; h++) {
        const int m = mm[h], n = SAME_N ? n0 : (int)(A.ref_off[pr[h] + 1] - A.ref_off[pr[h]]), pad_h = SAME_N ? pad0 : (need_bottom ? BAND - n : 0);
        // last column: H(i, m) = H(0, m) + sum of u over the rows above and including i
        int32_t incl = col_tot[h];
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t t = w.shfl(incl, (lane - d) & 31);
            if (lane >= d) incl += t;
        }
        const int32_t top = edge_gaps ? (wo + (m - 1) * we) : 0;
        const int32_t col_max = (col_best[h] == (int32_t)(-2147483647 - 1)) ? col_best[h] : top + (incl - col_tot[h]) + col_best[h];
        int32_t bot_max = (int32_t)(-2147483647 - 1);
        int bot_j = 0;
        if (need_bottom) {  // H(n, c) = sum of v(n, .) up to c (column 0 of the bottom row is 0 for these strategies)
            int32_t carry = 0;
            for (int c0 = 1; c0 <= m; c0 += 32) {
                const int c = c0 + lane;
                int32_t val = (c <= m) ? (int32_t)((botV[c] >> (16 * h)) & 0xffffu) - (int32_t)B : 0;
                for (int d = 1; d < 32; d <<= 1) {
                    const int32_t t = w.shfl(val, (lane - d) & 31);
                    if (lane >= d) val += t;
                }
                if (c <= m) botH[c] = carry + val;
                carry += w.shfl(val, 31);
            }
            w.sync();
            sw_bottom_summary(w, botH, n, m, bot_max, bot_j);
        }
        LK_SW16_TICK(3);
        // backtrack bits of matrix cell (i, j), 1-based, of pair h
        auto nibble = [&](int i, int j) -> uint32_t {
            const int ii = i - 1 + pad_h;
            const int ln = ii / K, k = ii % K;
            const int64_t el = (int64_t)(j + ln) * 32 + ln;  // the cell was written at step j + ln by lane ln
            // both loads unconditionally and without a branch, so that the loads of consecutive calls overlap
            uint32_t a, b, c, d;
            w.load4(planes + el * 4, a, b, c, d);
            const bool odd_row = HAS17 && k == 16;
            const uint32_t nb = odd_row ? (uint32_t)nibs[el] : 0u;
            a *= SW16_INV255; b *= SW16_INV255; c *= SW16_INV255; d *= SW16_INV255;
            const int bit = ((k & 1) ? 16 : 0) + (h ? 8 : 0) + (R16 / 2 - 1 - ((k & 15) >> 1));  // the first row pair ends up highest
            const uint32_t pd = odd_row ? (nb >> h) & 1u : (a >> bit) & 1u, prb = odd_row ? (nb >> (2 + h)) & 1u : (b >> bit) & 1u;
            const uint32_t vo = odd_row ? (nb >> (4 + h)) & 1u : (c >> bit) & 1u, ho = odd_row ? (nb >> (6 + h)) & 1u : (d >> bit) & 1u;
            const uint32_t mv = pd ? (prb ? 2u : 1u) : 0u;
            return mv | (ho << 2) | (vo << 3);
        };
        sw_finish<1>(w, A, pr[h], n, m, col_max, col_row[h], bot_max, bot_j, nibble);
        w.sync();
        LK_SW16_TICK(4);
    }
    return true;
}

}  // namespace lk
