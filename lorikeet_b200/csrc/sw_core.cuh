// sw_core.cuh — Smith-Waterman (GATK single-matrix affine-gap formulation) with CIGAR traceback,
// one warp per (reference, alternate) pair.
//
// Bit-exact re-design of src/smith_waterman/smith_waterman_aligner.rs:
//   align            :47-113   (exact-substring shortcut for SoftClip / Ignore, :70-80,102-107)
//   calculate_matrix :124-263  (fill)
//   calculate_cigar  :273-442  (start-cell search, traceback, overhang handling)
//
// Layout: rows = reference positions, columns = alternate positions, exactly as the reference.
// Each lane owns K consecutive rows; the warp sweeps the columns as a skewed wavefront (lane l is
// at column s-l at step s) and hands (H, F) of its last row to the next lane by shuffle.  Rows are
// processed in bands of 32*K; a band's last row travels through per-warp scratch.
//
// The reference keeps two (n+1)x(m+1) i32 matrices (sw, btrack).  Here nothing of `sw` is stored
// except the band boundary; `btrack` is compressed to 4 bits per cell, kept as four bit-planes
//     pd  set when the move is NOT diagonal (step_diag < max(step_down, step_right), :248-252)
//     pr  set when step_right < step_down: vertical; clear: horizontal (:253-260)
//     ho  a horizontal gap was (re)opened at this cell  (gap_size_h[i] reset to 1, :235-239)
//     vo  a vertical gap was (re)opened at this cell    (gap_size_v[j] reset to 1, :212-218)
// from which the stored gap length k of the reference's btrack (-ki / +kd) is recovered during the
// traceback by walking back to the cell where the gap was opened.  The planes of one (step, lane) are packed into
// TWO words -- (pd, pr) interleaved in one, (vo, ho) in the other, two bits per row for the lane's first 16 rows --
// stored step-major ([step][lane], 8 bytes each) so every warp store is one coalesced 256-byte run; a 17th row
// (K = 17, the variant 520-row haplotype pairs use) adds one byte per (step, lane) in a second step-major array.
// That is 0.53 B/cell of backtrack traffic at K = 17 (the first version stored four 32-bit planes: 0.94 B/cell).
// Per cell: compare+select+add (diagonal), add + VIADDMNMX + compare (each gap direction),
// max + 2 compares + VIMNMX3 (choice, the -1e8 clamp rides in the 3-input max), 4 predicated ORs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define LK_DEV __device__ __forceinline__
#define LK_HD __host__ __device__ __forceinline__
#else
#define LK_DEV inline
#define LK_HD inline
#endif

#ifndef LK_SW_LEAN_UNROLL
#define LK_SW_LEAN_UNROLL 1
#endif
// independent loads per lane and round trip of the traceback's diagonal scan
#ifndef LK_SW_TRACE_LOADS
#define LK_SW_TRACE_LOADS 2
#endif
#define LK_SW_PRAGMA_(x) _Pragma(#x)
#define LK_SW_PRAGMA(x) LK_SW_PRAGMA_(x)

namespace lk {

enum : int { SW_SOFTCLIP = 0, SW_INDEL = 1, SW_LEADING_INDEL = 2, SW_IGNORE = 3 };
enum : uint32_t { CIG_M = 0, CIG_I = 1, CIG_D = 2, CIG_S = 4 };
constexpr int32_t SW_MATRIX_MIN_CUTOFF = -100000000;            // smith_waterman_aligner.rs:31
constexpr int32_t SW_LOW_INIT = (int32_t)(-2147483647 - 1) / 2;  // i32::MIN / 2, :139

struct SwArgs {
    const uint8_t *refs, *alts;
    const int64_t *ref_off, *alt_off;
    const int32_t *order;      // pair ids this launch works through (dynamic, via cursor)
    int32_t n_order;
    int32_t *cursor;
    int32_t *packed_count;     // pairs the packed kernel (sw16_core.cuh) took, for introspection
    unsigned long long *prof;  // developer builds: clock ticks per phase (LK_SW16_PROFILE)
    int32_t w_match, w_mismatch, w_open, w_extend;
    int32_t strategy;
    int32_t max_ops;
    uint32_t *cigar_ops;       // [pair][max_ops]
    int32_t *n_ops;            // [pair]
    int32_t *offsets;          // [pair]
    uint32_t *scratch;         // per-warp backtrack + band boundary
    int64_t scratch_words;     // words per warp
    uint32_t one, mone, two16; // 1, 0xffffffff and 65536 as run-time values (sw16_core.cuh: adds issued as IMADs)
};

LK_DEV uint32_t sw_cigar_el(uint32_t op, int32_t len) { return ((uint32_t)len << 4) | op; }

LK_DEV int32_t sw_addmax(int32_t a, int32_t b, int32_t c) {  // max(a + b, c): one VIADDMNMX on sm_100a
#if defined(__CUDA_ARCH__)
    return __viaddmax_s32(a, b, c);
#else
    return a + b > c ? a + b : c;
#endif
}
LK_DEV int32_t sw_max3(int32_t a, int32_t b, int32_t c) {    // one VIMNMX3
#if defined(__CUDA_ARCH__)
    return __vimax3_s32(a, b, c);
#else
    const int32_t t = a > b ? a : b;
    return t > c ? t : c;
#endif
}
LK_DEV uint32_t sw_push_sign(uint32_t plane, int32_t d) {      // (plane << 1) | (d < 0): one SHF
#if defined(__CUDA_ARCH__)
    return __funnelshift_l((uint32_t)d, plane, 1);
#else
    return (plane << 1) | ((uint32_t)d >> 31);
#endif
}
// ---- substitution-weight table --------------------------------------------------------------
// The match / mismatch weight of a cell, (ref[i] == alt[j]) ? w_match : w_mismatch, costs a compare and a select per cell
// on the 16-lane ALU pipe.  Instead each lane keeps, in per-warp shared memory, the weight of each of its K rows against
// each base an alternate column can hold: tab[class][k/4][lane] is an int4, so a step fetches its K weights with
// ceil(K/4) conflict-free LDS.128 at an address picked by the column's byte.  Classes are the exact bytes 'A' 'C' 'G' 'T'
// 'N' (the reference compares raw bytes, smith_waterman_aligner.rs:197-202); a pair whose alternate holds any other byte
// takes the compare-and-select path (decided once per pair, warp-uniform), so nothing is assumed about the alphabet.
constexpr int SW_N_CLASSES = 5;
LK_DEV int sw_byte_class(uint32_t b) {  // 0..4, or SW_N_CLASSES for "some other byte"
    switch (b) {
        case 'A': return 0;
        case 'C': return 1;
        case 'G': return 2;
        case 'T': return 3;
        case 'N': return 4;
        default: return SW_N_CLASSES;
    }
}
template <int K> struct SwTable { static constexpr int words = SW_N_CLASSES * ((K + 3) / 4) * 32 * 4; };

// H[k] for a warp-uniform runtime k (a switch keeps it to one branch instead of a K-long select chain)
template <int K>
LK_DEV int32_t sw_pick(const int32_t (&H)[K], int k) {
    int32_t v = H[0];
    switch (k) {
#define LK_PICK(i) case i: if (i < K) v = H[i < K ? i : 0]; break;
        LK_PICK(1) LK_PICK(2) LK_PICK(3) LK_PICK(4) LK_PICK(5) LK_PICK(6) LK_PICK(7) LK_PICK(8) LK_PICK(9) LK_PICK(10)
        LK_PICK(11) LK_PICK(12) LK_PICK(13) LK_PICK(14) LK_PICK(15) LK_PICK(16) LK_PICK(17) LK_PICK(18) LK_PICK(19)
        LK_PICK(20) LK_PICK(21) LK_PICK(22) LK_PICK(23) LK_PICK(24) LK_PICK(25) LK_PICK(26) LK_PICK(27) LK_PICK(28)
        LK_PICK(29) LK_PICK(30) LK_PICK(31)
#undef LK_PICK
        default: break;
    }
    return v;
}

// number of backtrack words one pair needs (host + device agree on this): per band and (step, lane) two plane words,
// plus one byte when K > 16
LK_HD int64_t sw_band_words(int m, int K) {
    const int64_t steps = (int64_t)m + 32;
    return steps * 32 * 2 + (K > 16 ? steps * 8 : 0);
}
LK_HD int64_t sw_bt_words(int n, int m, int K) {
    const int band_rows = 32 * K;
    const int n_bands = (n + band_rows - 1) / band_rows;
    return (int64_t)n_bands * sw_band_words(m, K);
}

// CLAMP = false drops the max(-1e8, .) of smith_waterman_aligner.rs:31,248-260 from the cell: the host proves per launch
// that no score can get that low ((n + m + 2) * max|weight| < 1e8, sw_clamp_inactive), which is every realistic input.
LK_HD bool sw_clamp_inactive(int64_t nm_max, int32_t wm, int32_t wx, int32_t wo, int32_t we) {
    int64_t a = wm < 0 ? -(int64_t)wm : wm, b = wx < 0 ? -(int64_t)wx : wx, c = wo < 0 ? -(int64_t)wo : wo, d = we < 0 ? -(int64_t)we : we;
    int64_t mx = a > b ? a : b;
    mx = mx > c ? mx : c;
    mx = mx > d ? mx : d;
    return (nm_max + 2) * mx < 100000000LL;
}

// words of per-warp scratch one pair needs: backtrack planes, two band-boundary rows, the bottom row, one class byte per
// alternate base (even, so that every warp's slice keeps the 8-byte alignment of the plane stores)
LK_HD int64_t sw_scratch_words(int n, int m, int K) {
    const int64_t w = sw_bt_words(n, m, K) + 3 * ((int64_t)m + 1) + ((int64_t)m + 3) / 4 + 36;
    return (w + 1) & ~(int64_t)1;
}
struct SwLeanTag { static constexpr bool value = true; };
struct SwGenericTag { static constexpr bool value = false; };

// Empty inputs (:65-68) and the exact-substring shortcut, AlignmentUtils::last_index_of (alignment_utils.rs:717-735): true
// when the pair is finished without a matrix.
template <class Wp>
LK_DEV bool sw_prologue(Wp &w, const SwArgs &A, int pair) {
    const int lane = w.lane();
    const uint8_t *ref = A.refs + A.ref_off[pair];
    const uint8_t *alt = A.alts + A.alt_off[pair];
    const int n = (int)(A.ref_off[pair + 1] - A.ref_off[pair]);
    const int m = (int)(A.alt_off[pair + 1] - A.alt_off[pair]);
    uint32_t *ops_out = A.cigar_ops + (int64_t)pair * A.max_ops;
    if (n <= 0 || m <= 0) {  // :65-68 assert
        if (lane == 0) { A.n_ops[pair] = -1; A.offsets[pair] = 0; }
        return true;
    }
    const int strategy = A.strategy;
    if ((strategy == SW_SOFTCLIP || strategy == SW_IGNORE) && m <= n) {
        int found = -1;
        for (int base = n - m; base >= 0 && found < 0; base -= 32) {
            const int r = base - lane;
            bool ok = r >= 0;
            if (ok) {
                int q = 0;
                while (q < m && ref[r + q] == alt[q]) q++;
                ok = (q == m);
            }
            const unsigned b = w.ballot(ok);
            if (b) found = base - (w.ffs(b) - 1);  // lowest lane = largest r
        }
        if (found >= 0) {
            if (lane == 0) {
                if (A.max_ops < 1) A.n_ops[pair] = -2;
                else { ops_out[0] = sw_cigar_el(CIG_M, m); A.n_ops[pair] = 1; }
                A.offsets[pair] = found;
            }
            return true;
        }
    }
    return false;
}

// ---- bottom row summary (:316-329): the reference scans it left->right taking a cell if it is strictly better, or equal
// and closer to the diagonal than the current best.  Its outcome depends only on the row maximum and on the maximal cell
// closest to the diagonal (leftmost on ties): every lane scans its columns in that order, then the lanes are combined. ----
template <class Wp>
LK_DEV void sw_bottom_summary(Wp &w, const int32_t *botH, int n, int m, int32_t &bot_max, int &bot_j) {
    const int lane = w.lane();
    bot_max = (int32_t)(-2147483647 - 1);
    bot_j = 0;
    for (int c = lane + 1; c <= m; c += 32) {
        const int32_t cur = botH[c];
        const int dj = (n > c) ? (n - c) : (c - n), db = (n > bot_j) ? (n - bot_j) : (bot_j - n);
        if (cur > bot_max || (cur == bot_max && dj < db)) { bot_max = cur; bot_j = c; }
    }
    for (int d = 16; d >= 1; d >>= 1) {
        const int32_t ov = w.shfl_xor(bot_max, d);
        const int oj = w.shfl_xor(bot_j, d);
        const int dj = (n > oj) ? (n - oj) : (oj - n), db = (n > bot_j) ? (n - bot_j) : (bot_j - n);
        if (ov > bot_max || (ov == bot_max && (dj < db || (dj == db && oj < bot_j)))) { bot_max = ov; bot_j = oj; }
    }
}

// ---- start cell (:288-333), traceback (:335-442) and output.  col_max / col_p1: every lane's best cell of the last column
// among its own rows (scanned top -> bottom with >=); bot_max / bot_j: the bottom-row summary.  nibble(i, j) returns the
// backtrack bits stored for matrix cell (i, j), 1-based: bits 0-1 move (0 diagonal, 1 horizontal, 2 vertical), bit 2 ho,
// bit 3 vo.  FLAG_SHIFT = 0: ho / vo describe the gap ENDING at the cell they are stored with; 1: the horizontal gap of
// (i, j + 1) / the vertical gap of (i + 1, j) (the packed kernel computes them one cell early). ----
template <int FLAG_SHIFT, class Wp, class Nib>
LK_DEV void sw_finish(Wp &w, const SwArgs &A, int pair, int n, int m, int32_t col_max, int col_p1, int32_t bot_max, int bot_j,
                      Nib nibble) {
    const int lane = w.lane();
    const int strategy = A.strategy;
    uint32_t *ops_out = A.cigar_ops + (int64_t)pair * A.max_ops;
    // ---- start cell (:288-333) -------------------------------------------------------------------
    // last column: every lane scanned its own rows with >=; combine so the LARGEST row among maxima wins
    for (int d = 16; d >= 1; d >>= 1) {
        const int32_t ov = w.shfl_xor(col_max, d);
        const int op = w.shfl_xor(col_p1, d);
        if (ov > col_max || (ov == col_max && op > col_p1)) { col_max = ov; col_p1 = op; }
    }
    int p1, p2;
    int32_t segment_length = 0;
    if (strategy == SW_INDEL) { p1 = n; p2 = m; }
    else {
        p1 = col_p1; p2 = m;
        if (strategy != SW_LEADING_INDEL) {
            // The reference scans the bottom row left->right taking a cell if it is strictly better, or equal and
            // closer to the diagonal than the current best.  Its outcome depends only on the row maximum and on the
            // maximal cell closest to the diagonal (leftmost on ties), which is what (bot_max, bot_j) track.
            const int d_best = (p1 > p2) ? (p1 - p2) : (p2 - p1);
            const int d_bot = (n > bot_j) ? (n - bot_j) : (bot_j - n);
            if (bot_max > col_max || (bot_max == col_max && d_bot < d_best)) {
                p1 = n; p2 = bot_j;
                segment_length = m - bot_j;
            }
        }
    }

    // ---- traceback (:335-442), warp-cooperative: 32 cells of a run are fetched per round trip -----
    int n_out = 0;
    bool overflow = false;
    auto push = [&](uint32_t el) {
        if (n_out < A.max_ops) { if (lane == 0) ops_out[n_out] = el; }
        else overflow = true;
        n_out++;
    };
    if (segment_length > 0 && strategy == SW_SOFTCLIP) { push(sw_cigar_el(CIG_S, segment_length)); segment_length = 0; }
    int state = 0;  // 0 match, 1 insertion, 2 deletion
    auto op_of = [](int st) -> uint32_t { return st == 0 ? CIG_M : (st == 1 ? CIG_I : CIG_D); };
    for (;;) {
        // how many consecutive diagonal moves start at (p1, p2)?
        const int lim = (p1 < p2) ? p1 : p2;  // cells available along the diagonal
        int run = 0;
        uint32_t stop_nib = 0;
        for (;;) {  // 32 NL cells of the diagonal per round trip: the NL loads of a lane are independent
            constexpr int NL = LK_SW_TRACE_LOADS;
            uint32_t nb[NL];
            unsigned nd[NL], any = 0u;
#pragma unroll
            for (int q = 0; q < NL; q++) {
                const int t = run + 32 * q + lane;
                nb[q] = (t < lim) ? nibble(p1 - t, p2 - t) : 0xFFu;
            }
#pragma unroll
            for (int q = 0; q < NL; q++) {
                const int t = run + 32 * q + lane;
                nd[q] = w.ballot(t >= lim || (nb[q] & 3u) != 0u);
                any |= nd[q];
            }
            if (any == 0u) { run += 32 * NL; continue; }
            bool found = false;
#pragma unroll
            for (int q = 0; q < NL; q++) {
                if (!found && nd[q] != 0u) {
                    const int f = w.ffs(nd[q]) - 1;
                    run += 32 * q + f;
                    stop_nib = w.shfl(nb[q], f);
                    found = true;
                }
            }
            break;
        }
        if (run > 0) {  // `run` diagonal steps
            if (state == 0) segment_length += run;
            else { if (segment_length > 0) push(sw_cigar_el(op_of(state), segment_length)); segment_length = run; state = 0; }
            p1 -= run; p2 -= run;
            if (p1 <= 0 || p2 <= 0) break;
        }
        // (p1, p2) holds a gap move; recover its length = distance back to the cell where the gap was opened
        const uint32_t mv = stop_nib & 3u;
        int glen = 0;
        for (;;) {
            const int t = glen + lane;
            bool opened;
            if (mv == 1u) { opened = (p2 - t <= 1) ? true : ((nibble(p1, p2 - t - FLAG_SHIFT) & 4u) != 0u); }
            else { opened = (p1 - t <= 1) ? true : ((nibble(p1 - t - FLAG_SHIFT, p2) & 8u) != 0u); }
            const unsigned b = w.ballot(opened);
            if (b == 0) { glen += 32; continue; }
            glen += w.ffs(b);  // inclusive
            break;
        }
        const int new_state = (mv == 1u) ? 1 : 2;
        if (new_state == 1) p2 -= glen; else p1 -= glen;
        if (new_state == state) segment_length += glen;
        else { if (segment_length > 0) push(sw_cigar_el(op_of(state), segment_length)); segment_length = glen; state = new_state; }
        if (p1 <= 0 || p2 <= 0) break;
    }
    int32_t alignment_offset;
    if (strategy == SW_SOFTCLIP) {          // :413-419
        push(sw_cigar_el(op_of(state), segment_length));
        if (p2 > 0) push(sw_cigar_el(CIG_S, p2));
        alignment_offset = p1;
    } else if (strategy == SW_IGNORE) {     // :420-423
        push(sw_cigar_el(op_of(state), segment_length + p2));
        alignment_offset = p1 - p2;
    } else {                                // :424-436
        push(sw_cigar_el(op_of(state), segment_length));
        if (p1 > 0) push(sw_cigar_el(CIG_D, p1));
        else if (p2 > 0) push(sw_cigar_el(CIG_I, p2));
        alignment_offset = 0;
    }
    w.sync();
    if (lane == 0) {
        if (overflow) A.n_ops[pair] = -2;
        else {
            for (int a = 0, b = n_out - 1; a < b; a++, b--) { const uint32_t t = ops_out[a]; ops_out[a] = ops_out[b]; ops_out[b] = t; }  // :440
            A.n_ops[pair] = n_out;
        }
        A.offsets[pair] = alignment_offset;
    }
}

template <class Wp, int K, bool CLAMP = true>
LK_DEV void sw_dp_pair(Wp &w, const SwArgs &A, int pair, uint32_t *scratch, int32_t *wtab) {
    constexpr int KQ = (K + 3) / 4;
    const int lane = w.lane();
    const uint8_t *ref = A.refs + A.ref_off[pair];
    const uint8_t *alt = A.alts + A.alt_off[pair];
    const int n = (int)(A.ref_off[pair + 1] - A.ref_off[pair]);
    const int m = (int)(A.alt_off[pair + 1] - A.alt_off[pair]);
    const int strategy = A.strategy;

    // ---- fill -----------------------------------------------------------------------------------
    const int32_t wm = A.w_match, wx = A.w_mismatch, wo = A.w_open, we = A.w_extend;
    const bool edge_gaps = (strategy == SW_INDEL || strategy == SW_LEADING_INDEL);  // :146-179
    const bool need_bottom = (strategy == SW_SOFTCLIP || strategy == SW_IGNORE);    // the bottom row is searched only then (:316)
    const int BAND = 32 * K;
    const int n_bands = (n + BAND - 1) / BAND;
    static_assert(K <= 17, "two plane bits per row in one word for 16 rows, one extra byte for a 17th");
    constexpr int PR = K < 16 ? K : 16;  // rows of a lane whose bits live in the two plane words
    constexpr bool HAS_NIB = K > 16;     // row 16 of a lane: one byte per (step, lane)
    const int64_t band_words = sw_band_words(m, K);
    uint32_t *bt = scratch;                                         // [band]: [step][lane] x 2 words, then [step][lane] bytes
    int32_t *bndH = (int32_t *)(scratch + sw_bt_words(n, m, K));    // band-boundary row: H then F
    int32_t *bndF = bndH + (m + 1);
    int32_t *botH = bndF + (m + 1);                                 // bottom row of the matrix, one value per column
    // class of every alternate base, once per pair; the weight table is used when all of them are A C G T N (warp-uniform)
    uint8_t *acls = (uint8_t *)(botH + (m + 1));
    bool use_table;
    {
        bool other = false;
        for (int j0 = 0; j0 < m; j0 += 128) {  // four independent loads per lane and round trip
            uint32_t by[4];
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int j = j0 + q * 32 + lane;
                by[q] = (j < m) ? (uint32_t)w.ldg(alt + j) : (uint32_t)'A';
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int j = j0 + q * 32 + lane;
                const int cl = sw_byte_class(by[q]);
                other |= (cl == SW_N_CLASSES);
                if (j < m) acls[j] = (uint8_t)(cl == SW_N_CLASSES ? 0 : cl);
            }
        }
        use_table = w.ballot(other) == 0u;
        w.sync();
    }
    // best cell of the last column (scan top->bottom with >=, :301-307)
    int32_t col_max = (int32_t)(-2147483647 - 1);
    int col_p1 = 0;
    // Bottom alignment (the strategies that search the bottom row, SoftClip / Ignore): the rows are shifted down so that the
    // matrix's LAST row is always register K-1 of lane 31 of the last band -- its value is then the `h_last` the step hands
    // on anyway, instead of a register picked by a run-time index.  The pad_top rows this leaves above row 1 are VIRTUAL rows
    // with weight 0 against every column: they reproduce the top row exactly (H = max(0 + 0, open, open) = 0 for open < 0) and
    // hand row 1 a vertical-gap state of `open` instead of i32::MIN/2, which row 1 cannot tell apart as long as extend < 0
    // (max(0 + open, open + extend) = open = max(0 + open, MIN/2 + extend), gap opened either way).  The traceback never enters
    // them (it stops at row 0).
    // Only when it costs at most two extra lanes of skew (a reference that nearly fills its 32 x K rows, e.g. 520 of 544): a short
    // reference would pay more in additional wavefront steps than the picked register costs.
    const bool bottom_align = need_bottom && use_table && wo < 0 && we < 0 && (n_bands * BAND - n) < 2 * K;
    const int pad_top = bottom_align ? n_bands * BAND - n : 0;

    for (int band = 0; band < n_bands; band++) {
        const int row0 = band * BAND - pad_top;  // rows row0+1 .. row0+rows (1-based matrix rows; <= 0: virtual rows above row 1)
        const int rows = (n - row0 < BAND) ? (n - row0) : BAND;
        const int lanes = (rows + K - 1) / K;
        uint32_t rb[K];
        int32_t H[K], E[K];
#pragma unroll
        for (int k = 0; k < K; k++) {
            const int i = row0 + lane * K + k + 1;  // 1-based matrix row
            rb[k] = (i >= 1 && i <= n) ? (uint32_t)ref[i - 1] : (i < 1 ? 0x200u : 0x100u);  // 0x200: virtual row, 0x100: row past the reference
            H[k] = edge_gaps ? (wo + (i - 1) * we) : 0;  // sw[i][0]
            E[k] = SW_LOW_INIT;                           // best_gap_h[i]
        }
        // substitution weights of this lane's rows against each alternate class (rows past the reference never match)
        w.sync();
#pragma unroll
        for (int k = 0; k < KQ * 4; k++) {
            const int cls_of_row = (k < K) ? sw_byte_class(rb[k < K ? k : 0]) : SW_N_CLASSES;
            const bool virt = (k < K) && rb[k < K ? k : 0] == 0x200u;  // a virtual row scores 0 against every column
#pragma unroll
            for (int c = 0; c < SW_N_CLASSES; c++) wtab[((c * KQ + (k >> 2)) * 32 + lane) * 4 + (k & 3)] = virt ? 0 : ((cls_of_row == c) ? wm : wx);
        }
        w.sync();
        // column 0 of the row above this lane's strip
        const int i_above = row0 + lane * K;  // matrix row index of the row above (<= 0: the top row or a virtual row, both 0 in column 0)
        const int32_t hd0 = (i_above <= 0) ? 0 : (edge_gaps ? (wo + (i_above - 1) * we) : 0);
        int32_t Hd = hd0;
        // (Hu, Fu) = the row above at this lane's column of the coming step.  Lane 0's comes from the matrix top row (band 0:
        // sw[0][c] = 0, or open + (c-1) extend for the edge-gap strategies, :146-179) or from the previous band's last row.
        int32_t Hu = 0, Fu = SW_LOW_INIT;
        const int32_t hu_inc = (band == 0 && edge_gaps) ? we : 0;
        if (lane == 0) {
            if (band == 0) Hu = edge_gaps ? wo : 0;
            else { Hu = bndH[1]; Fu = bndF[1]; }
        }
        uint32_t *btb = bt + (int64_t)band * band_words;
        uint8_t *btn = (uint8_t *)(btb + ((int64_t)m + 32) * 64);  // row-16 nibbles of the band (K = 17)
        const bool has_bottom = (band == n_bands - 1);
        const int bottom_lane = (rows - 1) / K, bottom_k = (rows - 1) % K;
        const bool keeps_bottom = need_bottom && has_bottom && lane == bottom_lane;
        const bool writes_bnd = !has_bottom && lane == lanes - 1;
        const int n_steps = m + lanes - 1;

        int cls_ahead = (use_table && lane == 0) ? (int)acls[0] : 0;  // column of step 1 (only lane 0 is active then)
        // One step of the wavefront.  LEAN: every lane's column is inside [1, m-1] (steps 32 .. m-1), so nothing is gated; lanes
        // past the reference's last row then run on rows that never match -- their cells are never read.
        auto step = [&](int s, auto lean_tag, bool band0) {
            constexpr bool LEAN = decltype(lean_tag)::value;
            const int c = s - lane;
            const bool active = LEAN || ((c >= 1) && (c <= m) && (lane < lanes));
            // class of the NEXT step's column, fetched one step ahead so its latency hides behind this step's cells
            const int cls_now = cls_ahead;
            if (use_table) cls_ahead = (LEAN || (c >= 0 && c < m)) ? (int)acls[c] : 0;
            int32_t h_last = 0, f_last = SW_LOW_INIT;
            if (active) {
                int32_t wt[KQ * 4];
                if (use_table) {
#pragma unroll
                    for (int q = 0; q < KQ; q++) w.lds4i(wt + 4 * q, wtab + ((cls_now * KQ + q) * 32 + lane) * 4);
                } else {  // some alternate byte is outside ACGTN: raw byte compare for this pair, as the reference does everywhere
                    const uint32_t ab = w.ldg(alt + c - 1);
#pragma unroll
                    for (int k = 0; k < K; k++) wt[k] = (rb[k] == ab) ? wm : wx;
                }
                int32_t hdiag = Hd, hup = Hu, fup = Fu;
                uint32_t pw_m = 0u, pw_g = 0u, nib = 0u;  // (pd, pr) pairs; (vo, ho) pairs; row 16's four bits
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const int32_t hold = H[k];
                    const int32_t step_diag = hdiag + wt[k];                         // :197-202
                    const int32_t ext_v = fup + we;                                  // :210-222
                    const int32_t step_down = sw_addmax(hup, wo, ext_v);             // max(hup + wo, ext_v)
                    const int32_t ext_h = E[k] + we;                                 // :232-245
                    const int32_t step_right = sw_addmax(hold, wo, ext_h);
                    // choice: diagonal if >= both gaps, else the larger gap (:248-260).  Without the clamp one 3-input max gives
                    // the cell and "not diagonal" is best > step_diag; with it the 2-input max of the gaps is needed as well.
                    int32_t best, not_diag;
                    if (CLAMP) {
                        const int32_t gap = step_right > step_down ? step_right : step_down;
                        best = sw_max3(step_diag, gap, SW_MATRIX_MIN_CUTOFF);
                        not_diag = step_diag - gap;
                    } else {
                        best = sw_max3(step_diag, step_right, step_down);
                        not_diag = step_diag - best;
                    }
                    // each plane bit is the SIGN BIT of a difference, shifted in by one funnel shift (the subtractions can issue
                    // on the FMA pipe): row k ends up at bits 2 (PR-1-k) + {1, 0} of its word
                    if (k < 16) {
                        pw_g = sw_push_sign(pw_g, ext_v - step_down);      // vo set: opened (hup + wo > fup + we)
                        pw_g = sw_push_sign(pw_g, ext_h - step_right);     // ho set: opened
                        pw_m = sw_push_sign(pw_m, not_diag);               // pd set: NOT diagonal (:248-252 wants diag >= both)
                        pw_m = sw_push_sign(pw_m, step_right - step_down); // pr set: vertical (right < down), clear: horizontal
                    } else {
                        nib = sw_push_sign(nib, ext_v - step_down);
                        nib = sw_push_sign(nib, ext_h - step_right);
                        nib = sw_push_sign(nib, not_diag);
                        nib = sw_push_sign(nib, step_right - step_down);
                    }
                    E[k] = step_right;
                    H[k] = best;
                    hdiag = hold; hup = best; fup = step_down;
                }
                h_last = H[K - 1]; f_last = fup;
                w.store2(btb + ((int64_t)s * 32 + lane) * 2, pw_m, pw_g);
                if (HAS_NIB) btn[(int64_t)s * 32 + lane] = (uint8_t)nib;
                if (!LEAN && c == m) {  // this lane's rows of the last column, top -> bottom with >=  (:301-307)
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        const int i = row0 + lane * K + k + 1;
                        if (i >= 1 && i <= n && H[k] >= col_max) { col_max = H[k]; col_p1 = i; }
                    }
                }
                // the matrix's bottom row goes to scratch, one value per column; its summary (:316-329) is taken after the fill
                if (keeps_bottom) botH[c] = bottom_align ? h_last : sw_pick<K>(H, bottom_k);
                if (writes_bnd) { bndH[c] = h_last; bndF[c] = f_last; }
            }
            // hand the strip's last row to the next lane; keep this step's "up" as next step's diagonal
            Hd = Hu;
            const int32_t hs = w.shfl_up1(h_last), fs = w.shfl_up1(f_last);
            if (band0) {
                Hu = (lane != 0) ? hs : Hu + hu_inc;  // lane 0: next column of the top row
                Fu = (lane != 0) ? fs : Fu;
            } else {
                int32_t lh = hs, lf = fs;
                if (lane == 0 && (LEAN || s + 1 <= m)) { lh = bndH[s + 1]; lf = bndF[s + 1]; }  // previous band's last row
                Hu = lh; Fu = lf;
            }
            // a lane that has not started yet must keep seeing column 0 of the row above as its diagonal
            if (!LEAN && c < 1) Hd = hd0;
        };
        const int s_lean_end = m - 1;  // last step at which lane 0's column is still < m
        int s = 1, gen_end = n_steps < 31 ? n_steps : 31;
        for (;;) {  // generic start-up, lean middle, generic tail (one copy of the generic loop in the code)
            for (; s <= gen_end; ++s) step(s, SwGenericTag(), band == 0);
            if (gen_end == n_steps) break;
            if (band == 0) {
LK_SW_PRAGMA(unroll LK_SW_LEAN_UNROLL)
                for (; s <= s_lean_end; ++s) step(s, SwLeanTag(), true);
            } else {
                for (; s <= s_lean_end; ++s) step(s, SwLeanTag(), false);
            }
            gen_end = n_steps;
        }
        w.sync();
    }

    int32_t bot_max = (int32_t)(-2147483647 - 1);
    int bot_j = 0;
    if (need_bottom) sw_bottom_summary(w, botH, n, m, bot_max, bot_j);

    // nibble of matrix cell (i, j), 1-based: bits 0-1 move (0 diagonal, 1 horizontal, 2 vertical), bit 2 ho, bit 3 vo
    auto nibble = [&](int i, int j) -> uint32_t {
        const int ii = i - 1 + pad_top;  // position among the rows the bands cover (virtual rows first)
        const int band = ii / BAND, r = ii % BAND, ln = r / K, k = r % K;
        const uint32_t *bb = bt + (int64_t)band * band_words;
        const int64_t el = (int64_t)(j + ln) * 32 + ln;  // the cell was written at step j + ln by lane ln
        uint32_t pd, pr, ho, vo;
        if (HAS_NIB && k == 16) {
            const uint32_t nb = ((const uint8_t *)(bb + ((int64_t)m + 32) * 64))[el];
            vo = (nb >> 3) & 1u; ho = (nb >> 2) & 1u; pd = (nb >> 1) & 1u; pr = nb & 1u;
        } else {
            uint32_t pw_m, pw_g;
            w.load2(bb + el * 2, pw_m, pw_g);
            const int sh = 2 * (PR - 1 - k);  // the words were filled by shifting left twice per row
            pd = (pw_m >> (sh + 1)) & 1u; pr = (pw_m >> sh) & 1u;
            vo = (pw_g >> (sh + 1)) & 1u; ho = (pw_g >> sh) & 1u;
        }
        const uint32_t mv = pd ? (pr ? 2u : 1u) : 0u;
        return mv | (ho << 2) | (vo << 3);
    };
    sw_finish<0>(w, A, pair, n, m, col_max, col_p1, bot_max, bot_j, nibble);
}

template <class Wp, int K, bool CLAMP = true>
LK_DEV void sw_align_pair(Wp &w, const SwArgs &A, int pair, uint32_t *scratch, int32_t *wtab) {
    if (sw_prologue(w, A, pair)) return;
    sw_dp_pair<Wp, K, CLAMP>(w, A, pair, scratch, wtab);
}

}  // namespace lk
