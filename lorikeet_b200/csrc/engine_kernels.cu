// engine_kernels.cu — device side of PairHMMLikelihoodCalculationEngine's pre- and post-ops, so the
// region driver takes raw reads and returns normalised, filtered [allele][read] matrices without host passes.
//
//   engine_modify_reads_kernel   modify_read_qualities        pair_hmm_likelihood_calculation_engine.rs:352-426
//                                  apply_pcr_error_model        :502-526, find_tandem_repeat_units :528-611,
//                                  find_number_of_repetitions   variant_context_utils.rs:276-337
//                                  cap_minimum_read_qualities   :428-466
//   engine_epilogue_kernel       [r*A + h] -> [a][r]            pair_hmm.rs:247-266
//                                  normalize_likelihoods        allele_likelihoods.rs:378-444 (search_best_allele :457-554)
//                                  thresholds                   ...engine.rs:244-319 (table :23-39), from the ORIGINAL qualities
//                                  filter_poorly_modeled_evidence / remove_evidence_by_index   allele_likelihoods.rs:925-1018
//
// All byte / f64 compare work, HBM-bound and tiny next to the PairHMM kernels; the f64 arithmetic uses explicit
// un-fused operations so every threshold rounds exactly like the reference's.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "best_allele.h"
#include "lk_internal.h"

namespace lk {

namespace {

constexpr int kMaxStrUnitLength = 20;   // ...engine.rs:98
constexpr int kMaxRepeatLength = 100;   // :99
constexpr uint8_t kMinUsableQ = 6;      // quality_utils.rs:23

__device__ __forceinline__ bool same_bytes(const uint8_t *a, const uint8_t *b, int n) {
    for (int i = 0; i < n; i++)
        if (a[i] != b[i]) return false;
    return true;
}

// find_tandem_repeat_units(...).1 (...engine.rs:528-611), by run lengths.  Consecutive copies of a unit of length s are
// a run of positions with b[p] == b[p - s]: the reference's window-by-window counts are
//   copies ending at `offset`        = 1 + RB / s,  RB = run of equalities ending at offset
//   copies starting at `offset + 1`  = 1 + RF / s,  RF = run starting at offset + 1 + s
//   copies of a forward unit (length L) that precede it = RB' / L,  RB' = run ending at offset + L
// so each (offset, s) costs one short while loop instead of window compares behind a call -- which matters because the
// lanes of a warp sit on different offsets: with the call, some lane was inside it at almost every s and the warp ran
// at 8.7 active lanes of 32.
__device__ __forceinline__ int run_back(const uint8_t *b, int q, int s) {  // run of b[p] == b[p - s] ending at q
    int p = q;
    while (p >= s && b[p] == b[p - s]) p--;
    return q - p;
}
__device__ __noinline__ int tandem_repeat_length(const uint8_t *bases, int len, int offset) {
    int max_bw = 0, bw_len = 1;
    for (int str = 1; str <= kMaxStrUnitLength; str++) {
        if (offset + 1 < str) break;
        const int run = run_back(bases, offset, str);
        max_bw = run >= str ? 1 + run / str : 1;
        if (max_bw > 1) { bw_len = str; break; }
    }
    int max_rl = max_bw;
    if (offset < len - 1) {
        int fw_len = 1, max_fw = 0;
        for (int str = 1; str <= kMaxStrUnitLength; str++) {
            if (offset + str + 1 > len) break;
            const int p0 = offset + 1 + str;
            int p = p0;
            while (p < len && bases[p] == bases[p - str]) p++;
            max_fw = (p - p0) >= str ? 1 + (p - p0) / str : 1;
            if (max_fw > 1) { fw_len = str; break; }
        }
        if (fw_len == bw_len && same_bytes(bases + offset + 1, bases + offset + 1 - bw_len, fw_len)) {
            max_rl = max_bw + max_fw;
        } else {
            int before = 0;  // copies of the forward unit among the bases up to `offset`
            if (offset + 1 >= fw_len) {
                const int run = run_back(bases, offset + fw_len, fw_len);
                before = run >= fw_len ? run / fw_len : 0;
            }
            max_rl = max_fw + before;
        }
    }
    return max_rl < kMaxRepeatLength ? max_rl : kMaxRepeatLength;
}

// The same search for a read staged in shared memory (slab laid out as kEnginePad guard bytes, the read, guard bytes;
// `b` points at the read's first byte, 4-byte aligned).  A unit length s can only repeat at `offset` if the single byte
// pair b[offset] == b[offset - s] (backward) / b[offset + 1] == b[offset + 1 + s] (forward) matches, so all 20 pairs of
// a direction are compared at once -- six aligned LDS.32, five funnel shifts, a SWAR zero-byte test per word (exact for
// any byte value) and a multiply that gathers the four flags of a word -- and only the lengths whose bit survives (5 of
// 20 on random sequence) enter the run-length loop.  Guard bytes are never trusted: the validity masks cut the lengths
// that would reach outside the read.  Byte-identical to tandem_repeat_length (tests/test_gpu_engine.py).
constexpr int kEnginePad = 24;
// floor(run / s) for 0 <= run < 2^15, 1 <= s <= 20 without the ~25-instruction integer division: (run + 0.5) / s lies
// at least 0.5/20 away from every integer, far more than the error of the approximate float division
__device__ __forceinline__ int div_small(int run, int s) { return __float2int_rd(__fdividef((float)run + 0.5f, (float)s)); }
__device__ __forceinline__ uint32_t zero_byte_flags(uint32_t x) {  // bit 8b+7 set iff byte b of x is 0
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}
template <bool BACKWARD>
__device__ __forceinline__ uint32_t window_eq_mask(const uint8_t *b, int first, uint32_t c) {
    // 20 bytes b[first .. first+19] against byte c.  BACKWARD: byte t -> bit 19 - t (s = 20 - t); else byte t -> bit t (s = t + 1)
    const int p0 = first + kEnginePad;  // >= 4: index from the slab's (aligned) start
    const uint32_t *w = reinterpret_cast<const uint32_t *>(b - kEnginePad) + (p0 >> 2);
    const uint32_t sh = (uint32_t)(p0 & 3) * 8u, splat = c * 0x01010101u;
    uint32_t lo = w[0], mask = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const uint32_t hi = w[k + 1];
        const uint32_t z = zero_byte_flags(__funnelshift_r(lo, hi, sh) ^ splat);
        lo = hi;
        if (BACKWARD) mask |= (__umulhi(z, 0x10080402u) & 0xfu) << (16 - 4 * k);  // flags -> bits 3,2,1,0 of the high word
        else          mask |= (__umulhi(z, 0x02040810u) & 0xfu) << (4 * k);       // flags -> bits 0,1,2,3
    }
    return mask;
}
__device__ __forceinline__ int tandem_repeat_length_staged(const uint8_t *b, int len, int offset) {
    int bw_len = 1, bw_run = 0;  // found unit length and its run of equalities (0: none, one copy)
    uint32_t m = window_eq_mask<true>(b, offset - 20, b[offset]) & (offset >= 20 ? 0xfffffu : (1u << offset) - 1u);
    // a unit of length >= 2 needs the neighbouring byte pair to match as well: second screen, same window one byte down
    if (m & ~1u) m &= window_eq_mask<true>(b, offset - 21, b[offset - 1]) & (offset >= 21 ? 0xfffffu : (1u << (offset - 1)) - 1u) | 1u;
    while (m) {
        const int s = __ffs(m);
        m &= m - 1;
        const int run = 1 + run_back(b, offset - 1, s);
        if (run >= s) { bw_run = run; bw_len = s; break; }
    }
    const int max_bw = 1 + div_small(bw_run, bw_len);  // outside the divergent loop, every lane at once
    int max_rl = max_bw;
    if (offset < len - 1) {
        int fw_len = 1, fw_run = 0;
        const int room = len - 2 - offset;  // s <= room keeps b[offset + 1 + s] inside the read
        m = window_eq_mask<false>(b, offset + 2, b[offset + 1]) & (room >= 20 ? 0xfffffu : (1u << room) - 1u);
        if (m & ~1u) m &= window_eq_mask<false>(b, offset + 3, b[offset + 2]) & (room >= 21 ? 0xfffffu : (1u << (room - 1)) - 1u) | 1u;
        while (m) {
            const int s = __ffs(m);
            m &= m - 1;
            const int p0 = offset + 1 + s;
            int p = p0 + 1;
            while (p < len && b[p] == b[p - s]) p++;
            if (p - p0 >= s) { fw_run = p - p0; fw_len = s; break; }
        }
        const int max_fw = 1 + div_small(fw_run, fw_len);
        if (fw_len == bw_len && same_bytes(b + offset + 1, b + offset + 1 - bw_len, fw_len)) {
            max_rl = max_bw + max_fw;
        } else {
            int before = 0;
            if (offset + 1 >= fw_len) {
                const int run = run_back(b, offset + fw_len, fw_len);
                before = run >= fw_len ? div_small(run, fw_len) : 0;
            }
            max_rl = max_fw + before;
        }
    }
    return max_rl < kMaxRepeatLength ? max_rl : kMaxRepeatLength;
}

// ...engine.rs:23-39, (mean, variance) for baseQ 1..40
__constant__ double kDynMean[40] = {
    5.996842844, 5.870018422, 5.401558531, 4.818940919, 4.218758304, 3.646319832, 3.122346753, 2.654731979, 2.244479156, 1.88893867,
    1.583645342, 1.3233807, 1.102785365, 0.916703025, 0.760361881, 0.629457387, 0.520175654, 0.42918208, 0.353590663, 0.290923699,
    0.23906788, 0.196230431, 0.160897421, 0.131795374, 0.1078567, 0.088189063, 0.072048567, 0.058816518, 0.047979438, 0.039111985,
    0.031862437, 0.025940415, 0.021106532, 0.017163711, 0.013949904, 0.011332027, 0.009200898, 0.007467036, 0.006057179, 0.004911394 };
__constant__ double kDynVar[40] = {
    0.196616587, 1.388545569, 5.641990128, 10.33176216, 14.25799688, 17.02880749, 18.64537883, 19.27521677, 19.13584613, 18.43922003,
    17.36842261, 16.07088712, 14.65952563, 13.21718577, 11.80207947, 10.45304833, 9.194183767, 8.038657241, 6.991779595, 6.053379213,
    5.219610436, 4.484302033, 3.839943445, 3.27839108, 2.791361596, 2.370765375, 2.008921719, 1.698687797, 1.433525748, 1.207526336,
    1.015402928, 0.852465956, 0.714585285, 0.598145851, 0.500000349, 0.41742159, 0.348056286, 0.289881373, 0.241163527, 0.200422214 };

}  // namespace

// one warp per read, lanes over its bases; in place on the staged copies.  The read's bases are staged in shared memory
// (the repeat search reads every base ~40 times); reads longer than the slab are scanned in global memory.
// (A run-length formulation -- one ballot per 32 positions and unit length gives every repeat count of the read -- was
// written and is byte-identical, but at 20 unit lengths x 2 scans it issues ~3x more warp instructions than the
// per-base search, which stops at the first mismatching byte: 3.6 ms against 2.0 ms per 2 000 regions.  Dropped.)
constexpr int kEngineSlab = 1024;
__global__ void __launch_bounds__(128) engine_modify_reads_kernel(EngineDevArgs E, int64_t n_reads) {
    __shared__ __align__(16) uint8_t slab[4][kEnginePad + kEngineSlab + kEnginePad + 8];
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    uint8_t *mine = slab[threadIdx.x >> 5];
    for (int64_t r = warp; r < n_reads; r += n_warps) {
        const int64_t o = E.read_off[r];
        const int len = (int)(E.read_off[r + 1] - o);
        const uint8_t *bases = E.read_bases + o;
        const uint8_t mapq = E.cap_to_mapq ? E.read_mapq[r] : 255;
        const bool staged = E.use_pcr && len <= kEngineSlab;
        if (staged) {
            __syncwarp();
            for (int j = lane; j < len; j += 32) mine[kEnginePad + j] = bases[j];
            __syncwarp();
        }
        // never assigned from a global pointer, so the search compiles to LDS with 32-bit addresses (a pointer that may be
        // either costs a generic load and 64-bit address arithmetic per byte)
        const uint8_t *staged_bases = mine + kEnginePad;
        for (int j = lane; j < len; j += 32) {
            uint8_t q = E.read_qual[o + j], qi = E.read_ins[o + j], qd = E.read_del[o + j];
            if (E.use_pcr && j < len - 1) {  // apply_pcr_error_model writes i-1 for i in 1..len: the last base is never touched
                const uint8_t c = E.pcr_cache[staged ? tandem_repeat_length_staged(staged_bases, len, j) : tandem_repeat_length(bases, len, j)];
                qi = qi < c ? qi : c;
                qd = qd < c ? qd : c;
            }
            if (E.cap_to_mapq) q = q < mapq ? q : mapq;     // cap_minimum_read_qualities :428-458
            if (q < E.bq_threshold) q = kMinUsableQ;
            if (qi < kMinUsableQ) qi = kMinUsableQ;
            if (qd < kMinUsableQ) qd = kMinUsableQ;
            E.read_qual[o + j] = q; E.read_ins[o + j] = qi; E.read_del[o + j] = qd;
        }
    }
}

// one warp per (region, sample) block: the reference normalises and filters each sample's [A x R_s] matrix on its own
// (pair_hmm_likelihood_calculation_engine.rs:201-239, allele_likelihoods.rs:378-444,925-1018)
__global__ void engine_epilogue_kernel(EngineDevArgs E, int64_t n_blocks) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const double ninf = -INFINITY;
    for (int64_t b = warp; b < n_blocks; b += n_warps) {
        const int64_t g = E.block_region ? (int64_t)E.block_region[b] : b;
        const int64_t r0 = E.block_read_start[b];
        const int R = (int)(E.block_read_start[b + 1] - r0);
        const int A = (int)(E.region_hap_start[g + 1] - E.region_hap_start[g]);
        if (R == 0) continue;
        if (A == 0) {  // no alleles: the maximum over none is -inf, below any threshold -> every read is filtered
            if (E.keep) for (int r = lane; r < R; r += 32) E.keep[r0 + r] = 0;
            continue;
        }
        const int64_t base = E.read_out_base[r0];  // pairs before this region (chunk-local)
        const double *src = E.lk + base;            // [r*A + a]
        double *v = E.values + base;                // [a*R + r]
        const int ref_idx = E.ref_allele_index ? E.ref_allele_index[g] : 0;
        const bool normalise = !(E.cap_diff == ninf) && A >= 2;
        int kept_before = 0;
        for (int rb = 0; rb < R; rb += 32) {
            const int r = rb + lane;
            bool keep = false;
            double worst = ninf;
            if (r < R) {
                if (normalise) {  // (A >= 2, so an allele other than the reference exists)
                    // search_best_allele(.., can_be_reference = symmetric, None): only the best likelihood is needed
                    int first = (E.symmetric || ref_idx != 0) ? 0 : 1;
                    double best = src[(int64_t)r * A + first];
                    for (int a = first + 1; a < A; a++) {
                        if (!E.symmetric && a == ref_idx) continue;
                        const double c = src[(int64_t)r * A + a];
                        if (c > best) best = c;
                    }
                    worst = __dadd_rn(best, E.cap_diff);
                }
                // maximum over ALL alleles after normalisation (filter_poorly_modeled_evidence)
                double mx = ninf;
                for (int a = 0; a < A; a++) {
                    double c = src[(int64_t)r * A + a];
                    if (normalise && c < worst) c = worst;
                    if (c > mx) mx = c;
                }
                // threshold from the ORIGINAL qualities / length (the "HMMQuals" key quirk)
                const int64_t o = E.read_off[r0 + r];
                const int len = (int)(E.read_off[r0 + r + 1] - o);
                double e = ceil(__dmul_rn((double)len, E.expected_error_rate));
                double thr;
                if (E.dynamic_disq) {
                    double mean = 0.0, var = 0.0;
                    for (int i = 0; i < len; i++) {
                        const int bq = E.qual_orig[o + i];
                        const int entry = bq <= 1 ? 0 : (bq < 40 ? bq : 40) - 1;
                        mean = __dadd_rn(mean, kDynMean[entry]);
                        var = __dadd_rn(var, kDynVar[entry]);
                    }
                    const double dyn = __dmul_rn(__dadd_rn(mean, __dmul_rn(E.disq_scale, sqrt(var))), -0.1);
                    const double stat = __dmul_rn(e, -4.0);
                    thr = dyn < stat ? dyn : stat;
                } else {
                    thr = __dmul_rn(e < 2.0 ? e : 2.0, -4.0);
                }
                keep = !(mx < thr);
                if (E.keep) E.keep[r0 + r] = keep ? 1 : 0;
            }
            const unsigned kb = __ballot_sync(0xffffffffu, keep);
            if (keep) {  // compacted column of this read
                const int col = kept_before + __popc(kb & ((1u << lane) - 1u));
                for (int a = 0; a < A; a++) {
                    double c = src[(int64_t)r * A + a];
                    if (normalise && c < worst) c = worst;
                    v[(int64_t)a * R + col] = c;
                }
            }
            kept_before += __popc(kb);
        }
        // tail of every row: NaN (allele_likelihoods.rs:996-1009)
        const int tail = R - kept_before;
        for (int t = lane; t < tail * A; t += 32) {
            const int a = t / tail, c = kept_before + t % tail;
            v[(int64_t)a * R + c] = __longlong_as_double(0x7ff8000000000000LL);
        }
    }
}

cudaError_t launch_engine_modify(const EngineDevArgs &E, int64_t n_reads, cudaStream_t st) {
    if (n_reads == 0) return cudaSuccess;
    const int threads = 128;
    int64_t blocks = (n_reads * 32 + threads - 1) / threads;
    if (blocks > 148 * 32) blocks = 148 * 32;
    engine_modify_reads_kernel<<<(unsigned)blocks, threads, 0, st>>>(E, n_reads);
    return cudaGetLastError();
}

cudaError_t launch_engine_epilogue(const EngineDevArgs &E, int64_t n_blocks, cudaStream_t st) {
    if (n_blocks == 0) return cudaSuccess;
    const int threads = 128;
    int64_t blocks = (n_blocks * 32 + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    engine_epilogue_kernel<<<(unsigned)blocks, threads, 0, st>>>(E, n_blocks);
    return cudaGetLastError();
}

// Best haplotype of every read with the reference's tie-breaking, on likelihood matrices that are already in HBM
// (realign_reads_to_their_best_haplotype, assembly_based_caller_utils.rs:208-246): one thread per read.
__global__ void best_alleles_kernel(const double *__restrict__ values, const BestAlleleBlock *__restrict__ blocks, int64_t n_blocks,
                                    const int32_t *__restrict__ priorities, double informative_threshold,
                                    int32_t *__restrict__ best_allele, uint8_t *__restrict__ informative) {
    for (int64_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const BestAlleleBlock B = blocks[b];
        for (int r = threadIdx.x; r < B.n_reads; r += blockDim.x) {
            int32_t best = -1;
            uint8_t info = 0;
            if (r < B.n_evidence)
                best_allele_of_read(values + B.values_off, B.n_alleles, B.n_reads, r, priorities + B.prio_off, informative_threshold,
                                    &best, nullptr, &info);
            best_allele[B.first_read + r] = best;
            informative[B.first_read + r] = info;
        }
    }
}

cudaError_t launch_best_alleles(const double *values, const BestAlleleBlock *blocks, int64_t n_blocks, const int32_t *priorities,
                                double informative_threshold, int32_t *best_allele, uint8_t *informative, cudaStream_t st) {
    if (n_blocks == 0) return cudaSuccess;
    const int64_t grid = n_blocks < 148 * 16 ? n_blocks : 148 * 16;
    best_alleles_kernel<<<(unsigned)grid, 64, 0, st>>>(values, blocks, n_blocks, priorities, informative_threshold, best_allele, informative);
    return cudaGetLastError();
}

}  // namespace lk
