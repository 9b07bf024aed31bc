// pairhmm_kernels.cu — sm_100a kernels of the PairHMM path and their launchers.
//
//   encode_haps_kernel   bytes -> prior-table slice index per column (+ zero guard bands), exotic-byte flags
//   scan_reads_kernel    per-read domain flags for the fp32 hat form
//   pairhmm_fast_kernel  fp32 wavefront, one warp per task (reads of one region x all its haplotypes)
//   pairhmm_plain_kernel f64 pass for flagged pairs (persistent warps pulling from the rescue list)
//
// Reference semantics: src/pair_hmm/pair_hmm.rs:345-375 (driver), :503-615 (recurrence).
#include <cuda_runtime.h>

#include "lk_internal.h"
#include "pairhmm_core.cuh"

namespace lk {

struct DevWarp {
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    template <class T> __device__ __forceinline__ T shfl_up1(T v) const { return __shfl_up_sync(0xffffffffu, v, 1); }
    template <class T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    template <class T> __device__ __forceinline__ T shfl_xor(T v, int m) const { return __shfl_xor_sync(0xffffffffu, v, m); }
    __device__ __forceinline__ unsigned reduce_or(unsigned v) const { return __reduce_or_sync(0xffffffffu, v); }
    __device__ __forceinline__ unsigned reduce_max(unsigned v) const { return __reduce_max_sync(0xffffffffu, v); }
    __device__ __forceinline__ int reduce_min_i(int v) const { return __reduce_min_sync(0xffffffffu, v); }
    template <class T> __device__ __forceinline__ T ldg(const T *p) const { return __ldg(p); }
    __device__ __forceinline__ int atomic_add(int32_t *p, int v) const { return atomicAdd(p, v); }
    __device__ __forceinline__ void sync() const { __syncwarp(); __threadfence_block(); }
    __device__ __forceinline__ void lds4(float *dst, const float *src) const {
        const float4 v = *reinterpret_cast<const float4 *>(src);
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __device__ __forceinline__ void prefetch(const void *p) const { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
};

// ---------------------------------------------------------------------------------------------
__global__ void encode_haps_kernel(const uint8_t *__restrict__ hap_bases, const int64_t *__restrict__ hap_off,
                                   const int64_t *__restrict__ henc_off, int64_t n_haps,
                                   uint8_t *__restrict__ hidx, uint8_t *__restrict__ hap_flags) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t h = warp; h < n_haps; h += n_warps) {
        const int64_t ho = hap_off[h];
        const int H = (int)(hap_off[h + 1] - ho);
        uint8_t *dsti = hidx + henc_off[h];  // column 1
        bool exotic = false;
        for (int j = lane - HENC_PADL; j < H + HENC_PADR; j += 32) {
            uint8_t sl = 0;
            if (j >= 0 && j < H) { (void)encode_base(hap_bases[ho + j], exotic); sl = base_slice(hap_bases[ho + j]); }
            dsti[j] = sl;
        }
        const unsigned any = __ballot_sync(0xffffffffu, exotic);
        if (lane == 0) hap_flags[h] = any ? SEQ_EXOTIC : 0;
    }
}

__global__ void scan_reads_kernel(const uint8_t *__restrict__ read_bases, const uint8_t *__restrict__ read_qual,
                                  const uint8_t *__restrict__ read_ins, const uint8_t *__restrict__ read_del,
                                  const int64_t *__restrict__ read_off,
                                  int64_t n_reads, uint8_t *__restrict__ read_flags) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < n_reads; r += n_warps) {
        const int64_t ro = read_off[r];
        const int R = (int)(read_off[r + 1] - ro);
        bool exotic = false, qzero = false;
        int imin = 255, imax = 0;
        for (int i = lane; i < R; i += 32) {
            (void)encode_base(read_bases[ro + i], exotic);
            qzero |= (read_qual[ro + i] == 0);  // match prior would be exactly 0: hat form divides by it
            const int qi = read_ins[ro + i];
            qzero |= (qi < 4) || (read_del[ro + i] < 4);  // match->match may be exactly 0: scaled form divides by it
            imin = min(imin, qi);
            imax = max(imax, qi);
        }
        imin = __reduce_min_sync(0xffffffffu, imin);
        imax = __reduce_max_sync(0xffffffffu, imax);
        const unsigned ex = __ballot_sync(0xffffffffu, exotic);
        const unsigned qz = __ballot_sync(0xffffffffu, qzero);
        if (lane == 0) {
            uint8_t f = 0;
            if (ex) f |= SEQ_EXOTIC;
            if (qz || (imax - imin) > FAST_MAX_INS_RANGE) f |= SEQ_QUAL_DOMAIN;
            read_flags[r] = f;
        }
    }
}

// every pair goes to the f64 pass (lk_set_force_f64)
__global__ void list_all_pairs_kernel(PairHmmArgs A, int64_t n_reads) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const int32_t region = A.read_region[r];
    const int64_t h0 = A.region_hap_start[region], h1 = A.region_hap_start[region + 1];
    const int64_t base = A.read_out_base[r];
    for (int64_t h = h0; h < h1; ++h) {
        A.rescue_pairs[base + (h - h0)].read = (int32_t)r;
        A.rescue_pairs[base + (h - h0)].hap = (int32_t)h;
    }
}

// ---------------------------------------------------------------------------------------------
template <int K, int GM, int MINB, bool SPLIT = false>
__global__ void __launch_bounds__(LK_FAST_THREADS, MINB) pairhmm_fast_kernel(const PairHmmArgs A) {
    extern __shared__ float4 lk_fast_smem[];  // per-warp prior tables
    const int task = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (task >= A.n_tasks) return;
    DevWarp w;
    float *tab = reinterpret_cast<float *>(lk_fast_smem) + (threadIdx.x >> 5) * fast_table_floats(K);
    pairhmm_fast_task<DevWarp, K, GM, SPLIT>(w, A, task, tab);
}

template <int K, bool Q10>
__global__ void __launch_bounds__(LK_BAND_THREADS, LK_BAND_MINB) pairhmm_band_kernel(const PairHmmArgs A) {
    extern __shared__ float4 lk_band_smem[];  // per warp: prior table, then the ring of staged boundary elements
    float *tab = reinterpret_cast<float *>(lk_band_smem) + (threadIdx.x >> 5) * band_smem_floats(K);
    Bnd4 *ring = reinterpret_cast<Bnd4 *>(tab + fast_table_floats(K));
    DevWarp w;
    const int warp_in_grid = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    Bnd4 *scratch = (Bnd4 *)A.band_scratch + (int64_t)warp_in_grid * A.band_scratch_stride;
    for (;;) {
        int idx = 0;
        if (w.lane() == 0) idx = atomicAdd(A.band_cursor, 1);
        idx = w.shfl(idx, 0);
        if (idx >= A.n_band) break;
        const RescuePair p = A.band_pairs[idx];
        pairhmm_band_pair<DevWarp, K, Q10>(w, A, p.read, (int64_t)p.hap, scratch, tab, ring);
        __syncwarp();
    }
}

template <int K>
__global__ void __launch_bounds__(LK_PLAIN_THREADS) pairhmm_plain_kernel(const PairHmmArgs A) {
    DevWarp w;
    const int warp_in_grid = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    double *scratch = A.scratch + (int64_t)warp_in_grid * A.scratch_stride;
    const int n = min(*A.rescue_count, A.rescue_capacity);
    for (;;) {
        int idx = 0;
        if (w.lane() == 0) idx = atomicAdd(A.rescue_cursor, 1);
        idx = w.shfl(idx, 0);
        if (idx >= n) break;
        const RescuePair p = A.rescue_pairs[idx];
        pairhmm_plain_pair<DevWarp, K>(w, A, p.read, (int64_t)p.hap, scratch);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// launchers (host)
// ---------------------------------------------------------------------------------------------
cudaError_t launch_encode_haps(const uint8_t *hap_bases, const int64_t *hap_off, const int64_t *henc_off,
                               int64_t n_haps, uint8_t *hidx, uint8_t *hap_flags, cudaStream_t st) {
    if (n_haps == 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks = (n_haps * 32 + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    encode_haps_kernel<<<(unsigned)blocks, threads, 0, st>>>(hap_bases, hap_off, henc_off, n_haps, hidx, hap_flags);
    return cudaGetLastError();
}

cudaError_t launch_scan_reads(const uint8_t *read_bases, const uint8_t *read_qual, const uint8_t *read_ins,
                              const uint8_t *read_del, const int64_t *read_off, int64_t n_reads, uint8_t *read_flags, cudaStream_t st) {
    if (n_reads == 0) return cudaSuccess;
    const int threads = 256;
    int64_t blocks = (n_reads * 32 + threads - 1) / threads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    scan_reads_kernel<<<(unsigned)blocks, threads, 0, st>>>(read_bases, read_qual, read_ins, read_del, read_off, n_reads, read_flags);
    return cudaGetLastError();
}

cudaError_t launch_list_all_pairs(const PairHmmArgs &A, int64_t n_reads, cudaStream_t st) {
    if (n_reads == 0) return cudaSuccess;
    list_all_pairs_kernel<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(A, n_reads);
    return cudaGetLastError();
}

template <int K, int GM, int MINB, bool SPLIT = false>
static cudaError_t launch_fast_kg(const PairHmmArgs &A, cudaStream_t st) {
    const int warps_per_block = LK_FAST_THREADS / 32;
    const unsigned blocks = (unsigned)((A.n_tasks + warps_per_block - 1) / warps_per_block);
    const int smem = warps_per_block * FastTable<K>::floats * (int)sizeof(float);
    // the opt-in is per function AND per device: one flag per device ordinal (idempotent, so a race is harmless)
    static bool configured_on[64] = { false };
    int dev = 0;
    cudaGetDevice(&dev);
    bool &configured = configured_on[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(pairhmm_fast_kernel<K, GM, MINB, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    pairhmm_fast_kernel<K, GM, MINB, SPLIT><<<blocks, LK_FAST_THREADS, smem, st>>>(A);
    return cudaGetLastError();
}

template <int K, int MINB>
static cudaError_t launch_fast_k(const PairHmmArgs &A, bool vargcp, cudaStream_t st) {
    if (vargcp) return launch_fast_kg<K, GM_VAR, MINB>(A, st);
    if (A.gcp_constant == 10 && A.cf_uniform == 0.1f) {
        // tasks cut along the haplotype axis (small launches, host_prep.cpp) exist for the engine's default gcp only; with
        // any other gcp the whole-region kernel runs the same task list, each task then sweeping (and rewriting) all
        // haplotypes of its region: redundant, never wrong
        if (A.task_haps) return launch_fast_kg<K, GM_Q10, MINB, true>(A, st);
        return launch_fast_kg<K, GM_Q10, MINB>(A, st);
    }
    return launch_fast_kg<K, GM_UNIFORM, MINB>(A, st);
}

cudaError_t launch_fast(int K, const PairHmmArgs &A, bool vargcp, cudaStream_t st) {
    if (A.n_tasks == 0) return cudaSuccess;
    switch (K) {
#define LK_CASE(KV, MINB) case KV: return launch_fast_k<KV, MINB>(A, vargcp, st);
        LK_FAST_K_TABLE(LK_CASE)
#undef LK_CASE
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_band(const PairHmmArgs &A, int n_blocks, cudaStream_t st) {
    if (A.n_band == 0) return cudaSuccess;
    const int smem = (LK_BAND_THREADS / 32) * band_smem_floats(LK_BAND_K) * (int)sizeof(float);
    static bool configured_on[64] = { false };  // per device ordinal: the attribute is per function and device
    int dev = 0;
    cudaGetDevice(&dev);
    bool &configured = configured_on[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(pairhmm_band_kernel<LK_BAND_K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(pairhmm_band_kernel<LK_BAND_K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    if (A.gcp_constant == 10 && A.cf_uniform == 0.1f) pairhmm_band_kernel<LK_BAND_K, true><<<n_blocks, LK_BAND_THREADS, smem, st>>>(A);
    else pairhmm_band_kernel<LK_BAND_K, false><<<n_blocks, LK_BAND_THREADS, smem, st>>>(A);
    return cudaGetLastError();
}

cudaError_t launch_plain(const PairHmmArgs &A, int n_blocks, cudaStream_t st) {
    pairhmm_plain_kernel<LK_PLAIN_K><<<n_blocks, LK_PLAIN_THREADS, 0, st>>>(A);
    return cudaGetLastError();
}

}  // namespace lk
