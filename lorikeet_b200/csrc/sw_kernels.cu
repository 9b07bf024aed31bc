// sw_kernels.cu — sm_100a Smith-Waterman kernel (fill + in-kernel traceback) and its launcher.
// Persistent warps pull pair ids from a cursor; each warp owns a slice of scratch for the 4-bit backtrack.
#include <cuda_runtime.h>

#include <algorithm>

#include "lk_internal.h"
#include "sw16_core.cuh"

namespace lk {

struct SwDevWarp {
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    template <class T> __device__ __forceinline__ T shfl_up1(T v) const { return __shfl_up_sync(0xffffffffu, v, 1); }
    template <class T> __device__ __forceinline__ T shfl(T v, int src) const { return __shfl_sync(0xffffffffu, v, src); }
    template <class T> __device__ __forceinline__ T shfl_xor(T v, int m) const { return __shfl_xor_sync(0xffffffffu, v, m); }
    __device__ __forceinline__ unsigned ballot(bool p) const { return __ballot_sync(0xffffffffu, p); }
    __device__ __forceinline__ int ffs(unsigned v) const { return __ffs((int)v); }
    template <class T> __device__ __forceinline__ T ldg(const T *p) const { return __ldg(p); }
    __device__ __forceinline__ void sync() const { __threadfence_block(); __syncwarp(); }
    __device__ __forceinline__ void store4(uint32_t *p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) const {
        *reinterpret_cast<uint4 *>(p) = make_uint4(a, b, c, d);
    }
    __device__ __forceinline__ void store2(uint32_t *p, uint32_t a, uint32_t b) const { *reinterpret_cast<uint2 *>(p) = make_uint2(a, b); }
    __device__ __forceinline__ void load2(const uint32_t *p, uint32_t &a, uint32_t &b) const {
        const uint2 v = *reinterpret_cast<const uint2 *>(p);
        a = v.x; b = v.y;
    }
    __device__ __forceinline__ void lds4i(int32_t *dst, const int32_t *src) const {
        const int4 v = *reinterpret_cast<const int4 *>(src);
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __device__ __forceinline__ void lds4u(uint32_t *dst, const uint32_t *src) const {
        const uint4 v = *reinterpret_cast<const uint4 *>(src);
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
    __device__ __forceinline__ void load4(const uint32_t *p, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) const {
        const uint4 v = *reinterpret_cast<const uint4 *>(p);
        a = v.x; b = v.y; c = v.z; d = v.w;
    }
};

template <int K, bool CLAMP>
__global__ void __launch_bounds__(LK_SW_THREADS) sw_kernel(const SwArgs A) {
    extern __shared__ int4 lk_sw_smem[];  // per-warp substitution-weight tables
    int32_t *wtab = reinterpret_cast<int32_t *>(lk_sw_smem) + (threadIdx.x >> 5) * SwTable<K>::words;
    SwDevWarp w;
    const int warp_in_grid = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    uint32_t *scratch = A.scratch + (int64_t)warp_in_grid * A.scratch_words;
    for (;;) {
        int idx = 0;
        if (w.lane() == 0) idx = atomicAdd(A.cursor, 1);
        idx = w.shfl(idx, 0);
        if (idx >= A.n_order) break;
        sw_align_pair<SwDevWarp, K, CLAMP>(w, A, A.order[idx], scratch, wtab);
        __syncwarp();
    }
}

// Two pairs per warp (sw16_core.cuh) wherever two consecutive pairs of the order qualify, one at a time (sw_core.cuh)
// otherwise.
template <int K> struct SwDuoSmem { static constexpr int words = Sw16Table<K>::words > SwTable<K>::words ? Sw16Table<K>::words : SwTable<K>::words; };

// warps per block x blocks per SM of the packed kernel: 4 x 4 = 16 warps per SM (128 registers, 12.8 KB of shared memory per
// warp: the one-pair table of the fall-back path is the larger of the two layouts).  With sw16_core.cuh's LK_SW16_COMPACT=0
// tables it has to be 4 x 3.
#ifndef LK_SW16_BLOCK_WARPS
#define LK_SW16_BLOCK_WARPS 4
#define LK_SW16_BLOCKS_PER_SM (LK_SW16_COMPACT ? 4 : 3)
#endif

template <int K, bool SAME_N>
__global__ void __launch_bounds__(LK_SW16_BLOCK_WARPS * 32, LK_SW16_BLOCKS_PER_SM) sw_duo_kernel(const SwArgs A) {
    extern __shared__ int4 lk_sw_smem[];
    int32_t *wtab = reinterpret_cast<int32_t *>(lk_sw_smem) + (threadIdx.x >> 5) * SwDuoSmem<K>::words;
    SwDevWarp w;
    const int warp_in_grid = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    uint32_t *scratch = A.scratch + (int64_t)warp_in_grid * A.scratch_words;
    for (;;) {
        int idx = 0;
        if (w.lane() == 0) idx = atomicAdd(A.cursor, 2);
        idx = w.shfl(idx, 0);
        if (idx >= A.n_order) break;
        const int pa = A.order[idx], pb = (idx + 1 < A.n_order) ? A.order[idx + 1] : -1;
        LK_SW16_TICK_INIT;
        const bool done_a = sw_prologue(w, A, pa), done_b = (pb < 0) || sw_prologue(w, A, pb);
        LK_SW16_TICK(5);
        if (!done_a && !done_b) {
            const int64_t na = A.ref_off[pa + 1] - A.ref_off[pa], nb = A.ref_off[pb + 1] - A.ref_off[pb];
            if (na <= 32 * K && nb <= 32 * K && (!SAME_N || na == nb) && sw16_align_duo<SwDevWarp, K, SAME_N>(w, A, pa, pb, scratch, wtab)) {
                if (w.lane() == 0) atomicAdd(A.packed_count, 2);
                __syncwarp();
                continue;
            }
        }
        if (!done_a) { sw_dp_pair<SwDevWarp, K, false>(w, A, pa, scratch, wtab); __syncwarp(); }
        if (!done_b) { sw_dp_pair<SwDevWarp, K, false>(w, A, pb, scratch, wtab); __syncwarp(); }
    }
}

template <int K, bool SAME_N>
static cudaError_t launch_sw_duo_k(const SwArgs &A, int n_blocks, cudaStream_t st) {
    const int smem = LK_SW16_BLOCK_WARPS * SwDuoSmem<K>::words * (int)sizeof(int32_t);
    static bool configured_on[64] = { false };
    int dev = 0;
    cudaGetDevice(&dev);
    bool &configured = configured_on[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(sw_duo_kernel<K, SAME_N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    sw_duo_kernel<K, SAME_N><<<n_blocks, LK_SW16_BLOCK_WARPS * 32, smem, st>>>(A);
    return cudaGetLastError();
}

// K classes the packed kernel is built for (references of 129..544 rows); the launcher falls back to launch_sw otherwise
bool sw_duo_has_k(int K) { return K == 8 || K == 12 || K == 16 || K == 17; }
// n_warps: per-warp scratch slices available; the grid is the device's fill of the packed kernel, or fewer
// same_n: every two consecutive pairs of the order have references of the same length (the host arranges the order so)
cudaError_t launch_sw_duo(int K, bool same_n, const SwArgs &A, int n_sm, int64_t n_warps, cudaStream_t st) {
    if (A.n_order == 0) return cudaSuccess;
    const int n_blocks = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)n_sm * LK_SW16_BLOCKS_PER_SM, n_warps / LK_SW16_BLOCK_WARPS));
    if (n_warps < LK_SW16_BLOCK_WARPS) return cudaErrorInvalidValue;
    if (K == 8) return same_n ? launch_sw_duo_k<8, true>(A, n_blocks, st) : launch_sw_duo_k<8, false>(A, n_blocks, st);
    if (K == 12) return same_n ? launch_sw_duo_k<12, true>(A, n_blocks, st) : launch_sw_duo_k<12, false>(A, n_blocks, st);
    if (K == 16) return same_n ? launch_sw_duo_k<16, true>(A, n_blocks, st) : launch_sw_duo_k<16, false>(A, n_blocks, st);
    if (K == 17) return same_n ? launch_sw_duo_k<17, true>(A, n_blocks, st) : launch_sw_duo_k<17, false>(A, n_blocks, st);
    return cudaErrorInvalidValue;
}

template <int K, bool CLAMP>
static cudaError_t launch_sw_k(const SwArgs &A, int n_blocks, cudaStream_t st) {
    const int smem = (LK_SW_THREADS / 32) * SwTable<K>::words * (int)sizeof(int32_t);
    static bool configured_on[64] = { false };  // per device ordinal: the attribute is per function and device
    int dev = 0;
    cudaGetDevice(&dev);
    bool &configured = configured_on[dev & 63];
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(sw_kernel<K, CLAMP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    sw_kernel<K, CLAMP><<<n_blocks, LK_SW_THREADS, smem, st>>>(A);
    return cudaGetLastError();
}

// clamp: false when the host has proved the -1e8 clamp inactive for every pair of the launch (sw_clamp_inactive)
cudaError_t launch_sw(int K, bool clamp, const SwArgs &A, int n_blocks, cudaStream_t st) {
    if (A.n_order == 0) return cudaSuccess;
    switch (K) {
#define LK_CASE(KV) case KV: return clamp ? launch_sw_k<KV, true>(A, n_blocks, st) : launch_sw_k<KV, false>(A, n_blocks, st);
        LK_SW_K_TABLE(LK_CASE)
#undef LK_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace lk
