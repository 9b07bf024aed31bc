// colstrip.cu — inner loop of the "column strip" PairHMM layout as a standalone throughput probe.
//
// Each lane owns K consecutive HAPLOTYPE columns (M', I', D' of the current read row in registers) and the
// warp sweeps the READ rows as a skewed wavefront.  The per-row constants of the scaled hat form are then
// the same for all K cells a lane updates in a step: they are loaded once per step (two LDS.128) and sit in
// the same operand slot of K consecutive FFMAs, so the operand-reuse cache serves them and every FFMA reads
// only two fresh registers (a 3-distinct-register FFMA issues at 2/3 rate on B200, see ffma3.cu).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o colstrip.bin colstrip.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

template <int K, int MINB, int VAR>
__global__ void __launch_bounds__(128, MINB) colstrip(const float4 *__restrict__ tabA_g, const float4 *__restrict__ tabB_g,
                                                      const uint32_t *__restrict__ hap_g, float *out, int R, int reps, int lanes_per_hap) {
    extern __shared__ float4 sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int RP = R + 2;
    float4 *tA = sm + warp * 2 * RP;
    float4 *tB = tA + RP;
    for (int i = lane; i < RP; i += 32) { tA[i] = tabA_g[i]; tB[i] = tabB_g[i]; }
    __syncwarp();
    const int lig = lane % lanes_per_hap;
    const bool is_start = lig == 0;
    constexpr int KW = (K + 3) / 4;
    uint32_t hp[KW];
#pragma unroll
    for (int x = 0; x < KW; x++) hp[x] = hap_g[(blockIdx.x * 128 + threadIdx.x) % 1024 * KW + x];
    float total = 0.f;
    for (int rep = 0; rep < reps; rep++) {
        float M[K], I[K], D[K];
#pragma unroll
        for (int k = 0; k < K; k++) { M[k] = 0.f; I[k] = 0.f; D[k] = 1e20f; }
        float curM = 0.f, curI = 0.f, curD = 1e20f, prevM, prevI, prevD;
        const int n_steps = R + lanes_per_hap - 1;
        int row = 1 - lig;
        float4 ca = tA[max(0, min(row, R + 1))], cb = tB[max(0, min(row, R + 1))];
        for (int s = 1; s <= n_steps; ++s, ++row) {
            prevM = curM; prevI = curI; prevD = curD;
            curM = __shfl_up_sync(0xffffffffu, M[K - 1], 1);
            curI = __shfl_up_sync(0xffffffffu, I[K - 1], 1);
            curD = __shfl_up_sync(0xffffffffu, D[K - 1], 1);
            if (is_start) { curM = 0.f; curI = 0.f; curD = 0.f; }
            const int nrow = max(0, min(row + 1, R + 1));
            const float4 na = tA[nrow], nb = tB[nrow];
            if (row >= 1 && row <= R) {
                const float c_b = ca.x, c_c = ca.y, c_e = ca.z, c_f = ca.w;
                const float c_s = cb.x, c_sr = cb.y;
                const uint32_t rmask = __float_as_uint(cb.z);
                float pm = prevM, pi = prevI, pd = prevD, lm = curM, ld = curD;
#pragma unroll
                for (int k = 0; k < K; k++) {
                    const float om = M[k], oi = I[k], od = D[k];
                    float mm = fmaf(pi, c_b, pm);
                    mm = fmaf(pd, c_c, mm);
                    float mn;
                    if (VAR == 0) {  // full cell: compare + select + FMUL
                        const uint32_t match = hp[k >> 2] & rmask & (0xFFu << (8 * (k & 3)));
                        mn = mm * (match ? c_s : c_sr);
                    } else if (VAR == 1) {  // no compare / select: 5 FMA-pipe instructions
                        mn = mm * c_s;
                    } else if (VAR == 2) {  // 4 FFMAs only
                        mn = mm;
                    } else if (VAR == 3) {  // select on a per-step predicate (no per-cell compare)
                        mn = mm * ((rmask & hp[0]) ? c_s : c_sr);
                    } else if (VAR == 4) {  // compare feeds a predicated FMUL instead of FSEL
                        const uint32_t match = hp[k >> 2] & rmask & (0xFFu << (8 * (k & 3)));
                        asm("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tmul.f32 %0, %2, %4;\n\t@p mul.f32 %0, %2, %3;\n\t}" : "=&f"(mn) : "r"(match), "f"(mm), "f"(c_s), "f"(c_sr));
                    } else {  // VAR 5: arithmetic select: prior factor from an integer-indexed pair via FFMA (match as 0/1 float)
                        const uint32_t match = hp[k >> 2] & rmask & (0xFFu << (8 * (k & 3)));
                        mn = match ? mm * c_s : mm * c_sr;
                    }
                    if (VAR == 6) { mn = mm; I[k] = fmaf(oi, c_e, om); D[k] = fmaf(od, c_f, om); }       // 4 FFMAs, no serial chain
                    else if (VAR == 7) { mn = mm; }                                                        // M part only (2 FFMAs)
                    else if (VAR == 8) { mn = om + 0.f * mm; I[k] = fmaf(oi, c_e, om); D[k] = fmaf(ld, c_f, lm); }
                    else { I[k] = fmaf(oi, c_e, om); D[k] = fmaf(ld, c_f, lm); }
                    M[k] = mn;
                    pm = om; pi = oi; pd = od; lm = mn; ld = D[k];
                }
            }
            ca = na; cb = nb;
        }
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < K; k++) acc += M[k] + I[k];
        total += acc;
    }
    out[blockIdx.x * 128 + threadIdx.x] = total;
}

template <int K, int MINB, int VAR>
void run(int blocks_per_sm, int R, int lanes_per_hap) {
    const int blocks = 148 * blocks_per_sm, reps = 64;
    const int RP = R + 2;
    float4 *ta, *tb; uint32_t *hap; float *out;
    cudaMalloc(&ta, RP * 16); cudaMalloc(&tb, RP * 16); cudaMalloc(&hap, 1024 * 16 * 4); cudaMalloc(&out, blocks * 128 * 4);
    float4 *ha = (float4 *)malloc(RP * 16), *hb = (float4 *)malloc(RP * 16);
    uint32_t *hh = (uint32_t *)malloc(1024 * 16 * 4);
    srand(1);
    for (int i = 0; i < RP; i++) {
        ha[i] = make_float4(1e-4f, 1e-4f, 0.1f, 0.1f);
        const uint32_t m = (1u << (rand() & 3)) * 0x01010101u;
        hb[i] = make_float4(0.99f, 1e-3f, 0.f, 0.f);
        ((uint32_t *)&hb[i])[2] = m;
    }
    for (int i = 0; i < 1024 * 16; i++) { uint32_t w = 0; for (int b = 0; b < 4; b++) w |= (1u << (rand() & 3)) << (8 * b); hh[i] = w; }
    cudaMemcpy(ta, ha, RP * 16, cudaMemcpyHostToDevice); cudaMemcpy(tb, hb, RP * 16, cudaMemcpyHostToDevice);
    cudaMemcpy(hap, hh, 1024 * 16 * 4, cudaMemcpyHostToDevice);
    const size_t smem = 4 * 2 * RP * 16;
    cudaFuncSetAttribute(colstrip<K, MINB, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    colstrip<K, MINB, VAR><<<blocks, 128, smem>>>(ta, tb, hap, out, R, reps, lanes_per_hap);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("K=%d error %s\n", K, cudaGetErrorString(e)); return; }
    float best = 1e9;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0); colstrip<K, MINB, VAR><<<blocks, 128, smem>>>(ta, tb, hap, out, R, reps, lanes_per_hap); cudaEventRecord(e1);
        cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    // useful cells: every lane-step with 1 <= row <= R updates K cells
    const double cells = (double)blocks * 128 * reps * (double)R * K;
    const double cyc_per_cell = best * 1e-3 * 1.965e9 * 148 * 4 / ((double)blocks * 4 * reps * (double)R * K) ;
    printf("var=%d K=%2d minb=%d blocks/SM=%d R=%d lanes/hap=%d: %.3f ms  %.1f GCUPS  (%.2f SMSP-cycles per warp-cell at 1965 MHz)\n", VAR, K, MINB, blocks_per_sm, R,
           lanes_per_hap, best, cells / best / 1e6, cyc_per_cell);
    cudaFree(ta); cudaFree(tb); cudaFree(hap); cudaFree(out); free(ha); free(hb); free(hh);
}

int main() {
    run<32, 2, 0>(1, 150, 8);
    run<32, 2, 0>(2, 150, 8);
    run<32, 2, 2>(1, 150, 8);
    run<32, 2, 2>(2, 150, 8);
    run<32, 2, 6>(1, 150, 8);
    run<32, 2, 6>(2, 150, 8);
    run<32, 2, 7>(1, 150, 8);
    run<32, 2, 7>(2, 150, 8);
    run<24, 3, 0>(1, 150, 8);
    run<24, 3, 2>(1, 150, 8);
    run<24, 3, 2>(3, 150, 8);
    run<24, 3, 6>(3, 150, 8);
    run<24, 3, 7>(3, 150, 8);
    run<16, 4, 0>(4, 150, 8);
    run<16, 4, 2>(4, 150, 8);
    run<16, 4, 6>(4, 150, 8);
    run<16, 4, 7>(4, 150, 8);
    return 0;
}
