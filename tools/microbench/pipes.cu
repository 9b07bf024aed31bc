// pipes.cu — issue-rate probe for the instruction mix of a PairHMM cell on sm_100a:
// FFMA / FMUL (FMA pipe) against LOP3-with-predicate and FSEL (ALU pipe), alone and mixed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes.bin pipes.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

// NF FMA-pipe ops, NL LOP3.P ops, NS FSEL ops per "cell"; 16 independent cells per loop trip.
template <int NF, int NL, int NS, bool SEL_USES_LOP>
__global__ void __launch_bounds__(128) mix(float *out, const float *in, const uint32_t *inw, int iters) {
    float x[16], y[16];
    uint32_t w[16];
    for (int i = 0; i < 16; i++) { x[i] = in[threadIdx.x + i]; y[i] = in[64 + threadIdx.x + i]; w[i] = inw[threadIdx.x + i]; }
    const float c = in[0], d = in[1], e = in[2];
    const uint32_t m0 = inw[0];
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        const uint32_t m = m0 ^ (uint32_t)it;  // loop-variant mask: nothing below can be hoisted
#pragma unroll
        for (int i = 0; i < 16; i++) {
            bool p = (it & 1);
#pragma unroll
            for (int l = 0; l < NL; l++) {
                uint32_t t;
                asm volatile("and.b32 %0, %1, %2;" : "=r"(t) : "r"(w[(i + l) & 15]), "r"(m));
                if (SEL_USES_LOP) p = (t != 0u);
                else w[(i + l) & 15] = t | 1u;  // plain LOP3 whose result stays live
            }
#pragma unroll
            for (int s = 0; s < NS; s++) {
                asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\tselp.f32 %0, %1, %2, q;\n\t}" : "=f"(y[i]) : "f"(c), "f"(e), "r"((uint32_t)p));
            }
#pragma unroll
            for (int f = 0; f < NF; f++) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(c), "f"(NS ? y[i] : d));
        }
    }
    float s = 0;
    for (int i = 0; i < 16; i++) s += x[i] + y[i] + (float)w[i];
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <int NF, int NL, int NS, bool SUL>
void run(const char *name, float *d, float *in, uint32_t *inw, int bps) {
    const int blocks = 148 * bps, threads = 128, iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix<NF, NL, NS, SUL><<<blocks, threads>>>(d, in, inw, iters); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); mix<NF, NL, NS, SUL><<<blocks, threads>>>(d, in, inw, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    const double cells = (double)bps * iters * 16.0;  // warp-cells per SMSP (4 warps per block, 1 per SMSP per block)
    printf("%-44s blocks/SM=%d  %.3f ms  %.2f cycles per warp-cell (%d FMA + %d LOP3 + %d FSEL)\n", name, bps, best,
           best * 1e-3 * 1.965e9 / cells, NF, NL, NS);
}
int main() {
    float *d, *in; uint32_t *inw;
    cudaMalloc(&d, 148 * 16 * 128 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&inw, 4096); cudaMemset(inw, 0, 4096);
    for (int bps = 2; bps <= 4; bps += 2) {
        run<5, 0, 0, false>("5 FMA", d, in, inw, bps);
        run<0, 2, 0, false>("2 LOP3", d, in, inw, bps);
        run<0, 0, 2, false>("2 FSEL", d, in, inw, bps);
        run<0, 1, 1, true>("1 LOP3.P + 1 FSEL", d, in, inw, bps);
        run<5, 1, 0, false>("5 FMA + 1 LOP3", d, in, inw, bps);
        run<5, 0, 1, false>("5 FMA + 1 FSEL", d, in, inw, bps);
        run<5, 1, 1, true>("5 FMA + LOP3.P + FSEL (cell mix)", d, in, inw, bps);
        run<5, 2, 2, true>("5 FMA + 2 LOP3 + 2 FSEL", d, in, inw, bps);
        run<10, 1, 1, true>("10 FMA + LOP3.P + FSEL", d, in, inw, bps);
    }
    return 0;
}
