#include <cstdio>
#include <cuda_runtime.h>
// FFMA with a 32-bit immediate multiplier against the register form: issue rate alone and mixed.
// MODE 0: x = x*c + y (register c, reuse)   MODE 1: x = x*0.1f + y (immediate)   MODE 2: half and half
// MODE 3: x = x*y (FMUL)                    MODE 4: 2 imm FFMA + 2 reg FFMA (3 distinct) + 1 FMUL  (the PairHMM cell)
template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, const float *in, int iters) {
    float x[24], y[24], z[24];
    for (int i = 0; i < 24; i++) { x[i] = in[threadIdx.x + i]; y[i] = in[64 + threadIdx.x + i]; z[i] = in[128 + threadIdx.x + i]; }
    const float c = in[0];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 24; i++) {
                if (MODE == 0) x[i] = fmaf(x[i], c, y[i]);
                if (MODE == 1) x[i] = fmaf(x[i], 0.1f, y[i]);
                if (MODE == 2) x[i] = (i & 1) ? fmaf(x[i], 0.1f, y[i]) : fmaf(x[i], c, y[i]);
                if (MODE == 3) x[i] = x[i] * y[i];
                if (MODE == 4) {
                    if (i % 5 == 0) x[i] = fmaf(x[i], 0.1f, y[i]);
                    else if (i % 5 == 1) x[i] = fmaf(x[i], 0.1f, z[i]);
                    else if (i % 5 == 2) x[i] = fmaf(x[i], z[i], y[i]);
                    else if (i % 5 == 3) x[i] = fmaf(y[i], z[i], x[i]);
                    else x[i] = x[i] * y[i];
                }
            }
    }
    float s = 0;
    for (int i = 0; i < 24; i++) s += x[i] + y[i] + z[i];
    if (s == 123.456f) out[0] = s;
}
template <int MODE> void run(const char *name, float *d, float *in) {
    const int blocks = 148 * 16, threads = 128, iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, in, iters); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    double ops = (double)blocks * threads * iters * 4.0 * 24.0;
    printf("%-52s %.3f ms  %.2f cycles per warp-instruction per SMSP\n", name, best, best * 1e-3 * 1.965e9 * 148 * 4 * 32 / ops);
}
int main() {
    float *d, *in; cudaMalloc(&d, 64); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    run<0>("FFMA x = x*c + y (register c)", d, in);
    run<1>("FFMA x = x*0.1f + y (immediate)", d, in);
    run<2>("half immediate, half register", d, in);
    run<3>("FMUL x = x*y", d, in);
    run<4>("cell mix: 2 imm FFMA, 2 reg FFMA (3 regs), 1 FMUL", d, in);
    return 0;
}
