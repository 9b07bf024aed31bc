#include <cstdio>
#include <cuda_runtime.h>
// FFMA throughput as a function of how many DISTINCT register operands each instruction reads.
template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, const float *in, int iters) {
    float x[24], c[24], d[24];
    for (int i = 0; i < 24; i++) { x[i] = in[threadIdx.x + i]; c[i] = in[64 + threadIdx.x + i]; d[i] = in[128 + threadIdx.x + i]; }
    const float a = in[0], b = in[1];
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 24; i++) {
                if (MODE == 0) x[i] = fmaf(x[i], a, b);            // 1 distinct + 2 shared (reuse cache)
                if (MODE == 1) x[i] = fmaf(x[i], c[i], b);         // 2 distinct + 1 shared
                if (MODE == 2) x[i] = fmaf(x[i], c[i], d[i]);      // 3 distinct
                if (MODE == 3) x[i] = fmaf(x[(i + 7) % 24], c[i], d[(i + 11) % 24]);  // 3 distinct, dest differs from sources
            }
    }
    float s = 0;
    for (int i = 0; i < 24; i++) s += x[i] + c[i] + d[i];
    if (s == 123.456f) out[0] = s;
}
template <int MODE> void run(const char *name, float *d, float *in) {
    const int blocks = 148 * 16, threads = 128, iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, in, iters); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    double fma = (double)blocks * threads * iters * 4.0 * 24.0;
    printf("%-40s %.3f ms  %.1f TFLOP/s\n", name, best, 2 * fma / best / 1e9);
}
int main() {
    float *d, *in; cudaMalloc(&d, 64); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    run<0>("FFMA 1 distinct reg + 2 shared", d, in);
    run<1>("FFMA 2 distinct regs + 1 shared", d, in);
    run<2>("FFMA 3 distinct regs", d, in);
    run<3>("FFMA 3 distinct regs, rotating", d, in);
    return 0;
}
