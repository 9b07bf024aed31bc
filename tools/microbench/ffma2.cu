#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k_ffma(float *out, int iters, float a, float b) {
    float x[16];
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
    }
    float s = 0; for (int i = 0; i < 16; i++) s += x[i];
    if (s == 123.456f) out[0] = s;
}
__global__ void __launch_bounds__(256) k_ffma2(float *out, int iters, float a, float b) {
    float2 x[8];
    for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 1e-3f + i + 0.5f);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = __ffma2_rn(x[i], a2, b2);
    }
    float s = 0; for (int i = 0; i < 8; i++) s += x[i].x + x[i].y;
    if (s == 123.456f) out[0] = s;
}
// mixed: FFMA2 + independent ALU ops (LOP3/SEL) to see whether they co-issue
__global__ void __launch_bounds__(256) k_mix(float *out, int iters, float a, float b, unsigned m) {
    float2 x[8]; unsigned y[8];
    for (int i = 0; i < 8; i++) { x[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 1e-3f + i + 0.5f); y[i] = threadIdx.x + i; }
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int i = 0; i < 8; i++) { x[i] = __ffma2_rn(x[i], a2, b2); y[i] = (y[i] & m) ^ (y[(i + 1) & 7] | 0x55u); }
    }
    float s = 0; for (int i = 0; i < 8; i++) s += x[i].x + x[i].y + y[i];
    if (s == 123.456f) out[0] = s;
}
template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    return best;
}
int main() {
    float *d; cudaMalloc(&d, 64);
    const int blocks = 148 * 8, threads = 256, iters = 2048;
    double fma_per_launch = (double)blocks * threads * iters * 8.0 * 16.0;
    float t1 = timeit([&] { k_ffma<<<blocks, threads>>>(d, iters, 0.999f, 1e-3f); });
    float t2 = timeit([&] { k_ffma2<<<blocks, threads>>>(d, iters, 0.999f, 1e-3f); });
    float t3 = timeit([&] { k_mix<<<blocks, threads>>>(d, iters, 0.999f, 1e-3f, 0xF0F0F0F0u); });
    printf("FFMA  : %.3f ms  %.1f TFLOP/s\n", t1, 2 * fma_per_launch / t1 / 1e9);
    printf("FFMA2 : %.3f ms  %.1f TFLOP/s (same FMA count, half the instructions)\n", t2, 2 * fma_per_launch / t2 / 1e9);
    printf("FFMA2 + 1 LOP3 per FFMA2: %.3f ms  %.1f TFLOP/s\n", t3, 2 * fma_per_launch / t3 / 1e9);
    return 0;
}
