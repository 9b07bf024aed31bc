// tools_kernels.cu — measurement helpers (not on the hot path): an FFMA-only microbenchmark that gives the
// fp32 roofline denominator on the device the bench runs on ("re-measure with an FMA-only
// microbenchmark under sustained load and use the measured figure").
#include <cuda_runtime.h>

#include "lk_api_internal.h"

__global__ void __launch_bounds__(256) fma_peak_kernel(float *out, int iters, float a, float b) {
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f,
          x7 = x0 + 7.f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 16; u++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    if (x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 == 123.456f) out[0] = x0;  // keep the chains alive
}

// Sustained fp32 FMA throughput in TFLOP/s (2 flop per FMA), measured for ~duration_ms.
extern "C" lk_status lk_measure_fp32_fma_tflops(lk_ctx *ctx, double duration_ms, double *tflops) {
    if (!ctx || !tflops) return LK_ERR_INVALID;
    std::lock_guard<std::mutex> g(ctx->mu);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "cudaSetDevice");
    float *d = nullptr;
    if ((e = cudaMalloc(&d, 64)) != cudaSuccess) return lk_cuda_fail(ctx, e, "cudaMalloc");
    cudaStream_t st = ctx->stream[0];
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = ctx->n_sm * 8, threads = 256, iters = 4096;
    const double flop_per_launch = (double)blocks * threads * (double)iters * 16.0 * 8.0 * 2.0;
    fma_peak_kernel<<<blocks, threads, 0, st>>>(d, iters, 0.999f, 1e-3f);  // warm-up
    cudaStreamSynchronize(st);
    double best = 0.0, spent = 0.0;
    while (spent < duration_ms) {
        cudaEventRecord(e0, st);
        fma_peak_kernel<<<blocks, threads, 0, st>>>(d, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1, st);
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        spent += ms;
        // sustained figure: the LAST launches run with the clock settled under load
        best = flop_per_launch / (ms * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d);
    if (e != cudaSuccess) return lk_cuda_fail(ctx, e, "fma_peak_kernel");
    *tflops = best;
    return LK_OK;
}
