#include <cstdio>
#include <cuda_runtime.h>
// Register-file cost of packed fp32x2 FMAs (Blackwell FFMA2) against scalar FFMAs as a function of how many
// DISTINCT register operands each instruction reads.  Scalar FFMAs with three distinct register sources issue
// at 2/3 rate on B200 (ffma3.cu); does an FFMA2 with three distinct 64-bit sources keep one FMA per lane-cycle?
template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, const float *in, int iters) {
    float2 x[12], c[12], d[12];
    for (int i = 0; i < 12; i++) {
        x[i] = make_float2(in[threadIdx.x + i], in[threadIdx.x + i + 1]);
        c[i] = make_float2(in[64 + threadIdx.x + i], in[65 + threadIdx.x + i]);
        d[i] = make_float2(in[128 + threadIdx.x + i], in[129 + threadIdx.x + i]);
    }
    const float2 a = make_float2(in[0], in[1]), b = make_float2(in[2], in[3]);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int i = 0; i < 12; i++) {
                if (MODE == 0) x[i] = __ffma2_rn(x[i], a, b);                               // FFMA2: 1 distinct + 2 shared
                if (MODE == 1) x[i] = __ffma2_rn(x[i], c[i], b);                            // FFMA2: 2 distinct + 1 shared
                if (MODE == 2) x[i] = __ffma2_rn(x[i], c[i], d[i]);                         // FFMA2: 3 distinct
                if (MODE == 3) x[i] = __ffma2_rn(x[(i + 5) % 12], c[i], d[(i + 7) % 12]);   // FFMA2: 3 distinct, rotating
                if (MODE == 4) { x[i].x = fmaf(x[i].x, c[i].x, d[i].x); x[i].y = fmaf(x[i].y, c[i].y, d[i].y); }  // scalar, 3 distinct
                if (MODE == 5) { x[i].x = fmaf(x[(i + 5) % 12].x, c[i].x, d[(i + 7) % 12].x); x[i].y = fmaf(x[(i + 5) % 12].y, c[i].y, d[(i + 7) % 12].y); }
                if (MODE == 6) x[i] = __fmul2_rn(x[(i + 5) % 12], c[i]);                    // FMUL2: 2 distinct
            }
    }
    float s = 0;
    for (int i = 0; i < 12; i++) s += x[i].x + x[i].y + c[i].x + d[i].y;
    if (s == 123.456f) out[0] = s;
}
template <int MODE> void run(const char *name, float *d, float *in) {
    const int blocks = 148 * 8, threads = 128, iters = 2048;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(d, in, iters); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); k<MODE><<<blocks, threads>>>(d, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    double fma = (double)blocks * threads * iters * 4.0 * 24.0;
    printf("%-44s %.3f ms  %.1f TFLOP/s\n", name, best, 2 * fma / best / 1e9);
}
int main() {
    float *d, *in; cudaMalloc(&d, 64); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    run<0>("FFMA2 1 distinct + 2 shared", d, in);
    run<1>("FFMA2 2 distinct + 1 shared", d, in);
    run<2>("FFMA2 3 distinct", d, in);
    run<3>("FFMA2 3 distinct, rotating", d, in);
    run<4>("FFMA scalar 3 distinct", d, in);
    run<5>("FFMA scalar 3 distinct, rotating", d, in);
    run<6>("FMUL2 2 distinct, rotating", d, in);
    return 0;
}
