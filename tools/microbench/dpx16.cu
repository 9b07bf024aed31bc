// dpx16.cu — issue-rate probe for the integer ops of a Smith-Waterman cell on sm_100a: the 32-bit DPX forms the SW kernel
// uses today (VIADDMNMX, VIMNMX3, SHF, IMAD.IADD) against the packed 16x2 forms (VIADD.16x2, VIADDMNMX.S16x2,
// VIMNMX3.S16x2, VIMNMX.S16x2 with predicate outputs) and the PRMT / LOP3 pair that gathers sign bits.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dpx16.bin dpx16.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

enum Op { ADDMAX32, MAX3_32, SHF32, IADD32, ADD16, ADDMAX16, MAX3_16, BMAX16, PRMT_, LOP3_, CELL32, CELL16, CELL16_M0 = 16 };

template <int OP>
__device__ __forceinline__ void one(uint32_t &x, uint32_t &y, uint32_t c, uint32_t d) {
    if (OP == ADDMAX32) x = (uint32_t)__viaddmax_s32((int)x, (int)c, (int)d);
    if (OP == MAX3_32) x = (uint32_t)__vimax3_s32((int)x, (int)c, (int)d);
    if (OP == SHF32) x = __funnelshift_l(c, x, 1);
    if (OP == IADD32) x = x + c;
    if (OP == ADD16) x = __vadd2(x, c);
    if (OP == ADDMAX16) x = __viaddmax_s16x2(x, c, d);
    if (OP == MAX3_16) x = __vimax3_s16x2(x, c, d);
    if (OP == BMAX16) { bool ph, pl; x = __vibmax_s16x2(x, c, &ph, &pl); if (ph) y |= 0x10000u; if (pl) y |= 1u; }
    if (OP == PRMT_) x = __byte_perm(x, c, 0xfb73);
    if (OP == LOP3_) x = (x & c) | d;
}

// today's cell: 3 adds, 2 VIADDMNMX, 1 VIMNMX3, 4 subtractions, 4 funnel shifts (14 ops, one cell)
__device__ __forceinline__ void cell32(int &h, int &e, int &hd, int &hu, int &fu, uint32_t &pm, uint32_t &pg, int wt, int wo, int we) {
    const int hold = h, sd = hd + wt, ev = fu + we, dn = __viaddmax_s32(hu, wo, ev), eh = e + we, rt = __viaddmax_s32(hold, wo, eh);
    const int best = __vimax3_s32(sd, rt, dn);
    pg = __funnelshift_l((uint32_t)(ev - dn), pg, 1);
    pg = __funnelshift_l((uint32_t)(eh - rt), pg, 1);
    pm = __funnelshift_l((uint32_t)(sd - best), pm, 1);
    pm = __funnelshift_l((uint32_t)(rt - dn), pm, 1);
    e = rt; h = best; hd = hold; hu = best; fu = dn;
}
// packed difference cell (two pairs per word, unsigned biased halves, no carry between halves): the cell of sw16_core.cuh.
// Two rows at a time so that one PRMT + one LOP3 gather four flag bits of each kind.
// a + b and a - b issued as IMADs (FMA pipe) through a multiplier the compiler cannot see through
__device__ __forceinline__ uint32_t fadd(uint32_t a, uint32_t b, uint32_t one) { uint32_t d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b)); return d; }
__device__ __forceinline__ uint32_t fsub(uint32_t a, uint32_t b, uint32_t mone) { uint32_t d; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(mone), "r"(a)); return d; }
__device__ __forceinline__ uint32_t signs4(uint32_t a, uint32_t b) { uint32_t d; asm("prmt.b32 %0, %1, %2, 0xfdb9;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t fadd_v(uint32_t a, uint32_t b, uint32_t one) { uint32_t d; asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(b)); return d; }
// MODE: which of the three-input sums go to the ALU pipe as one IADD3 (the rest are two IMADs): bit 0 a, bit 1 b, bit 2 dpd
template <int MODE>
__device__ __forceinline__ void cell16x2(uint32_t &u0, uint32_t &y0, uint32_t &u1, uint32_t &y1, uint32_t &v, uint32_t &x,
                                         uint32_t &a_pd, uint32_t &a_pr, uint32_t &a_vo, uint32_t &a_ho, uint32_t tA0, uint32_t tB0,
                                         uint32_t tA1, uint32_t tB1, uint32_t Ca, uint32_t D, uint32_t cflag, uint32_t c7f, uint32_t one,
                                         uint32_t mone, uint32_t msk) {
    uint32_t dpd[2], dpr[2], dvo[2], dho[2];
    uint32_t *uu[2] = { &u0, &u1 }, *yy[2] = { &y0, &y1 };
    const uint32_t ss[2] = { fadd_v(tA0, tB0, one), fadd_v(tA1, tB1, one) };
#pragma unroll
    for (int r = 0; r < 2; r++) {
        uint32_t &u = *uu[r], &y = *yy[r];
        const uint32_t a = (MODE & 1) ? v - x + Ca : fadd(fsub(v, x, mone), Ca, one);
        const uint32_t b = (MODE & 2) ? u - y + Ca : fadd(fsub(u, y, mone), Ca, one);
        const uint32_t z = __vimax3_u16x2(ss[r], a, b);
        const uint32_t un = fsub(z, v, mone), vn = fsub(z, u, mone), nf = fsub(z, a, mone), ne = fsub(z, b, mone);
        dpd[r] = (MODE & 4) ? z - ss[r] + 0x7fff7fffu : fadd(fsub(z, ss[r], mone), c7f, one);
        dpr[r] = fadd(ne, c7f, one);
        dvo[r] = fadd(nf, cflag, one);
        dho[r] = fadd(ne, cflag, one);
        x = __vminu2(nf, D); y = __vminu2(ne, D);
        u = un; v = vn;
    }
    a_pd |= signs4(dpd[0], dpd[1]) & msk;
    a_pr |= signs4(dpr[0], dpr[1]) & msk;
    a_vo |= signs4(dvo[0], dvo[1]) & msk;
    a_ho |= signs4(dho[0], dho[1]) & msk;
}

template <int OP>
__global__ void __launch_bounds__(128) probe(uint32_t *out, const uint32_t *in, int iters) {
    uint32_t x[16], y[16];
    for (int i = 0; i < 16; i++) { x[i] = in[threadIdx.x + i]; y[i] = in[64 + threadIdx.x + i]; }
    uint32_t c = in[0], d = in[1];
    if (OP == CELL32) {
        int h[8], e[8];
        for (int i = 0; i < 8; i++) { h[i] = x[i]; e[i] = y[i]; }
        int hd = in[3], hu = in[4], fu = in[5];
        uint32_t pm = 0, pg = 0;
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
            int hdd = hd, huu = hu, fuu = fu;
#pragma unroll
            for (int i = 0; i < 8; i++) cell32(h[i], e[i], hdd, huu, fuu, pm, pg, (int)c + (it & 3), (int)d, -11);
            hd = huu; hu = hdd; fu = fuu;
        }
        for (int i = 0; i < 8; i++) { x[i] = h[i]; y[i] = e[i]; }
        x[0] += pm + pg + hd + hu + fu;
    } else if (OP >= CELL16_M0) {
        uint32_t v = in[3], xx = in[4];
        uint32_t t[16];
        for (int i = 0; i < 16; i++) t[i] = in[128 + threadIdx.x + i];
        const uint32_t Ca = in[6], D = in[7], cf = in[8], c7f = in[9], one = in[10], mone = in[11];
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
            uint32_t vv = v, xv = xx, ap = 0, aq = 0, ar = 0, as = 0;
#pragma unroll
            for (int i = 0; i < 8; i += 2)
                cell16x2<OP - CELL16_M0>(x[i], y[i], x[i + 1], y[i + 1], vv, xv, ap, aq, ar, as, t[i], t[8 + i], t[i + 1], t[9 + i], Ca, D, cf, c7f, one, mone,
                         0x01010101u << (i >> 1));
            v = xv ^ ap ^ aq; xx = vv ^ ar ^ as;
        }
        x[0] += v + xx;
    } else {
#pragma unroll 1
        for (int it = 0; it < iters; it++) {
            const uint32_t cc = c ^ (uint32_t)it;
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 16; i++) one<OP>(x[i], y[i], cc, d);
        }
    }
    uint32_t s = 0;
    for (int i = 0; i < 16; i++) s += x[i] + y[i];
    out[blockIdx.x * 128 + threadIdx.x] = s;
}

template <int OP>
void run(const char *name, uint32_t *d, uint32_t *in, int bps, double units_per_iter, const char *unit) {
    const int blocks = 148 * bps, threads = 128, iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<OP><<<blocks, threads>>>(d, in, iters); cudaDeviceSynchronize();
    float best = 1e9;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); probe<OP><<<blocks, threads>>>(d, in, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best; }
    const double units = (double)bps * iters * units_per_iter;  // warp-level units per SMSP (one warp of each block per SMSP)
    printf("%-34s blocks/SM=%d  %.3f ms  %.3f cycles per warp-%s per SMSP\n", name, bps, best, best * 1e-3 * 1.965e9 / units, unit);
}
__global__ void semantics(uint32_t *o) {
    o[0] = __vimax3_u16x2(0x80000001u, 0x7fff0002u, 0x0001ffffu);   // expect 0x8000ffff (unsigned halves)
    o[1] = __vminu2(0x80000001u, 0x7fffffffu);                      // expect 0x7fff0001
    o[2] = signs4(0x80007fffu, 0x0000ff00u);           // signs of the four halves: expect 0xff0000ff -> bytes [lo(a),hi(a),lo(b),hi(b)] = 00,ff,ff,00 -> 0x00ffff00
    o[3] = __viaddmax_u16x2(0xfff00005u, 0x00200003u, 0x00010009u); // wraps per half?  hi: 0xfff0+0x20 = 0x10 (wrap) vs 1 -> 0x0010; lo: 8 vs 9 -> 9
}
int main() {
    { uint32_t *o, h[4]; cudaMalloc(&o, 16); semantics<<<1, 1>>>(o); cudaMemcpy(h, o, 16, cudaMemcpyDeviceToHost);
      printf("semantics: max3_u16x2=%08x minu2=%08x prmt_signs=%08x addmax_u16x2=%08x\n", h[0], h[1], h[2], h[3]); }
    uint32_t *d, *in;
    cudaMalloc(&d, 148 * 16 * 128 * 4); cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096);
    for (int bps = 2; bps <= 4; bps += 2) {
        run<ADDMAX32>("VIADDMNMX (32)", d, in, bps, 64, "op");
        run<MAX3_32>("VIMNMX3 (32)", d, in, bps, 64, "op");
        run<SHF32>("SHF.L.W", d, in, bps, 64, "op");
        run<IADD32>("IADD", d, in, bps, 64, "op");
        run<ADD16>("VIADD.16x2", d, in, bps, 64, "op");
        run<ADDMAX16>("VIADDMNMX.S16x2", d, in, bps, 64, "op");
        run<MAX3_16>("VIMNMX3.S16x2", d, in, bps, 64, "op");
        run<BMAX16>("VIMNMX.S16x2 + 2 predicated ORs", d, in, bps, 64, "group");
        run<PRMT_>("PRMT", d, in, bps, 64, "op");
        run<LOP3_>("LOP3", d, in, bps, 64, "op");
        run<CELL32>("today's cell (32-bit, 1 cell)", d, in, bps, 8, "cell");
        run<CELL16_M0 + 7>("packed cell, a b dpd as IADD3", d, in, bps, 16, "cell");
        run<CELL16_M0 + 5>("packed cell, a dpd as IADD3", d, in, bps, 16, "cell");
        run<CELL16_M0 + 3>("packed cell, a b as IADD3", d, in, bps, 16, "cell");
        run<CELL16_M0 + 1>("packed cell, a as IADD3", d, in, bps, 16, "cell");
        run<CELL16_M0 + 4>("packed cell, dpd as IADD3", d, in, bps, 16, "cell");
        run<CELL16_M0 + 0>("packed cell, all IMAD", d, in, bps, 16, "cell");
    }
    return 0;
}
