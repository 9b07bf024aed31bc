/* lk_oracle_simd.c — TEST / BASELINE INFRASTRUCTURE, like the rest of oracle/: OUR fp32 SIMD PairHMM standing in for the
 * AVX path the reference takes through the un-vendored crate gkl (gkl::pairhmm::forward, src/pair_hmm/pair_hmm.rs:348-366;
 * Cargo.toml:42 `gkl = "^0.1.1"`, no lockfile, source not on this machine).  It is NOT gkl: it is an anti-diagonal fp32
 * kernel in the style Intel GKL documents (fp32 first, f64 redo on underflow), written against GCC vector extensions and
 * compiled three times -- 16 lanes under target("avx512f"), 8 lanes under target("avx2,fma"), 4 lanes for plain x86-64 --
 * with the widest one the CPU supports chosen at run time.  Used only as the second cpu_baseline of bench.py and by
 * `bench.py --impl reference --ref-kind vector`; checked against the 104 golden rows at the reference's 1e-5
 * (tests/test_oracle_simd.py).  Only tests/, smoke() and bench.py's CPU legs may use anything under oracle/. */
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "lk_oracle.h"

/* Per-row constants of one read, rows padded to a multiple of 16 with zeros (a padding row computes zeros). */
typedef struct {
    int R, cap;
    float *tMM, *tGM, *tMI, *tII, *tMD, *pm, *px;
    int32_t *rb, *rN;
} simd_read_t;

static void simd_read_free(simd_read_t *r) {
    free(r->tMM); free(r->tGM); free(r->tMI); free(r->tII); free(r->tMD); free(r->pm); free(r->px); free(r->rb); free(r->rN);
    memset(r, 0, sizeof *r);
}

/* transitions pair_hmm_model.rs:142-156, priors pair_hmm.rs:638-654, once per read */
static void simd_read_build(simd_read_t *r, const uint8_t *read, const uint8_t *qual, const uint8_t *ins, const uint8_t *del,
                            const uint8_t *gcp, int R, int use_tristate) {
    const int padded = ((R + 15) / 16) * 16 + 16;
    if (padded > r->cap) {
        simd_read_free(r);
        r->cap = padded;
        r->tMM = malloc(sizeof(float) * (size_t)padded); r->tGM = malloc(sizeof(float) * (size_t)padded);
        r->tMI = malloc(sizeof(float) * (size_t)padded); r->tII = malloc(sizeof(float) * (size_t)padded);
        r->tMD = malloc(sizeof(float) * (size_t)padded); r->pm = malloc(sizeof(float) * (size_t)padded);
        r->px = malloc(sizeof(float) * (size_t)padded); r->rb = malloc(sizeof(int32_t) * (size_t)padded);
        r->rN = malloc(sizeof(int32_t) * (size_t)padded);
    }
    r->R = R;
    for (int i = 0; i < padded; i++) {
        if (i < R) {
            const double err = lko_qual_to_error_prob(qual[i]);
            r->tMM[i] = (float)lko_match_to_match_prob(ins[i], del[i]);
            r->tGM[i] = (float)lko_qual_to_prob(gcp[i]);
            r->tMI[i] = (float)lko_qual_to_error_prob(ins[i]);
            r->tII[i] = (float)lko_qual_to_error_prob(gcp[i]);
            r->tMD[i] = (float)lko_qual_to_error_prob(del[i]);
            r->pm[i] = (float)(1.0 - err);
            r->px[i] = (float)(err / (use_tristate ? 3.0 : 1.0));
            r->rb[i] = read[i];
            r->rN[i] = read[i] == 'N' ? -1 : 0;
        } else {
            r->tMM[i] = r->tGM[i] = r->tMI[i] = r->tII[i] = r->tMD[i] = r->pm[i] = r->px[i] = 0.f;
            r->rb[i] = 0x100; r->rN[i] = 0;
        }
    }
}

#define SIMD_L 16
#define SIMD_NAME forward_f32_l16
#define SIMD_TARGET __attribute__((target("avx512f,avx512bw,avx512dq,avx512vl")))
#include "lk_oracle_simd_kernel.inc"

#define SIMD_L 8
#define SIMD_NAME forward_f32_l8
#define SIMD_TARGET __attribute__((target("avx2,fma")))
#include "lk_oracle_simd_kernel.inc"

#define SIMD_L 4
#define SIMD_NAME forward_f32_l4
#define SIMD_TARGET
#include "lk_oracle_simd_kernel.inc"

typedef double (*forward_fn)(const uint8_t *, int, const simd_read_t *, float *, uint8_t *);

/* fp32 cells far from the alignment path decay into the denormal range, where x86 arithmetic takes a microcode assist per
 * operation: flush-to-zero + denormals-are-zero for the duration of the SIMD kernel (as Intel GKL does); such values are
 * 2^-126 of a 2^100-scaled sum and cannot move the result.  The previous MXCSR is restored for the f64 scalar code. */
static unsigned simd_ftz_on(void) {
    const unsigned old = __builtin_ia32_stmxcsr();
    __builtin_ia32_ldmxcsr(old | 0x8040u);
    return old;
}
static void simd_ftz_restore(unsigned old) { __builtin_ia32_ldmxcsr(old); }

static int g_forced_lanes = 0;
/* 16 / 8 / 4: force that variant (tests); 0: the widest the CPU supports */
void lko_simd_force_lanes(int lanes) { g_forced_lanes = lanes; }

/* 1 if the CPU can run the variant (a forced variant the CPU lacks would fault) */
int lko_simd_supported(int lanes) {
    __builtin_cpu_init();
    if (lanes == 16) return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512dq") && __builtin_cpu_supports("avx512vl");
    if (lanes == 8) return __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
    return lanes == 4;
}

static forward_fn pick(int *lanes) {
    int l = g_forced_lanes;
    if (l == 0) l = lko_simd_supported(16) ? 16 : (lko_simd_supported(8) ? 8 : 4);
    if (lanes) *lanes = l;
    return l == 16 ? forward_f32_l16 : (l == 8 ? forward_f32_l8 : forward_f32_l4);
}

/* "avx512 x16 fp32", "avx2+fma x8 fp32" or "sse2 x4 fp32" */
const char *lko_simd_isa(void) {
    int l = 0;
    (void)pick(&l);
    return l == 16 ? "avx512 x16 fp32" : (l == 8 ? "avx2+fma x8 fp32" : "sse2 x4 fp32");
}

/* one pair: fp32 SIMD, scalar f64 restatement when fp32 underflows; *redone (may be NULL) counts the f64 redos */
double lko_pairhmm_forward_vector(const uint8_t *hap, int hap_len, const uint8_t *read, const uint8_t *qual, const uint8_t *ins,
                                  const uint8_t *del, const uint8_t *gcp, int read_len, int use_tristate, int *redone) {
    float *bnd = (float *)malloc(sizeof(float) * 3 * (size_t)(hap_len + 2));
    uint8_t *hrev = (uint8_t *)malloc((size_t)hap_len + 64);
    simd_read_t rd;
    memset(&rd, 0, sizeof rd);
    simd_read_build(&rd, read, qual, ins, del, gcp, read_len, use_tristate);
    const unsigned mx = simd_ftz_on();
    double v = pick(NULL)(hap, hap_len, &rd, bnd, hrev);
    simd_ftz_restore(mx);
    free(bnd); free(hrev);
    simd_read_free(&rd);
    if (v > 0.0) {
        if (redone) (*redone)++;
        v = lko_pairhmm_forward(hap, hap_len, read, qual, ins, del, gcp, read_len, use_tristate);
    }
    return v;
}

typedef struct {
    int n_regions;
    const int64_t *region_read_start, *region_hap_start, *read_off, *hap_off, *out_off;
    const uint8_t *read_bases, *read_qual, *read_ins, *read_del, *read_gcp, *hap_bases;
    int use_tristate;
    double *out;
    atomic_int next;
    atomic_int redone;
    forward_fn fn;
} vjob_t;

static void *vworker(void *arg) {
    vjob_t *j = (vjob_t *)arg;
    float *bnd = NULL;
    uint8_t *hrev = NULL;
    int cap = 0;
    simd_read_t rd;
    memset(&rd, 0, sizeof rd);
    for (;;) {
        const int g = atomic_fetch_add(&j->next, 1);
        if (g >= j->n_regions) break;
        const int64_t r0 = j->region_read_start[g], r1 = j->region_read_start[g + 1];
        const int64_t h0 = j->region_hap_start[g], h1 = j->region_hap_start[g + 1];
        const int A = (int)(h1 - h0);
        /* the reference's loop order: for read { for haplotype } (pair_hmm.rs:349-369), result [r * A + h] */
        for (int64_t r = r0; r < r1; r++) {
            const int64_t ro = j->read_off[r];
            const int R = (int)(j->read_off[r + 1] - ro);
            simd_read_build(&rd, j->read_bases + ro, j->read_qual + ro, j->read_ins + ro, j->read_del + ro, j->read_gcp + ro, R, j->use_tristate);
            for (int64_t h = h0; h < h1; h++) {
                const int64_t ho = j->hap_off[h];
                const int H = (int)(j->hap_off[h + 1] - ho);
                if (H > cap) {
                    cap = H + 64;
                    free(bnd); free(hrev);
                    bnd = (float *)malloc(sizeof(float) * 3 * (size_t)(cap + 2));
                    hrev = (uint8_t *)malloc((size_t)cap + 64);
                }
                const unsigned mx = simd_ftz_on();
                double v = j->fn(j->hap_bases + ho, H, &rd, bnd, hrev);
                simd_ftz_restore(mx);
                if (v > 0.0) {
                    atomic_fetch_add(&j->redone, 1);
                    v = lko_pairhmm_forward(j->hap_bases + ho, H, j->read_bases + ro, j->read_qual + ro, j->read_ins + ro,
                                            j->read_del + ro, j->read_gcp + ro, R, j->use_tristate);
                }
                j->out[j->out_off[g] + (r - r0) * A + (h - h0)] = v;
            }
        }
    }
    free(bnd); free(hrev);
    simd_read_free(&rd);
    return NULL;
}

/* Same contract as lko_pairhmm_batch_mt (one region per task, the granularity Rayon gives the reference); returns the number
 * of pairs redone in f64. */
int lko_pairhmm_batch_vector_mt(int n_regions, const int64_t *region_read_start, const int64_t *region_hap_start,
                                const int64_t *read_off, const uint8_t *read_bases, const uint8_t *read_qual,
                                const uint8_t *read_ins, const uint8_t *read_del, const uint8_t *read_gcp,
                                const int64_t *hap_off, const uint8_t *hap_bases, const int64_t *out_off,
                                int use_tristate, int n_threads, double *out) {
    lko_init();
    vjob_t job = { n_regions, region_read_start, region_hap_start, read_off, hap_off, out_off, read_bases, read_qual, read_ins,
                   read_del, read_gcp, hap_bases, use_tristate, out, 0, 0, pick(NULL) };
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, vworker, &job);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th);
    return atomic_load(&job.redone);
}
