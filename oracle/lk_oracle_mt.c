/*
 * lk_oracle_mt.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * Multi-threaded batch drivers over lk_oracle.c, used (a) by tests to produce
 * expected values for whole batches quickly and (b) by bench.py's
 * cpu_baseline / --impl reference leg.  Parallel granularity follows the
 * reference: one active region per worker task
 * (assembly/assembly_region_walker.rs:140,211: pending_regions.into_par_iter()),
 * with the (read x haplotype) double loop inside a region serial as in
 * pair_hmm/pair_hmm.rs:349-369.  SW pairs are one task each (the reference
 * calls SmithWatermanAligner::align once per pair from those same tasks).
 */
#include "lk_oracle.h"

#include <pthread.h>
#include <stdatomic.h>
#include <stdlib.h>

void lko_init(void) {
    /* force the lazy tables before threads start */
    (void)lko_match_to_match_prob(10, 10);
    uint8_t b = 'A', q = 30;
    (void)lko_pairhmm_forward(&b, 1, &b, &q, &q, &q, &q, 1, 1);
}

typedef struct {
    int n_regions;
    const int64_t *region_read_start, *region_hap_start, *read_off, *hap_off, *out_off;
    const uint8_t *read_bases, *read_qual, *read_ins, *read_del, *read_gcp, *hap_bases;
    int use_tristate, rescaled;
    double *out;
    atomic_int next;
} hmm_job_t;

static void *hmm_worker(void *arg) {
    hmm_job_t *j = (hmm_job_t *)arg;
    for (;;) {
        int g = atomic_fetch_add(&j->next, 1);
        if (g >= j->n_regions) break;
        int64_t r0 = j->region_read_start[g], r1 = j->region_read_start[g + 1];
        int64_t h0 = j->region_hap_start[g], h1 = j->region_hap_start[g + 1];
        lko_pairhmm_region((int)(r1 - r0), j->read_off + r0, j->read_bases, j->read_qual, j->read_ins,
                           j->read_del, j->read_gcp, (int)(h1 - h0), j->hap_off + h0, j->hap_bases,
                           j->use_tristate, j->rescaled, j->out + j->out_off[g]);
    }
    return NULL;
}

/* out_off[g] = first element of region g's [reads x haps] block in out. */
void lko_pairhmm_batch_mt(int n_regions, const int64_t *region_read_start, const int64_t *region_hap_start,
                          const int64_t *read_off, const uint8_t *read_bases, const uint8_t *read_qual,
                          const uint8_t *read_ins, const uint8_t *read_del, const uint8_t *read_gcp,
                          const int64_t *hap_off, const uint8_t *hap_bases, const int64_t *out_off,
                          int use_tristate, int rescaled, int n_threads, double *out) {
    lko_init();
    hmm_job_t job = { n_regions, region_read_start, region_hap_start, read_off, hap_off, out_off,
                      read_bases, read_qual, read_ins, read_del, read_gcp, hap_bases,
                      use_tristate, rescaled, out, 0 };
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, hmm_worker, &job);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th);
}

typedef struct {
    int n_pairs;
    const int64_t *ref_off, *alt_off;
    const uint8_t *refs, *alts;
    const int32_t *params;
    int strategy, max_ops, use_shortcut;
    uint32_t *cigar_ops;
    int32_t *n_ops, *offsets;
    atomic_int next;
} sw_job_t;

static void *sw_worker(void *arg) {
    sw_job_t *j = (sw_job_t *)arg;
    for (;;) {
        int p = atomic_fetch_add(&j->next, 1);
        if (p >= j->n_pairs) break;
        int rl = (int)(j->ref_off[p + 1] - j->ref_off[p]), al = (int)(j->alt_off[p + 1] - j->alt_off[p]);
        int (*fn)(const uint8_t *, int, const uint8_t *, int, const int32_t *, int, uint32_t *, int, int32_t *) =
            j->use_shortcut ? lko_sw_align : lko_sw_align_full;
        j->n_ops[p] = fn(j->refs + j->ref_off[p], rl, j->alts + j->alt_off[p], al, j->params, j->strategy,
                         j->cigar_ops + (size_t)p * (size_t)j->max_ops, j->max_ops, &j->offsets[p]);
    }
    return NULL;
}

/* cigar_ops is [n_pairs x max_ops]; n_ops[p] < 0 signals the per-pair error codes of lko_sw_align. */
void lko_sw_batch_mt(int n_pairs, const int64_t *ref_off, const uint8_t *refs, const int64_t *alt_off,
                     const uint8_t *alts, const int32_t params[4], int strategy, int use_shortcut,
                     int max_ops, int n_threads, uint32_t *cigar_ops, int32_t *n_ops, int32_t *offsets) {
    sw_job_t job = { n_pairs, ref_off, alt_off, refs, alts, params, strategy, max_ops, use_shortcut,
                     cigar_ops, n_ops, offsets, 0 };
    if (n_threads < 1) n_threads = 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
    for (int t = 0; t < n_threads; t++) pthread_create(&th[t], NULL, sw_worker, &job);
    for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
    free(th);
}
