/*
 * lorikeet_b200_synth.h — deterministic synthetic workload generators for the configurations
 * BASELINE.json names.  Host-only helpers built into their OWN shared library
 * (liblorikeet_b200_synth.so, no CUDA dependency); used by bench.py (both arms), the tests and
 * __graft_entry__.smoke().  Not part of the drop-in boundary.
 *
 * Every region / pair draws from its own xoshiro256** stream seeded by splitmix64(seed_base + id), so any
 * sub-range of ids can be generated independently (multi-GPU sharding generates only its own shard).
 */
#ifndef LORIKEET_B200_SYNTH_H
#define LORIKEET_B200_SYNTH_H

#include <stdint.h>

#include "lorikeet_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lk_synth_params {
    uint64_t seed_base;        /* cfg#2: 0x4C4B0000, cfg#3: 0x4C4B1000 */
    int32_t n_reads;           /* reads per region                      */
    int32_t read_len;
    int32_t n_haps;            /* haplotypes per region (first = base)  */
    int32_t hap_len;           /* length of the base haplotype          */
    int32_t min_edits, max_edits; /* edits per variant haplotype (used when hap_edit_rate == 0) */
    double hap_edit_rate;      /* if > 0: edits per variant = max(1, round(hap_len * rate))   */
    int32_t qual_profile;      /* 0: {Q2 1%, Q11-25 9%, Q30-40 90%}; 1: uniform Q7..Q20        */
    double read_indel_rate;    /* per-base probability of an insertion or deletion in the read */
    int32_t pcr_model;         /* lk_pcr_model applied to flat Q45 ins/del                      */
    uint8_t base_quality_score_threshold; /* base quals below it become 6 (engine floor)     */
    uint8_t mapq;
} lk_synth_params;

/* Upper bound of haplotype bytes one region can produce. */
int64_t lk_synth_max_hap_bytes_per_region(const lk_synth_params *p);

/*
 * Generate regions [first_region, first_region + n_regions).  Read arenas must hold
 * n_regions*n_reads*read_len bytes; hap_bases must hold n_regions*lk_synth_max_hap_bytes_per_region().
 * Offsets arrays: region_*_start n_regions+1, read_off n_regions*n_reads+1, hap_off n_regions*n_haps+1.
 * The engine pre-ops (PCR model, quality floor) are already applied, i.e. the arrays are what
 * PairHMM::compute_log10_likelihoods receives.  Returns LK_OK / LK_ERR_INVALID.
 */
lk_status lk_synth_pairhmm_regions(const lk_synth_params *p, int64_t first_region, int64_t n_regions,
                                   int64_t *region_read_start, int64_t *region_hap_start, int64_t *read_off,
                                   uint8_t *read_bases, uint8_t *read_qual, uint8_t *read_ins, uint8_t *read_del,
                                   int64_t *hap_off, uint8_t *hap_bases, int n_threads);

/*
 * cfg#4: pairs [first_pair, first_pair + n_pairs), seed 0x4C4B2000 + id: ref = ref_len iid ACGT; alt = ref with
 * 1 % SNPs and 1-3 indels of geometric length (mean 8, max 60); both wrapped in `pad` x 'N'
 * (reads/cigar_utils.rs:10,387-398).  refs needs n_pairs*(ref_len+2*pad) bytes, alts n_pairs*(ref_len+2*pad+180).
 */
lk_status lk_synth_sw_pairs(uint64_t seed_base, int64_t first_pair, int64_t n_pairs, int32_t ref_len, int32_t pad,
                            int64_t *ref_off, uint8_t *refs, int64_t *alt_off, uint8_t *alts, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
