// lk_internal.h — declarations shared between the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lk_config.h"

namespace lk {

struct PairHmmArgs;
struct SwArgs;

cudaError_t launch_encode_haps(const uint8_t *hap_bases, const int64_t *hap_off, const int64_t *henc_off,
                               int64_t n_haps, uint8_t *hidx, uint8_t *hap_flags, cudaStream_t st);
cudaError_t launch_scan_reads(const uint8_t *read_bases, const uint8_t *read_qual, const uint8_t *read_ins,
                              const uint8_t *read_del, const int64_t *read_off, int64_t n_reads, uint8_t *read_flags, cudaStream_t st);
cudaError_t launch_list_all_pairs(const PairHmmArgs &A, int64_t n_reads, cudaStream_t st);
cudaError_t launch_fast(int K, const PairHmmArgs &A, bool vargcp, cudaStream_t st);
cudaError_t launch_band(const PairHmmArgs &A, int n_blocks, cudaStream_t st);
cudaError_t launch_plain(const PairHmmArgs &A, int n_blocks, cudaStream_t st);
// engine pre/post-ops on the device (engine_kernels.cu)
struct EngineDevArgs {
    // reads of the staged chunk (arena pointers biased like PairHmmArgs)
    const uint8_t *read_bases;
    uint8_t *read_qual, *read_ins, *read_del;
    const uint8_t *qual_orig;          // untouched copy of read_qual (dynamic disqualification only)
    const int64_t *read_off;
    const uint8_t *read_mapq;          // per read (chunk-local index)
    const int64_t *read_out_base;      // per read: first pair of (read, haplotype 0)
    const int64_t *block_read_start;   // chunk-local: reads of (region, sample) block b = [block_read_start[b], block_read_start[b+1])
    const int32_t *block_region;       // chunk-local region of block b; null = one block per region (b == region)
    const int64_t *region_hap_start;   // chunk-local
    const int32_t *ref_allele_index;   // per region, may be null
    const double *lk;                  // PairHMM result [r*A + h] per region
    double *values;                    // out: [a*R + r] per region
    uint8_t *keep;                     // out: per read, may be null
    double cap_diff, expected_error_rate, disq_scale;
    int32_t use_pcr, cap_to_mapq, symmetric, dynamic_disq;
    uint8_t bq_threshold;
    uint8_t pcr_cache[101];
};
cudaError_t launch_engine_modify(const EngineDevArgs &E, int64_t n_reads, cudaStream_t st);
cudaError_t launch_engine_epilogue(const EngineDevArgs &E, int64_t n_blocks, cudaStream_t st);
struct BestAlleleBlock;
cudaError_t launch_best_alleles(const double *values, const BestAlleleBlock *blocks, int64_t n_blocks, const int32_t *priorities,
                                double informative_threshold, int32_t *best_allele, uint8_t *informative, cudaStream_t st);
cudaError_t launch_sw(int K, bool clamp, const SwArgs &A, int n_blocks, cudaStream_t st);
bool sw_duo_has_k(int K);
cudaError_t launch_sw_duo(int K, bool same_n, const SwArgs &A, int n_sm, int64_t n_warps, cudaStream_t st);

}  // namespace lk
