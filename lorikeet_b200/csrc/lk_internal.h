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
cudaError_t launch_sw(int K, const SwArgs &A, int n_blocks, cudaStream_t st);

}  // namespace lk
