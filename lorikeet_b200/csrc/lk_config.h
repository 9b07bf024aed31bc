// lk_config.h — compile-time configuration shared by host and device code.
#pragma once

#ifndef LK_FAST_THREADS
#define LK_FAST_THREADS 128
#endif
#ifndef LK_K19_MINB
#define LK_K19_MINB 3
#endif
#define LK_PLAIN_THREADS 128
#define LK_PLAIN_K 4
// band path (fp32, reads longer than 32 * max fast K): rows per lane, threads per block, min blocks per SM
#ifndef LK_BAND_K
#define LK_BAND_K 16
#define LK_BAND_MINB 3
#endif
#define LK_BAND_THREADS 128

// Rows per lane (K) the fp32 fast kernel is instantiated for, with the launch-bounds
// "min blocks per SM" that sets its register budget (128 threads/block: 4 -> 128 regs, 3 -> 168,
// 2 -> 255).  The host picks K per read from this list (host_prep.cpp: choose_k).
#define LK_FAST_K_TABLE(X) X(4, 4) X(6, 4) X(8, 4) X(10, 4) X(12, 3) X(14, 3) X(15, 3) X(16, 3) X(19, LK_K19_MINB) X(20, 3)
#define LK_FAST_KS 4, 6, 8, 10, 12, 14, 15, 16, 19, 20
// Smith-Waterman: rows per lane (K) variants; a pair with n reference rows uses the smallest K with 32*K >= n,
// longer references are processed in bands of 32*17 rows.  K >= 8 also exist as packed two-pairs-per-warp kernels.
#define LK_SW_THREADS 128
#define LK_SW_K_TABLE(X) X(2) X(4) X(8) X(12) X(16) X(17)
#define LK_SW_KS 2, 4, 8, 12, 16, 17
