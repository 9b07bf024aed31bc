// host_prep.h — host-side work-list construction for the PairHMM fast path (no CUDA dependency, so the
// tests/emul lock-step emulator can link it too).
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/lorikeet_b200.h"
#include "lk_config.h"
#include "pairhmm_core.cuh"

namespace lk {

struct FastLaunch { int K; int32_t first_task; int32_t n_tasks; bool split; };

struct HostPrep {
    std::vector<int32_t> read_region;
    std::vector<int64_t> read_out_base;
    std::vector<int64_t> henc_off;
    std::vector<WarpTask> tasks;
    std::vector<TaskHaps> task_haps;  // parallel to tasks; read by the kernel only in launches marked split
    std::vector<TaskGroup> groups;
    std::vector<RescuePair> direct;   // pairs that never enter the fast path (read longer than 32*K_max)
    std::vector<FastLaunch> launches;
    int64_t henc_words = 0, n_pairs = 0, n_cells = 0;
    int h_max = 0;
};

const int *fast_k_values(int *n);
int choose_k(int R, int force_k);
// work lists for the region range [g0, g1) of b, with chunk-local indices
// split_target_tasks > 0: a launch with fewer warp tasks than that is cut finer along the haplotype axis (>= 4 haplotypes
// per task) so that a small batch still spreads over the whole GPU
void build_prep(const lk_pairhmm_batch *b, int64_t g0, int64_t g1, int force_k, bool fast_enabled, HostPrep &P,
                int64_t split_target_tasks = 0);

}  // namespace lk
