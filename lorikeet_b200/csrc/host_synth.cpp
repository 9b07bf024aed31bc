// host_synth.cpp — deterministic synthetic workloads (see include/lorikeet_b200_synth.h).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/lorikeet_b200_synth.h"

namespace {

struct Rng {  // splitmix64 -> xoshiro256**
    uint64_t s[4];
    static uint64_t splitmix(uint64_t &x) {
        uint64_t z = (x += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    explicit Rng(uint64_t seed) { for (auto &v : s) v = splitmix(seed); }
    static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    uint64_t next() {
        const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
        return r;
    }
    double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    int below(int n) { return (int)(uniform() * n); }            // [0, n)
    int range(int lo, int hi) { return lo + below(hi - lo + 1); } // [lo, hi]
    uint8_t base() { return "ACGT"[next() >> 62]; }
    uint8_t other_base(uint8_t b) { uint8_t c; do c = base(); while (c == b); return c; }
};

void edit_haplotype(Rng &rng, std::vector<uint8_t> &h, int n_edits) {
    for (int e = 0; e < n_edits; e++) {
        const double u = rng.uniform();
        if (h.size() < 32) return;
        const int pos = rng.below((int)h.size() - 20) + 10;
        if (u < 0.70) {
            h[(size_t)pos] = rng.other_base(h[(size_t)pos]);
        } else if (u < 0.85) {
            const int len = std::min<int>(rng.range(1, 10), (int)h.size() - pos - 5);
            h.erase(h.begin() + pos, h.begin() + pos + len);
        } else {
            const int len = rng.range(1, 10);
            std::vector<uint8_t> ins((size_t)len);
            for (auto &b : ins) b = rng.base();
            h.insert(h.begin() + pos, ins.begin(), ins.end());
        }
    }
}

uint8_t draw_qual(Rng &rng, int profile) {
    if (profile == 1) return (uint8_t)rng.range(7, 20);
    const double u = rng.uniform();
    if (u < 0.01) return 2;
    if (u < 0.10) return (uint8_t)rng.range(11, 25);
    return (uint8_t)rng.range(30, 40);
}

template <class F>
void parallel_for(int64_t n, int n_threads, F f) {
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    if (n_threads == 1 || n < 2) { for (int64_t i = 0; i < n; i++) f(i); return; }
    std::atomic<int64_t> next{ 0 };
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++)
        th.emplace_back([&]() { for (int64_t i; (i = next.fetch_add(1)) < n;) f(i); });
    for (auto &t : th) t.join();
}

int edits_per_variant_max(const lk_synth_params *p) {
    if (p->hap_edit_rate > 0) return std::max(1, (int)std::lround(p->hap_len * p->hap_edit_rate));
    return p->max_edits;
}

}  // namespace

extern "C" int64_t lk_synth_max_hap_bytes_per_region(const lk_synth_params *p) {
    if (!p) return 0;
    return (int64_t)p->n_haps * ((int64_t)p->hap_len + 10LL * edits_per_variant_max(p));
}

extern "C" lk_status lk_synth_pairhmm_regions(const lk_synth_params *p, int64_t first_region, int64_t n_regions,
                                              int64_t *region_read_start, int64_t *region_hap_start, int64_t *read_off,
                                              uint8_t *read_bases, uint8_t *read_qual, uint8_t *read_ins, uint8_t *read_del,
                                              int64_t *hap_off, uint8_t *hap_bases, int n_threads) {
    if (!p || n_regions < 0 || p->n_reads < 1 || p->n_haps < 1 || p->read_len < 1 || p->hap_len < p->read_len + 64)
        return LK_ERR_INVALID;
    if (!region_read_start || !region_hap_start || !read_off || !read_bases || !read_qual || !read_ins || !read_del ||
        !hap_off || !hap_bases)
        return LK_ERR_INVALID;
    const int64_t slot = lk_synth_max_hap_bytes_per_region(p);
    std::vector<std::vector<int32_t>> hap_lens((size_t)n_regions);
    parallel_for(n_regions, n_threads, [&](int64_t g) {
        Rng rng(p->seed_base + (uint64_t)(first_region + g));
        std::vector<std::vector<uint8_t>> haps((size_t)p->n_haps);
        haps[0].resize((size_t)p->hap_len);
        for (auto &b : haps[0]) b = rng.base();
        for (int h = 1; h < p->n_haps; h++) {
            haps[(size_t)h] = haps[0];
            const int n_edits = p->hap_edit_rate > 0 ? edits_per_variant_max(p) : rng.range(p->min_edits, p->max_edits);
            edit_haplotype(rng, haps[(size_t)h], n_edits);
        }
        uint8_t *hdst = hap_bases + g * slot;  // compacted afterwards
        for (int h = 0; h < p->n_haps; h++) {
            memcpy(hdst, haps[(size_t)h].data(), haps[(size_t)h].size());
            hdst += haps[(size_t)h].size();
            hap_lens[(size_t)g].push_back((int32_t)haps[(size_t)h].size());
        }
        const int R = p->read_len;
        for (int r = 0; r < p->n_reads; r++) {
            const int64_t o = (g * p->n_reads + r) * (int64_t)R;
            const auto &src = haps[(size_t)rng.below(p->n_haps)];
            const int slack = (int)src.size() - R - (int)(R * p->read_indel_rate * 2) - 8;
            int pos = rng.below(std::max(1, slack));
            for (int i = 0; i < R; i++) {
                const uint8_t q = draw_qual(rng, p->qual_profile);
                uint8_t b;
                if (p->read_indel_rate > 0) {
                    const double u = rng.uniform();
                    if (u < p->read_indel_rate * 0.5) {
                        b = rng.base();                      // inserted base, haplotype position unchanged
                    } else {
                        if (u < p->read_indel_rate) pos++;   // deletion: skip one haplotype base
                        b = src[(size_t)std::min<int>(pos++, (int)src.size() - 1)];
                    }
                } else {
                    b = src[(size_t)pos++];
                }
                if (rng.uniform() < std::pow(10.0, -(double)q / 10.0)) b = rng.other_base(b);
                read_bases[o + i] = b;
                read_qual[o + i] = q;
                read_ins[o + i] = 45;  // ReadUtils default when BI/BD are absent (read_utils.rs:23,372-416)
                read_del[o + i] = 45;
            }
        }
    });
    // compact haplotypes and write offsets
    region_read_start[0] = 0; region_hap_start[0] = 0; read_off[0] = 0; hap_off[0] = 0;
    int64_t w = 0;
    for (int64_t g = 0; g < n_regions; g++) {
        region_read_start[g + 1] = (g + 1) * p->n_reads;
        region_hap_start[g + 1] = (g + 1) * p->n_haps;
        int64_t len = 0;
        for (int h = 0; h < p->n_haps; h++) {
            len += hap_lens[(size_t)g][(size_t)h];
            hap_off[g * p->n_haps + h + 1] = w + len;
        }
        if (w != g * slot) memmove(hap_bases + w, hap_bases + g * slot, (size_t)len);
        w += len;
    }
    const int64_t n_reads = n_regions * p->n_reads;
    for (int64_t r = 0; r < n_reads; r++) read_off[r + 1] = (r + 1) * (int64_t)p->read_len;
    // engine pre-ops (apply_pcr_error_model + cap_minimum_read_qualities), as compute_read_likelihoods would
    lk_engine_args ea;
    memset(&ea, 0, sizeof ea);
    ea.pcr_error_model = p->pcr_model;
    ea.base_quality_score_threshold = p->base_quality_score_threshold;
    std::vector<uint8_t> mapq((size_t)n_reads, p->mapq);
    return lk_engine_modify_read_qualities(&ea, n_reads, read_off, read_bases, read_qual, read_ins, read_del, mapq.data(),
                                           n_threads);
}

extern "C" lk_status lk_synth_sw_pairs(uint64_t seed_base, int64_t first_pair, int64_t n_pairs, int32_t ref_len,
                                       int32_t pad, int64_t *ref_off, uint8_t *refs, int64_t *alt_off, uint8_t *alts,
                                       int n_threads) {
    if (n_pairs < 0 || ref_len < 64 || pad < 0 || !ref_off || !refs || !alt_off || !alts) return LK_ERR_INVALID;
    const int64_t rslot = ref_len + 2 * pad, aslot = ref_len + 2 * pad + 180;
    std::vector<int32_t> alt_len((size_t)n_pairs);
    parallel_for(n_pairs, n_threads, [&](int64_t i) {
        Rng rng(seed_base + (uint64_t)(first_pair + i));
        std::vector<uint8_t> ref((size_t)ref_len);
        for (auto &b : ref) b = rng.base();
        std::vector<uint8_t> alt = ref;
        for (auto &b : alt)
            if (rng.uniform() < 0.01) b = rng.other_base(b);
        const int n_indels = rng.range(1, 3);
        for (int k = 0; k < n_indels; k++) {
            int len = 1;
            while (len < 60 && rng.uniform() > 1.0 / 8.0) len++;  // geometric, mean 8
            const int pos = rng.below((int)alt.size() - 80) + 20;
            if (rng.uniform() < 0.5) {
                alt.erase(alt.begin() + pos, alt.begin() + pos + std::min<int>(len, (int)alt.size() - pos - 20));
            } else {
                std::vector<uint8_t> ins((size_t)len);
                for (auto &b : ins) b = rng.base();
                alt.insert(alt.begin() + pos, ins.begin(), ins.end());
            }
        }
        uint8_t *r = refs + i * rslot, *a = alts + i * aslot;
        memset(r, 'N', (size_t)pad); memcpy(r + pad, ref.data(), ref.size()); memset(r + pad + ref.size(), 'N', (size_t)pad);
        memset(a, 'N', (size_t)pad); memcpy(a + pad, alt.data(), alt.size()); memset(a + pad + alt.size(), 'N', (size_t)pad);
        alt_len[(size_t)i] = (int32_t)alt.size() + 2 * pad;
    });
    ref_off[0] = 0; alt_off[0] = 0;
    int64_t w = 0;
    for (int64_t i = 0; i < n_pairs; i++) {
        ref_off[i + 1] = (i + 1) * rslot;
        if (w != i * aslot) memmove(alts + w, alts + i * aslot, (size_t)alt_len[(size_t)i]);
        w += alt_len[(size_t)i];
        alt_off[i + 1] = w;
    }
    return LK_OK;
}
