// bench_batcher.cpp — developer check of the drop-in call pattern: T host threads each submit ONE region at a time to
// lk_batcher (what Lorikeet's Rayon workers would do through the Rust binding, assembly_region_walker.rs:140-277),
// next to the same regions handed to lk_engine_compute_read_likelihoods in one call.
//   (built by lorikeet_b200/csrc/build.sh into tools/native/bench_batcher)
//   tools/native/bench_batcher [regions=4000] [threads=32] [max_regions=512] [max_wait_us=200] [in_flight=1] [devices=1] [first_device=0] [pinned=0]
// pinned = 1: the caller's arenas and result arrays are page-locked with lk_host_register (what an integrator would do once)
// in_flight > 1: every thread keeps that many regions queued through lk_batcher_submit_async / lk_batcher_wait
// devices: GPUs in the lk_pool (0 = every visible one); ONE process drives them all.  The last line is JSON.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "lorikeet_b200.h"
#include "lorikeet_b200_synth.h"

static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
    const int64_t G = argc > 1 ? atoll(argv[1]) : 4000;
    const int T = argc > 2 ? atoi(argv[2]) : 32;
    const int max_regions = argc > 3 ? atoi(argv[3]) : 512;
    const int max_wait_us = argc > 4 ? atoi(argv[4]) : 200;
    const int depth = argc > 5 ? std::max(1, atoi(argv[5])) : 1;
    const int n_dev_arg = argc > 6 ? atoi(argv[6]) : 1;
    const int first_dev = argc > 7 ? atoi(argv[7]) : 0;
    const int pinned = argc > 8 ? atoi(argv[8]) : 0;
    lk_synth_params p;
    memset(&p, 0, sizeof p);
    p.seed_base = 0x4C4B0000; p.n_reads = 64; p.read_len = 150; p.n_haps = 16; p.hap_len = 300; p.min_edits = 1; p.max_edits = 4;
    p.qual_profile = 0; p.pcr_model = 0 /* raw Q45: the engine applies the PCR model itself */; p.base_quality_score_threshold = 0; p.mapq = 60;
    const int64_t R = G * p.n_reads, H = G * p.n_haps;
    std::vector<int64_t> rrs(G + 1), rhs(G + 1), roff(R + 1), hoff(H + 1);
    std::vector<uint8_t> rb(R * p.read_len), rq(rb.size()), ri(rb.size()), rd(rb.size()), hb(G * lk_synth_max_hap_bytes_per_region(&p));
    if (lk_synth_pairhmm_regions(&p, 0, G, rrs.data(), rhs.data(), roff.data(), rb.data(), rq.data(), ri.data(), rd.data(), hoff.data(), hb.data(), 0) != LK_OK) return 1;
    std::vector<uint8_t> mapq(R, 60);
    lk_pool *pool = nullptr;
    {
        std::vector<int> devs;
        for (int d = 0; d < n_dev_arg; d++) devs.push_back(first_dev + d);
        if (lk_pool_create(n_dev_arg > 0 ? devs.data() : nullptr, n_dev_arg, &pool) != LK_OK) { fprintf(stderr, "no CUDA device / bad device list\n"); return 2; }
    }
    const int n_dev = lk_pool_device_count(pool);
    lk_engine_args args;
    memset(&args, 0, sizeof args);
    args.constant_gcp = 10; args.log10_global_read_mismapping_rate = -4.5; args.pcr_error_model = LK_PCR_CONSERVATIVE;
    args.base_quality_score_threshold = 18; args.read_disqualification_scale = 1.0; args.expected_error_rate_per_base = 0.02;
    args.symmetrically_normalize_alleles_to_reference = 1;
    lk_pairhmm_batch all;
    memset(&all, 0, sizeof all);
    all.n_regions = G; all.region_read_start = rrs.data(); all.region_hap_start = rhs.data(); all.read_off = roff.data();
    all.read_bases = rb.data(); all.read_qual = rq.data(); all.read_ins = ri.data(); all.read_del = rd.data(); all.gcp_constant = 10;
    all.hap_off = hoff.data(); all.hap_bases = hb.data();
    int64_t n_pairs = 0, n_cells = 0;
    lk_pairhmm_batch_dims(&all, &n_pairs, &n_cells);
    std::vector<double> direct(n_pairs), viab(n_pairs);
    std::vector<uint8_t> keep_d(R), keep_b(R);
    if (pinned) {
        int bad_pin = 0;
        bad_pin += lk_host_register(rb.data(), rb.size()) != LK_OK; bad_pin += lk_host_register(rq.data(), rq.size()) != LK_OK;
        bad_pin += lk_host_register(ri.data(), ri.size()) != LK_OK; bad_pin += lk_host_register(rd.data(), rd.size()) != LK_OK;
        bad_pin += lk_host_register(hb.data(), hb.size()) != LK_OK; bad_pin += lk_host_register(mapq.data(), mapq.size()) != LK_OK;
        bad_pin += lk_host_register(direct.data(), direct.size() * 8) != LK_OK; bad_pin += lk_host_register(viab.data(), viab.size() * 8) != LK_OK;
        bad_pin += lk_host_register(keep_d.data(), keep_d.size()) != LK_OK; bad_pin += lk_host_register(keep_b.data(), keep_b.size()) != LK_OK;
        if (bad_pin) { fprintf(stderr, "lk_host_register failed\n"); return 7; }
    }
    // ---- one call ----
    double best = 1e30;
    for (int it = 0; it < 3; it++) {
        const double t0 = now_ms();
        if (lk_pool_compute_read_likelihoods(pool, &args, &all, mapq.data(), nullptr, direct.data(), keep_d.data()) != LK_OK) { fprintf(stderr, "%s\n", lk_pool_last_error(pool)); return 3; }
        best = std::min(best, now_ms() - t0);
    }
    const double one_call_gcups = n_cells / best / 1e6;
    printf("one call (%d device%s)  regions=%lld  %.1f ms  %.0f GCUPS\n", n_dev, n_dev > 1 ? "s" : "", (long long)G, best, one_call_gcups);
    // single-device reference for the bit-for-bit check of the pooled results
    std::vector<double> single(n_pairs);
    std::vector<uint8_t> keep_s(R);
    if (lk_engine_compute_read_likelihoods(lk_pool_ctx(pool, 0), &args, &all, mapq.data(), nullptr, single.data(), keep_s.data()) != LK_OK) return 3;
    const bool pool_same = memcmp(direct.data(), single.data(), n_pairs * 8) == 0 && memcmp(keep_d.data(), keep_s.data(), R) == 0;
    // ---- T threads, one region per submit ----
    lk_batcher *bt = nullptr;
    if (lk_pool_batcher_create(pool, &args, max_regions, max_wait_us, &bt) != LK_OK) return 4;
    std::vector<int64_t> pair_base(G + 1, 0);
    for (int64_t g = 0; g < G; g++) pair_base[g + 1] = pair_base[g] + (rrs[g + 1] - rrs[g]) * (rhs[g + 1] - rhs[g]);
    best = 1e30;
    for (int it = 0; it < 3; it++) {
        std::atomic<int64_t> next{ 0 };
        std::atomic<int> bad{ 0 };
        const double t0 = now_ms();
        std::vector<std::thread> pool;
        for (int t = 0; t < T; t++)
            pool.emplace_back([&]() {
                // `depth` one-region views of the big arrays (offsets rebased to the region, arenas advanced to its first
                // byte), kept alive until their tickets are waited on
                struct Slot { int64_t rs1[2], hs1[2]; std::vector<int64_t> ro, ho; lk_pairhmm_batch one; lk_ticket *t; };
                std::vector<Slot> slots((size_t)depth);
                for (;;) {
                    int used = 0;
                    for (; used < depth; used++) {
                        const int64_t g = next.fetch_add(1);
                        if (g >= G) break;
                        Slot &S = slots[(size_t)used];
                        const int64_t r0 = rrs[g], nr = rrs[g + 1] - r0, h0 = rhs[g], nh = rhs[g + 1] - h0;
                        S.rs1[0] = 0; S.rs1[1] = nr; S.hs1[0] = 0; S.hs1[1] = nh;
                        S.ro.resize(nr + 1); S.ho.resize(nh + 1);
                        for (int64_t i = 0; i <= nr; i++) S.ro[i] = roff[r0 + i] - roff[r0];
                        for (int64_t i = 0; i <= nh; i++) S.ho[i] = hoff[h0 + i] - hoff[h0];
                        memset(&S.one, 0, sizeof S.one);
                        S.one.n_regions = 1; S.one.region_read_start = S.rs1; S.one.region_hap_start = S.hs1; S.one.read_off = S.ro.data();
                        S.one.read_bases = rb.data() + roff[r0]; S.one.read_qual = rq.data() + roff[r0]; S.one.read_ins = ri.data() + roff[r0];
                        S.one.read_del = rd.data() + roff[r0]; S.one.gcp_constant = 10; S.one.hap_off = S.ho.data(); S.one.hap_bases = hb.data() + hoff[h0];
                        S.t = nullptr;
                        if (lk_batcher_submit_async(bt, &S.one, mapq.data() + r0, 0, viab.data() + pair_base[g], keep_b.data() + r0, &S.t) != LK_OK) bad++;
                    }
                    for (int i = 0; i < used; i++)
                        if (slots[(size_t)i].t && lk_batcher_wait(bt, slots[(size_t)i].t) != LK_OK) bad++;
                    if (used < depth) break;
                }
            });
        for (auto &th : pool) th.join();
        if (bad) { fprintf(stderr, "submit failed\n"); return 5; }
        best = std::min(best, now_ms() - t0);
    }
    int64_t nb = 0, nr = 0;
    lk_batcher_stats(bt, &nb, &nr);
    printf("batcher %2d threads x %d in flight  regions=%lld  %.1f ms  %.0f GCUPS  (%.1f regions per launch)\n", T, depth, (long long)G, best, n_cells / best / 1e6, (double)nr / (double)nb);
    const bool same = pool_same && memcmp(single.data(), viab.data(), n_pairs * 8) == 0 && memcmp(keep_s.data(), keep_b.data(), R) == 0;
    printf("pooled one-call and batcher results %s the single-device one-call results\n", same ? "are bit-identical to" : "DIFFER from");
    printf("{\"devices\": %d, \"threads\": %d, \"in_flight\": %d, \"regions\": %lld, \"one_call_gcups\": %.1f, \"batcher_gcups\": %.1f, "
           "\"regions_per_launch\": %.1f, \"bit_identical\": %s, \"pinned\": %d}\n", n_dev, T, depth, (long long)G, one_call_gcups, n_cells / best / 1e6,
           (double)nr / (double)nb, same ? "true" : "false", pinned);
    lk_batcher_destroy(bt);
    lk_pool_destroy(pool);
    return same ? 0 : 6;
}
