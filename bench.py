#!/usr/bin/env python3
"""bench.py — PairHMM GCUPS on BASELINE.json's cfg#2 (synthetic active-region batch).

    python bench.py --gpus N --steps K --warmup W            # our arm (CUDA, one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...   # reference arm: the CPU restatement of the reference's
                                                              # scalar path on the host cores (rank 0 only)

A "step" = one pass of the hot path (PairHMM::compute_likelihoods for every region of the batch) over one synthetic
batch of `--regions` active regions (default 10 000, the configuration the metric is quoted on).
  value : whole-job GCUPS (read-bp x hap-bp / s / 1e9) with the batch resident in HBM, device-timed (CUDA events).
  e2e   : the same metric through the reference-facing C-ABI call lk_engine_compute_read_likelihoods (the region driver
          compute_read_likelihoods) with HOST (pinned) buffers: H2D of raw reads + pre-ops + PairHMM + post-ops + D2H inside
          the timed region; e2e.pairhmm_only is the bare lk_pairhmm_batch_host call.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "PairHMM GCUPS (read-bp x hap-bp / s)"
UNIT = "GCUPS"
FLOP_PER_CELL = 12  # pair_hmm.rs:575-591: M = 4 mul + 2 add, I = 2 mul + 1 add, D = 2 mul + 1 add
NCU_DRAM_BYTES_PER_REGION = 47.3e3  # ncu: 46.9 MB read + 0.4 MB written by pairhmm_fast_kernel over 1000 regions (profiles/r01_pairhmm_fast_k19_seltab_ncu_details.txt)


def workload_name(regions):
    return "cfg2 synthetic active-region batch: %d regions x (64 reads x 150 bp, 16 haps x ~300 bp), fp32 scaled-linear" % regions


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for s in self.samples:
            t = [x.strip() for x in s.split(",")]
            if len(t) < 7:
                continue
            try:
                sm.append(float(t[0])); mx.append(float(t[1])); pw.append(float(t[2]))
            except ValueError:
                continue
            for name, v in zip(self.NAMES, t[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # "under load" = samples drawing clearly more than idle power
        load = [c for c, p in zip(sm, pw) if p > 250.0] or sm
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(pw) if pw else None}


def cpu_baseline_run(batch, n_regions_sample, native=True):
    """Time the oracle (C restatement of the reference's scalar f64 path) on a bounded sample with all host cores.
    One region per worker task = the granularity Rayon gives the reference (assembly_region_walker.rs:140)."""
    from oracle import lk_oracle as O
    lib_path = None
    if native:
        try:
            lib_path = O.build(native=True)  # -march=native copy for the host we are on
        except Exception:
            lib_path = None
    sample = batch.region_slice(0, min(n_regions_sample, batch.n_regions))
    cores = os.cpu_count() or 1
    t0 = time.perf_counter()
    O.pairhmm_batch(sample, n_threads=cores, lib_path=lib_path)
    dt = time.perf_counter() - t0
    return sample.n_cells / dt / 1e9, cores, dt, sample


def run_extras(ctx, args, peak_fp32):
    """Side measurements of BASELINE.json's other single-GPU configs (kernel-resident; NOT the bench metric), each with
    its own roofline block: cfg#4 Smith-Waterman (100 000 pairs) and cfg#3 long reads at its FULL size (2 000 regions)."""
    from lorikeet_b200 import synth
    import numpy as np
    import torch
    out = {}
    try:
        ro, refs, ao, alts = synth.sw_pairs(100000)  # cfg4: 100k pairs, 500 bp + 10xN pads, NEW_SW_PARAMETERS, SoftClip
        res = ctx.sw_upload(ro, refs, ao, alts, [200, -150, -260, -11], "SoftClip", max_ops=64)
        res.run()
        ms = min(res.run() for _ in range(3))
        n_cells = res.n_cells
        gcups = n_cells / ms / 1e6
        n_sm = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        sm_hz = (measured_peaks() or {}).get("sm_max_mhz", 1965.0) * 1e6
        issue_peak = n_sm * 4 * 32 * sm_hz          # thread-instructions / s at one warp instruction per cycle and scheduler
        int32_peak = n_sm * 64 * sm_hz              # 64 int32 lanes per SM (the 16-lane ALU pipe x 4)
        out["cfg4_smith_waterman"] = {
            "pairs": 100000, "cells": n_cells, "ms": ms, "gcups": gcups, "packed_pairs": res.packed_pairs,
            "note": "fill + in-kernel traceback, CIGARs bit-exact vs oracle (tests/test_gpu_sw.py, test_gpu_fullsize.py); "
                    "packed_pairs = pairs that went two to a warp through the 16-bit difference kernel (sw16_core.cuh)",
            "roofline": {
                "bound": "integer issue (ALU and FMA-heavy pipes, 16 lanes each, balanced by construction)", "unit": "fraction",
                "instructions_per_cell": SW_NCU["instr_per_cell"], "instructions_source": SW_NCU["source"],
                "issue_slot_frac": gcups * 1e9 * SW_NCU["instr_per_cell"] / issue_peak,
                "int32_frac_at_15_ops_per_cell": gcups * 1e9 * 15.0 / int32_peak,
                "int32_frac_note": "against 64 32-bit integer lanes per SM at the reference's 15 ops per cell; above 1 because the "
                                   "packed kernel serves two cells with every 32-bit lane operation",
                "dram_bytes_per_cell": SW_NCU["dram_bytes_per_cell"],
                "hbm_frac": gcups * 1e9 * SW_NCU["dram_bytes_per_cell"] / ((measured_peaks() or {}).get("hbm_gbs", 6550.0) * 1e9),
                "algorithmic_bytes_per_cell": float(refs.nbytes + alts.nbytes + 100000 * (64 * 4 + 8)) / n_cells,
            },
        }
        res.free()
        # the same batch end to end through lk_sw_batch_host: pinned host buffers in, CIGARs + offsets in pinned host buffers out
        pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
        pro, prefs, pao, palts = pin(ro), pin(refs), pin(ao), pin(alts)
        outs = (torch.empty((100000, 64), dtype=torch.int32).pin_memory().numpy().view(np.uint32),
                torch.empty(100000, dtype=torch.int32).pin_memory().numpy(), torch.empty(100000, dtype=torch.int32).pin_memory().numpy())
        ctx.sw_batch(pro, prefs, pao, palts, [200, -150, -260, -11], "SoftClip", max_ops=64, out=outs)
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            ctx.sw_batch(pro, prefs, pao, palts, [200, -150, -260, -11], "SoftClip", max_ops=64, out=outs)
            best = min(best, time.perf_counter() - t0)
        out["cfg4_smith_waterman"]["e2e"] = {"ms": best * 1e3, "gcups": n_cells / best / 1e9,
                                             "api": "lk_sw_batch_host (the gkl::smithwaterman::align seam, batched), pinned host buffers"}
        if not args.no_cpu_baseline:
            # the reference's scalar SW (C restatement, the checker of the parity tests) on the first 4 000 pairs, all host cores
            from oracle import lk_oracle as O
            ns = 4000
            r_, a_ = ro[:ns + 1], ao[:ns + 1]
            cores = os.cpu_count() or 1
            t0 = time.perf_counter()
            O.sw_batch(r_, refs[:r_[-1]], a_, alts[:a_[-1]], [200, -150, -260, -11], "SoftClip", max_ops=64, n_threads=cores)
            dt = time.perf_counter() - t0
            sc = float((np.diff(r_) * np.diff(a_)).sum())
            out["cfg4_smith_waterman"]["cpu_baseline"] = {
                "value": sc / dt / 1e9, "unit": "GCUPS", "cores": cores, "kind": "port",
                "sample": "first %d pairs of the same generator (%.3e cells, %.2f s)" % (ns, sc, dt),
                "note": "C restatement of the reference's in-tree scalar path (smith_waterman_aligner.rs:124-442); gkl's AVX path is not vendored"}
    except Exception as e:  # side measurements must never break the bench line
        out["cfg4_error"] = str(e)
    try:
        # the rate Lorikeet's call pattern sees: Rayon-like host threads submitting ONE region each to lk_batcher (native
        # harness tools/native/bench_batcher.cpp, one process, this GPU only; multi-device pool runs are in profiles/)
        dev = str(torch.cuda.current_device())
        tool = os.path.join(ROOT, "tools", "native", "bench_batcher")
        runs = []
        for threads, in_flight in ((64, 1), (256, 1), (16, 16)):
            p = subprocess.run([tool, "4000", str(threads), "512", "200", str(in_flight), "1", dev], capture_output=True, text=True, timeout=300)
            last = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
            runs.append(json.loads(last[-1]) if last else {"error": (p.stderr or p.stdout)[-300:], "threads": threads})
        out["batcher"] = {"api": "lk_pool_batcher_create + lk_batcher_submit[_async] (one cfg2 region per submit, pageable buffers)",
                          "runs": runs}
    except Exception as e:
        out["batcher_error"] = str(e)
    try:
        # cfg3 at BASELINE's full size: 2 000 regions x (8 reads x 10 kbp, 4 haps x 12 kbp) = 7.68e12 cells
        n3 = args.cfg3_regions
        b = synth.pairhmm_regions(n3, **synth.CFG3)
        res = ctx.pairhmm_upload(b)
        res.run()
        runs = []
        for _ in range(2):
            ms = res.run()
            runs.append((ms, res.last_fast_ms))
        ms, band_ms = min(runs)
        _, flags = res.download(want_flags=True)
        achieved = FLOP_PER_CELL * res.n_cells / (band_ms * 1e-3) / 1e12
        out["cfg3_long_reads"] = {
            "regions": n3, "pairs": res.n_pairs, "cells": res.n_cells, "ms": ms, "gcups": res.n_cells / ms / 1e6,
            "pairs_redone_in_f64": int((flags & 1).sum()),
            "note": "fp32 band kernel (pairhmm_band_kernel) with per-lane block floating point; full BASELINE size when regions = 2000",
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_fp32, "unit": "TFLOP/s", "frac": achieved / peak_fp32,
                         "kernel": "pairhmm_band_kernel", "kernel_ms": band_ms, "algorithmic_flop_per_cell": FLOP_PER_CELL,
                         "traffic": None},
        }
        res.free()
        del b
    except Exception as e:
        out["cfg3_error"] = str(e)
    return out


# figures read from the committed ncu capture of the SW kernel (never measured inside the bench)
SW_NCU = {"instr_per_cell": 12.2, "dram_bytes_per_cell": 0.77,
          "source": "profiles/r02_sw_duo_kernel_k17_metrics.txt (packed two-pairs-per-warp kernel, 20 000 cfg4 pairs: 6.58e10 thread "
                    "instructions in 2.28e9 warp instructions, 2.93 GB written + 1.22 GB read over 5.41e9 cells; the one-pair "
                    "kernel it replaces on this workload: 17.9 instructions and 0.80 bytes per cell)"}


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm (oracle port; the Rust original cannot be built here).
    --ref-kind scalar (default): the C restatement of the in-tree scalar f64 path.  --ref-kind vector: OUR stand-in for the
    un-vendored gkl AVX path (fp32 anti-diagonal SIMD kernel under oracle/, runtime-dispatched ISA)."""
    if rank != 0:
        return
    from lorikeet_b200 import synth  # generator library only: the product library is not mapped by this arm
    regions = args.ref_sample_regions
    batch = synth.pairhmm_regions(regions)
    from oracle import lk_oracle as O
    try:
        lib_path = O.build(native=True)
    except Exception:
        lib_path = None
    cores = os.cpu_count() or 1
    vector = args.ref_kind == "vector"
    run = (lambda b: O.pairhmm_batch_vector(b, n_threads=cores, lib_path=lib_path)) if vector else \
          (lambda b: O.pairhmm_batch(b, n_threads=cores, lib_path=lib_path))
    for _ in range(args.warmup):
        run(batch.region_slice(0, max(1, regions // 8)))
    t = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        run(batch)
        t.append(time.perf_counter() - t0)
    total = sum(t)
    value = batch.n_cells * args.steps / total / 1e9
    sample = "%d regions of the cfg2 generator (%.3e cells) per step" % (regions, batch.n_cells)
    kind_note = ("OUR fp32 SIMD stand-in for the un-vendored gkl AVX path (%s), not gkl itself" % O.vector_isa(lib_path)) if vector else \
                "C restatement of the reference's in-tree SCALAR f64 path, not the gkl AVX path"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "vs_baseline_note": "no published number; same_config false: %s; %s" % (sample, kind_note),
        "dtype": "f32" if vector else "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.regions), "sample": sample,
                   "note": "reference is Rust and cannot be built here; " + kind_note + " (oracle/), one region per thread "
                           "as Rayon schedules it"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "ref_kind": args.ref_kind},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else written to fd 1 during the run (NCCL's version banner,
    library chatter) has been diverted to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


class ShardedBench:
    """One sharded configuration measured both ways on this rank's share of the regions:
      value : kernels on the HBM-resident shard + THE collective (all_gather of the result slabs), device-timed;
      e2e   : raw reads in pinned HOST memory -> lk_engine_compute_read_likelihoods (H2D, device pre-ops, PairHMM, post-ops)
              -> result slab in HBM -> all_gather -> the full gathered matrix set copied to pinned HOST memory, wall-timed.
    Region batches are dealt round-robin (lorikeet_b200.dispatch.ShardedRegions); after the timed loops every rank
    recomputes foreign batches locally and compares them bit for bit with what it received (gather_verified)."""

    def __init__(self, ctx, n_regions, batch_regions, rank, world, local_rank, gather_to_host=False):
        import numpy as np
        import torch
        from lorikeet_b200 import synth
        from lorikeet_b200.dispatch import ShardedRegions, device_view
        from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine, PCRErrorModel
        self.np, self.torch, self.synth, self.device_view = np, torch, synth, device_view
        self.ctx, self.rank, self.world = ctx, rank, world
        self.gather_to_host = gather_to_host
        self.eng = PairHMMLikelihoodCalculationEngine(ctx, 10, -4.5, PCRErrorModel.CONSERVATIVE, 18, False, 1.0, 0.02, True, False, False)
        self.sh = ShardedRegions(n_regions, batch_regions)
        self.n_regions, self.batch_regions = n_regions, batch_regions
        # RAW reads of this rank's batches (flat Q45 indel quals, un-floored base quals): what the region driver takes
        self.raw = self.sh.build(self.gen_raw)
        # the same reads after modify_read_qualities (device launch, byte-identical to the host utility): what the bare
        # PairHMM seam takes; staged in HBM once for the kernel-resident measurement
        pre = self.preprocessed(self.raw)
        self.cells = pre.n_cells
        self.lay = self.sh.agree_layout(pre, device="cuda")
        self.res = ctx.pairhmm_upload(pre)
        self.n_mine = self.lay.sizes[self.rank]
        assert self.n_mine == self.res.n_pairs
        self.input_bytes = self.raw.input_bytes()
        del pre
        self.equal_slabs = len(set(self.lay.sizes)) == 1
        stride = self.lay.stride
        self.local_out = device_view(self.res.out_device_ptr, self.n_mine) if self.n_mine else torch.zeros(0, dtype=torch.float64, device="cuda")
        if world > 1:
            self.slab = self.local_out if self.equal_slabs else torch.zeros(stride, dtype=torch.float64, device="cuda")
            self.gathered = torch.empty(world * stride, dtype=torch.float64, device="cuda")
        else:
            self.slab, self.gathered = self.local_out, self.local_out

    def gen_raw(self, g0, g1):
        return self.synth.pairhmm_regions(g1 - g0, first_region=g0, pcr_model=0, base_quality_score_threshold=0)

    def preprocessed(self, raw):
        from lorikeet_b200 import PairHmmBatch
        pre = PairHmmBatch(raw.region_read_start, raw.region_hap_start, raw.read_off, raw.read_bases, raw.read_qual.copy(),
                           raw.read_ins.copy(), raw.read_del.copy(), raw.hap_off, raw.hap_bases, read_gcp=None, gcp_constant=10)
        return self.eng.modify_read_qualities_device(pre)

    # ---- kernel-resident step: kernels + collective, device-timed ----
    def step(self):
        torch = self.torch
        ms = self.res.run()  # prologue + fp32 wavefront + f64 pass; CUDA events on the launching stream
        fast_ms = self.res.last_fast_ms
        if self.world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if not self.equal_slabs and self.n_mine:
                self.slab[:self.n_mine].copy_(self.local_out)
            self.sh.all_gather(self.slab, self.gathered)  # the one collective of the path: result slabs
            e1.record()
            e1.synchronize()
            ms += e0.elapsed_time(e1)
        return ms, fast_ms

    def verify(self, gathered_host_fn, compute_fn, n_checks=2, equal_nan=False):
        """Recompute `n_checks` random FOREIGN batches locally and compare bit for bit with the gathered copy."""
        np = self.np
        rng = np.random.default_rng(1234 + self.rank)
        ok, checked = True, 0
        others = [r for r in range(self.world) if r != self.rank and self.lay.plan[r]] or [self.rank]
        for _ in range(n_checks):
            r = int(rng.choice(others))
            i = int(rng.integers(0, len(self.lay.plan[r])))
            g0, g1 = self.lay.plan[r][i]
            pos, n = self.lay.position(r, i)
            exp = compute_fn(self.gen_raw(g0, g1))
            got = gathered_host_fn(pos, n)
            ok = ok and exp.size == n and bool(np.array_equal(got, exp, equal_nan=equal_nan))
            checked += 1
        return ok, checked

    def verify_resident(self):
        return self.verify(lambda pos, n: self.gathered[pos:pos + n].cpu().numpy(),
                           lambda raw: self.ctx.pairhmm_batch(self.preprocessed(raw)))

    # ---- end to end ----
    def setup_e2e(self):
        np, torch = self.np, self.torch
        import lorikeet_b200 as lk

        def pin(a):
            t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
            return t, t.numpy()
        self._keepalive, rp = [], {}
        for k in ("read_bases", "read_qual", "read_ins", "read_del", "hap_bases", "read_off", "hap_off", "region_read_start", "region_hap_start"):
            t, a = pin(getattr(self.raw, k))
            self._keepalive.append(t)
            rp[k] = a
        self.rbatch = lk.PairHmmBatch(rp["region_read_start"], rp["region_hap_start"], rp["read_off"], rp["read_bases"], rp["read_qual"],
                                      rp["read_ins"], rp["read_del"], rp["hap_off"], rp["hap_bases"], read_gcp=None, gcp_constant=10)
        n_reads = self.raw.n_reads
        self.raw = None  # the pinned copy is the one the timed calls read
        t, self.mapq = pin(np.full(n_reads, 60, np.uint8))
        self._keepalive.append(t)
        self.keep_t = torch.empty(n_reads, dtype=torch.uint8).pin_memory()
        stride = self.lay.stride
        # e2e at N > 1: the region driver leaves this rank's matrices in HBM (values may be device memory), then -- on a side
        # stream, double-buffered, overlapping the NEXT step's H2D and kernels -- the slab is (a) copied to this rank's pinned
        # host memory (the step's result read, as at N = 1) and (b) all_gathered, so that every GPU holds the full likelihood set
        # (what north_star's allgather is for).  gather_to_host additionally copies the full gathered set (N x the slab) to every
        # rank's host: measured once at N = 8 (profiles/r02_bench_n8_full_set_to_host.json), eight ranks writing the same 656 MB
        # to host memory per step cost 48 ms of a 155 ms step even overlapped.  The timed loop ends with e2e_finish().
        full = self.world * stride if self.gather_to_host else max(self.n_mine, 1)
        self.host_bufs = [torch.empty(full, dtype=torch.float64).pin_memory() for _ in range(2 if self.world > 1 else 1)]
        self.host_all = self.host_bufs[0]
        self.e_it = 0
        if self.world > 1:
            self.e_slab = [torch.zeros(stride, dtype=torch.float64, device="cuda") for _ in range(2)]
            self.e_gathered = [torch.empty(self.world * stride, dtype=torch.float64, device="cuda") for _ in range(2)]
            self.e_side = torch.cuda.Stream()
            self.e_done = [None, None]
            self.e_last = 0
        self.h2d = self.input_bytes + int(self.mapq.nbytes)
        self.d2h = int(self.host_all.numel() * 8 + self.keep_t.numel())

    def e2e_step(self):
        torch = self.torch
        keep = self.keep_t.numpy()
        if self.world == 1:  # matrices straight into pinned host memory
            self.eng.compute_read_likelihoods(self.rbatch, self.mapq, values=self.host_all.numpy()[:self.n_mine], keep=keep)
            return
        b = self.e_it & 1
        self.e_it += 1
        if self.e_done[b] is not None:
            self.e_done[b].synchronize()  # buffer pair b was last used two steps ago: its gather + copy must be through
        # matrices stay in HBM for the collective (the call returns when this rank's slab is complete)
        self.eng.compute_read_likelihoods(self.rbatch, self.mapq, keep=keep, values_device_ptr=self.e_slab[b].data_ptr())
        with torch.cuda.stream(self.e_side):
            if not self.gather_to_host and self.n_mine:
                self.host_bufs[b][:self.n_mine].copy_(self.e_slab[b][:self.n_mine], non_blocking=True)
            self.sh.all_gather(self.e_slab[b], self.e_gathered[b])
            if self.gather_to_host:
                self.host_bufs[b].copy_(self.e_gathered[b], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.e_side)
        self.e_done[b] = ev
        self.e_last = b
        self.host_all = self.host_bufs[b]

    def e2e_finish(self):
        if self.world > 1:
            for ev in self.e_done:
                if ev is not None:
                    ev.synchronize()

    def verify_e2e(self):
        np = self.np
        mq = lambda raw: np.full(raw.n_reads, 60, np.uint8)
        got = (lambda pos, n: self.host_all.numpy()[pos:pos + n]) if self.gather_to_host else \
              (lambda pos, n: self.e_gathered[self.e_last][pos:pos + n].cpu().numpy())
        ok, n = self.verify(got, lambda raw: self.eng.compute_read_likelihoods(raw, mq(raw))[0], n_checks=1, equal_nan=True)
        if not self.gather_to_host and self.n_mine:  # this rank's host copy equals its slab of the gathered set
            mine = self.e_gathered[self.e_last][self.rank * self.lay.stride:self.rank * self.lay.stride + self.n_mine].cpu().numpy()
            ok = ok and bool(np.array_equal(self.host_all.numpy()[:self.n_mine], mine, equal_nan=True))
        return ok, n

    def close(self):
        self.res.free()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg2", "cfg5"],
                    help="cfg2 (default): weak scaling, --regions per GPU, the line the metric is quoted on; cfg5: BASELINE's "
                         "scale-out config as THE line: 200 000 regions dealt in 256-region batches, strong scaling")
    ap.add_argument("--regions", type=int, default=10000, help="regions per GPU (cfg2 = 10000)")
    ap.add_argument("--cfg5-regions", type=int, default=200000)
    ap.add_argument("--cfg5-steps", type=int, default=2, help="timed steps of the cfg5 side measurement in a default run")
    ap.add_argument("--cfg3-regions", type=int, default=2000, help="cfg3 side measurement (2000 = BASELINE's full size)")
    ap.add_argument("--ref-sample-regions", type=int, default=192, help="regions per step of the reference arm")
    ap.add_argument("--ref-kind", default="scalar", choices=["scalar", "vector"])
    ap.add_argument("--cpu-sample-regions", type=int, default=256, help="regions of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the SW (cfg4), long-read (cfg3) and scale-out (cfg5) side measurements")
    ap.add_argument("--no-cfg5", action="store_true")
    ap.add_argument("--e2e-gather-to-host", action="store_true",
                    help="N > 1: e2e also copies the FULL gathered set (N x the slab) to every rank's host memory each step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import lorikeet_b200 as lk

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(*vals):
        if world == 1:
            return [float(v) for v in vals]
        t = torch.tensor(vals, dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t]

    def reduce_sum(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    def all_true(flag):
        if world == 1:
            return bool(flag)
        t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t[0] > 0.5)

    ctx = lk.Context(local_rank)

    def measure(n_total, batch_regions, steps, warmup, sample_clocks):
        """Both measurements of one sharded configuration; every rank returns the same reduced numbers."""
        sb = ShardedBench(ctx, n_total, batch_regions, rank, world, local_rank, gather_to_host=args.e2e_gather_to_host)
        for _ in range(warmup):
            sb.step()
        sampler = ClockSampler(local_rank) if sample_clocks else None
        barrier()
        if sampler:
            sampler.start()
        t_wall0 = time.perf_counter()
        step_ms, fast_ms = [], []
        for _ in range(steps):
            m, f = sb.step()
            step_ms.append(m)
            fast_ms.append(f)
        barrier()
        t_wall = time.perf_counter() - t_wall0
        clocks = sampler.stop() if sampler else None
        ok_res, n_chk = sb.verify_resident() if world > 1 else (True, 0)
        sb.setup_e2e()
        sb.e2e_step()  # warm the staging buffers
        sb.e2e_finish()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            sb.e2e_step()
        sb.e2e_finish()  # the last step's gathered set is in host memory
        barrier()
        e2e_s = time.perf_counter() - t0
        ok_e2e, n_chk2 = sb.verify_e2e() if world > 1 else (True, 0)
        dev_ms, e2e_s = reduce_max(float(sum(step_ms)), e2e_s)
        total_cells = reduce_sum(sb.cells)
        r = {
            "value": total_cells * steps / (dev_ms * 1e-3) / 1e9, "ms_per_step": dev_ms / steps,
            "e2e_value": total_cells * steps / e2e_s / 1e9, "e2e_ms_per_step": e2e_s / steps * 1e3,
            "total_cells": total_cells, "cells_rank0": sb.cells, "pairs_rank0": sb.n_mine, "regions_total": n_total,
            "fast_ms": float(sum(fast_ms) / max(1, len(fast_ms))), "launches": sb.res.launch_count * steps,
            "h2d": sb.h2d, "d2h": sb.d2h, "t_wall": t_wall, "clocks": clocks, "input_bytes": sb.input_bytes,
            "gather_verified": (all_true(ok_res and ok_e2e) if world > 1 else None),
            "gather_checks_per_rank": n_chk + n_chk2, "equal_slabs": sb.equal_slabs,
        }
        sb.close()
        del sb
        return r

    if args.config == "cfg5":
        main_r = measure(args.cfg5_regions, 256, args.steps, args.warmup, True)
        scaling, workload = "strong", ("cfg5 scale-out: %d cfg2-style regions dealt round-robin in batches of 256 to %d GPU(s), "
                                       "all_gather of the result slabs" % (args.cfg5_regions, world))
        cfg5 = None
    else:
        # weak scaling: rank r owns regions [r*R, (r+1)*R) of the generator's id space (= one batch of R regions per rank)
        main_r = measure(args.regions * world, args.regions, args.steps, args.warmup, True)
        scaling, workload = "weak", workload_name(args.regions)
        cfg5 = None
        if not args.no_extras and not args.no_cfg5:
            try:
                c = measure(args.cfg5_regions, 256, args.cfg5_steps, 1, False)
                cfg5 = {"regions": args.cfg5_regions, "batch_regions": 256, "scaling": "strong", "n_gpus": world, "steps": args.cfg5_steps,
                        "cells": c["total_cells"], "value": c["value"], "unit": UNIT, "ms_per_step": c["ms_per_step"],
                        "e2e": {"value": c["e2e_value"], "ms_per_step": c["e2e_ms_per_step"], "h2d_bytes_per_step": c["h2d"],
                                "d2h_bytes_per_step": c["d2h"]},
                        "gather_verified": c["gather_verified"], "gather_checks_per_rank": c["gather_checks_per_rank"],
                        "note": "BASELINE cfg#5: round-robin 256-region batches through lorikeet_b200.dispatch.ShardedRegions, the "
                                "all_gather inside both numbers; total work fixed as N grows"}
            except Exception as e:
                cfg5 = {"error": str(e)}

    if rank == 0:
        peaks = measured_peaks()
        sm_max = (peaks or {}).get("sm_max_mhz", 1965.0)
        n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
        peak_derived = n_sm * 128 * 2 * sm_max * 1e6 / 1e12
        fma_measured = ctx.measure_fp32_fma_tflops(300.0)
        cells = main_r["cells_rank0"]
        avg_fast_ms = main_r["fast_ms"]
        achieved = FLOP_PER_CELL * cells / (avg_fast_ms * 1e-3) / 1e12
        hbm = (peaks or {}).get("hbm_gbs", 6650.0)
        alg_bytes = main_r["input_bytes"] + main_r["pairs_rank0"] * 8
        regions_rank0 = cells / (main_r["total_cells"] / main_r["regions_total"])
        # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture
        # (profiles/r01_pairhmm_fast_k19_seltab_ncu_details.txt: dram read+write of a 1000-region launch), scaled to
        # this launch's region count; ncu itself is never run inside the bench
        traffic = NCU_DRAM_BYTES_PER_REGION * regions_rank0
        roofline = {
            "bound": "fp32", "achieved": achieved, "peak": peak_derived, "unit": "TFLOP/s", "frac": achieved / peak_derived,
            "traffic": traffic, "traffic_unit": "bytes of DRAM read+write per launch set (ncu capture, scaled by regions)",
            "peak_source": "SMs x 128 lanes x 2 flop x sm_max_mhz of MEASURED_PEAKS.json" if peaks else
                           "SMs x 128 x 2 x 1965 MHz (fallback: MEASURED_PEAKS.json absent)",
            "peak_fma_microbench": fma_measured, "frac_of_fma_microbench": achieved / fma_measured if fma_measured else None,
            "kernel": "pairhmm_fast_kernel", "kernel_ms_per_launch_set": avg_fast_ms,
            "algorithmic_flop_per_cell": FLOP_PER_CELL,
            "hbm_view": {"bound": "hbm", "achieved": alg_bytes / (avg_fast_ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                         "frac": alg_bytes / (avg_fast_ms * 1e-3) / 1e9 / hbm,
                         "note": "scalar DP with ~1e-3 B/cell: far below the HBM roof by construction (SURVEY 8d)"},
        }
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from lorikeet_b200 import synth
            sample_batch = synth.pairhmm_regions(args.cpu_sample_regions)
            v, cores, dt, sample = cpu_baseline_run(sample_batch, args.cpu_sample_regions)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "first %d regions of the same generator (%.3e cells, %.1f s)" % (sample.n_regions, sample.n_cells, dt),
                   "note": "C restatement of the reference's in-tree scalar f64 path (the reference is Rust; gkl's AVX path is "
                           "not vendored)"}
            try:
                from oracle import lk_oracle as O
                lib_path = O.build(native=True)
                t0 = time.perf_counter()
                O.pairhmm_batch_vector(sample_batch, n_threads=cores, lib_path=lib_path)
                dtv = time.perf_counter() - t0
                cpu["vector_stand_in"] = {"value": sample_batch.n_cells / dtv / 1e9, "unit": UNIT, "cores": cores, "isa": O.vector_isa(lib_path),
                                          "kind": "port", "note": "OUR fp32 SIMD anti-diagonal kernel standing in for the un-vendored gkl "
                                                                  "AVX path; same sample (%.1f s)" % dtv}
            except Exception as e:
                cpu["vector_stand_in"] = {"error": str(e)}
        extras = None
        if world == 1 and not args.no_extras:
            extras = run_extras(ctx, args, peak_derived)
        if cfg5 is not None:
            extras = extras or {}
            extras["cfg5_scale_out"] = cfg5
        line = {
            "metric": METRIC, "value": main_r["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main_r["ms_per_step"], "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "vs_baseline_note": "BASELINE.md holds no published number for this metric (published: {})",
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "config": args.config, "regions_total": main_r["regions_total"],
                       "cells_per_step": main_r["total_cells"], "cells_rank0_per_step": cells, "pairs_rank0": main_r["pairs_rank0"],
                       "l2_policy": "inputs (%.0f MB per GPU) exceed the 126 MB L2; no flush needed" % (main_r["input_bytes"] / 1e6),
                       "collective": "all_gather of f64 result slabs per step, inside value AND e2e" if world > 1 else "none (N=1)",
                       "gather_verified": main_r["gather_verified"], "gather_checks_per_rank": main_r["gather_checks_per_rank"],
                       "parallelism": "region batches dealt round-robin, one rank per GPU (lorikeet_b200.dispatch.ShardedRegions)"},
            "e2e": {"value": main_r["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": main_r["h2d"], "d2h_bytes_per_step": main_r["d2h"],
                    "ms_per_step": main_r["e2e_ms_per_step"],
                    "api": "lk_engine_compute_read_likelihoods = compute_read_likelihoods: raw reads + MAPQ in pinned host "
                           "buffers -> modify_read_qualities, PairHMM, normalise, filter on the device -> [allele][read] "
                           "matrices + keep mask" + ((" -> this rank's slab to its pinned host memory + all_gather (every GPU holds the "
                                                      "full set)" + (" + the full gathered set to every rank's host" if args.e2e_gather_to_host else "") +
                                                      "; collective and D2H on a side stream, overlapping the next step") if world > 1 else " on the host")},
            "gpu_launches": main_r["launches"],
            "clocks": main_r["clocks"],
            "roofline": roofline,
            "cpu_baseline": cpu,
            "wall_s_timed_region": main_r["t_wall"],
            "other_configs": extras,
        }
        emit(line)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
