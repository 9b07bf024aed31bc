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


def run_extras(ctx):
    """Side measurements of BASELINE.json's other configs (kernel-resident, best of 3; NOT the bench metric)."""
    from lorikeet_b200 import synth
    out = {}
    try:
        ro, refs, ao, alts = synth.sw_pairs(100000)  # cfg4: 100k pairs, 500 bp + 10xN pads, NEW_SW_PARAMETERS, SoftClip
        res = ctx.sw_upload(ro, refs, ao, alts, [200, -150, -260, -11], "SoftClip", max_ops=64)
        res.run()
        ms = min(res.run() for _ in range(3))
        out["cfg4_smith_waterman"] = {"pairs": 100000, "cells": res.n_cells, "ms": ms, "gcups": res.n_cells / ms / 1e6,
                                      "note": "fill + in-kernel traceback, CIGARs bit-exact vs oracle (tests/test_gpu_sw.py)"}
        n_cells = res.n_cells
        res.free()
        # the same batch end to end through lk_sw_batch_host: pinned host buffers in, CIGARs + offsets in pinned host buffers out
        import numpy as np
        import torch
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
        b = synth.pairhmm_regions(111, **synth.CFG3)  # cfg3 slice: 111 of 2000 regions (8 reads x 10 kbp, 4 haps x 12 kbp);
        # 111 x 32 pairs = 2 full waves of the band kernel's 1776 resident warps, as the 36-wave full config is
        res = ctx.pairhmm_upload(b)
        res.run()
        ms = min(res.run() for _ in range(2))
        _, flags = res.download(want_flags=True)
        out["cfg3_long_reads_slice"] = {"regions": 111, "cells": res.n_cells, "ms": ms, "gcups": res.n_cells / ms / 1e6,
                                        "pairs_redone_in_f64": int((flags & 1).sum()),
                                        "note": "fp32 band kernel with per-lane block floating point"}
        res.free()
    except Exception as e:  # side measurements must never break the bench line
        out["error"] = str(e)
    return out


def run_reference(args, rank):
    """--impl reference: the reference's own CPU algorithm (oracle port; the Rust original cannot be built here)."""
    if rank != 0:
        return
    from lorikeet_b200 import synth
    regions = args.ref_sample_regions
    batch = synth.pairhmm_regions(regions)
    from oracle import lk_oracle as O
    try:
        lib_path = O.build(native=True)
    except Exception:
        lib_path = None
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        O.pairhmm_batch(batch.region_slice(0, max(1, regions // 8)), n_threads=cores, lib_path=lib_path)
    t = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        O.pairhmm_batch(batch, n_threads=cores, lib_path=lib_path)
        t.append(time.perf_counter() - t0)
    total = sum(t)
    value = batch.n_cells * args.steps / total / 1e9
    sample = "%d regions of the cfg2 generator (%.3e cells) per step" % (regions, batch.n_cells)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.regions), "sample": sample,
                   "note": "reference is Rust and cannot be built here; this is the C restatement of its scalar f64 path "
                           "(oracle/), one region per thread as Rayon schedules it"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else written to fd 1 during the run (NCCL's version banner,
    library chatter) has been diverted to stderr by main()."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


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
    ap.add_argument("--regions", type=int, default=10000, help="regions per GPU (cfg2 = 10000)")
    ap.add_argument("--ref-sample-regions", type=int, default=192, help="regions per step of the reference arm")
    ap.add_argument("--cpu-sample-regions", type=int, default=256, help="regions of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the SW (cfg4) and long-read (cfg3 slice) side measurements")
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
    from lorikeet_b200 import synth

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

    # ---- this rank's shard: regions [rank*R, (rank+1)*R) of the generator's id space (weak scaling) ----
    batch = synth.pairhmm_regions(args.regions, first_region=rank * args.regions)
    cells = batch.n_cells
    ctx = lk.Context(local_rank)
    res = ctx.pairhmm_upload(batch)

    gathered = None
    if world > 1:
        class _Ptr:  # zero-copy torch view of the library's device result block
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}
        local_out = torch.as_tensor(_Ptr(res.out_device_ptr, res.n_pairs), device="cuda")
        gathered = torch.empty(world * res.n_pairs, dtype=torch.float64, device="cuda")

    def step():
        ms = res.run()  # prologue + fp32 wavefront + f64 pass; CUDA events on the launching stream
        fast_ms = res.last_fast_ms
        if world > 1:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_gather_into_tensor(gathered, local_out)  # the one collective of the path: result slabs
            e1.record()
            e1.synchronize()
            ms += e0.elapsed_time(e1)
        return ms, fast_ms

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t_wall0 = time.perf_counter()
    step_ms, fast_ms = [], []
    for _ in range(args.steps):
        m, f = step()
        step_ms.append(m)
        fast_ms.append(f)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    dev_ms = float(sum(step_ms))
    launches = res.launch_count * args.steps

    # ---- end to end through the C-ABI host call, inputs in pinned host memory ----
    def pin(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t, t.numpy()
    keep = []
    pinned = {}
    for k in ("read_bases", "read_qual", "read_ins", "read_del", "hap_bases", "read_off", "hap_off",
              "region_read_start", "region_hap_start"):
        t, a = pin(getattr(batch, k))
        keep.append(t)
        pinned[k] = a
    pbatch = lk.PairHmmBatch(pinned["region_read_start"], pinned["region_hap_start"], pinned["read_off"],
                             pinned["read_bases"], pinned["read_qual"], pinned["read_ins"], pinned["read_del"],
                             pinned["hap_off"], pinned["hap_bases"], read_gcp=None, gcp_constant=batch.gcp_constant)
    e2e_steps = args.steps
    out_t = torch.empty(res.n_pairs, dtype=torch.float64).pin_memory()
    flg_t = torch.empty(res.n_pairs, dtype=torch.uint8).pin_memory()
    out_host, flg_host = out_t.numpy(), flg_t.numpy()
    ctx.pairhmm_batch(pbatch, out=out_host, flags=flg_host)  # warm the staging buffers
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.pairhmm_batch(pbatch, out=out_host, flags=flg_host)
    barrier()
    e2e_bare_s = time.perf_counter() - t0

    # The reference-facing call of the path is the region driver compute_read_likelihoods
    # (pair_hmm_likelihood_calculation_engine.rs:195-242): RAW reads in (same generator, pre-ops not applied), MAPQ,
    # modify_read_qualities + PairHMM + normalise + filter on the device, [allele][read] matrices and the keep mask out.
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine, PCRErrorModel
    raw = synth.pairhmm_regions(args.regions, first_region=rank * args.regions, pcr_model=0, base_quality_score_threshold=0)
    rkeep = []
    rp = {}
    for k in ("read_bases", "read_qual", "read_ins", "read_del", "hap_bases", "read_off", "hap_off",
              "region_read_start", "region_hap_start"):
        t, a = pin(getattr(raw, k))
        rkeep.append(t)
        rp[k] = a
    rbatch = lk.PairHmmBatch(rp["region_read_start"], rp["region_hap_start"], rp["read_off"], rp["read_bases"], rp["read_qual"],
                             rp["read_ins"], rp["read_del"], rp["hap_off"], rp["hap_bases"], read_gcp=None,
                             gcp_constant=raw.gcp_constant)
    mapq_t, mapq = pin(np.full(raw.n_reads, 60, np.uint8))
    eng = PairHMMLikelihoodCalculationEngine(ctx, 10, -4.5, PCRErrorModel.CONSERVATIVE, 18, False, 1.0, 0.02, True, False, False)
    keep_t = torch.empty(raw.n_reads, dtype=torch.uint8).pin_memory()
    values_host, keep_host = out_host, keep_t.numpy()  # results land in pinned host memory, as in the bare call
    eng.compute_read_likelihoods(rbatch, mapq, values=values_host, keep=keep_host)  # warm-up
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.compute_read_likelihoods(rbatch, mapq, values=values_host, keep=keep_host)
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = raw.input_bytes() + int(mapq.nbytes)
    d2h = int(values_host.nbytes + keep_host.nbytes)  # f64 matrices + keep mask

    # ---- reduce over ranks: time = max, work = sum ----
    if world > 1:
        t = torch.tensor([dev_ms, e2e_s, e2e_bare_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_s, e2e_bare_s = float(t[0]), float(t[1]), float(t[2])
        c = torch.tensor([float(cells)], dtype=torch.float64, device="cuda")
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
        total_cells = float(c[0])
    else:
        total_cells = float(cells)

    value = total_cells * args.steps / (dev_ms * 1e-3) / 1e9
    e2e_value = total_cells * e2e_steps / e2e_s / 1e9

    if rank == 0:
        peaks = measured_peaks()
        sm_max = (peaks or {}).get("sm_max_mhz", 1965.0)
        n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
        peak_derived = n_sm * 128 * 2 * sm_max * 1e6 / 1e12
        fma_measured = ctx.measure_fp32_fma_tflops(300.0)
        avg_fast_ms = float(sum(fast_ms) / len(fast_ms))
        achieved = FLOP_PER_CELL * cells / (avg_fast_ms * 1e-3) / 1e12
        hbm = (peaks or {}).get("hbm_gbs", 6650.0)
        alg_bytes = batch.input_bytes() + out_host.nbytes
        # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture
        # (profiles/r01_pairhmm_fast_k19_seltab_ncu_details.txt: dram read+write of a 1000-region launch), scaled to
        # this launch's region count; ncu itself is never run inside the bench
        traffic = NCU_DRAM_BYTES_PER_REGION * args.regions
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
            v, cores, dt, sample = cpu_baseline_run(batch, args.cpu_sample_regions)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "first %d regions of the same batch (%.3e cells, %.1f s)" % (sample.n_regions, sample.n_cells, dt)}
        extras = None
        if world == 1 and not args.no_extras:
            extras = run_extras(ctx)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.regions), "regions_per_gpu": args.regions,
                       "cells_per_gpu_per_step": cells, "pairs_per_gpu": res.n_pairs,
                       "l2_policy": "inputs (%.0f MB per GPU) exceed the 126 MB L2; no flush needed" % (h2d / 1e6),
                       "collective": "all_gather of f64 result slabs per step" if world > 1 else "none (N=1)",
                       "parallelism": "regions sharded round-robin by id range, one rank per GPU"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / e2e_steps * 1e3,
                    "api": "lk_engine_compute_read_likelihoods = compute_read_likelihoods: raw reads + MAPQ in pinned host "
                           "buffers -> modify_read_qualities, PairHMM, normalise, filter on the device -> [allele][read] "
                           "matrices + keep mask on the host",
                    "pairhmm_only": {"value": total_cells * e2e_steps / e2e_bare_s / 1e9, "ms_per_step": e2e_bare_s / e2e_steps * 1e3,
                                     "api": "lk_pairhmm_batch_host (the gkl::pairhmm::forward seam, pre-processed reads)"}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "wall_s_timed_region": t_wall,
            "other_configs": extras,
        }
        emit(line)
    res.free()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
