import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import lorikeet_b200 as lk
from lorikeet_b200 import synth
b = synth.pairhmm_regions(10000)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
keep = {k: pin(getattr(b, k)) for k in ("read_bases", "read_qual", "read_ins", "read_del", "hap_bases", "read_off", "hap_off", "region_read_start", "region_hap_start")}
pb = lk.PairHmmBatch(keep["region_read_start"].numpy(), keep["region_hap_start"].numpy(), keep["read_off"].numpy(), keep["read_bases"].numpy(), keep["read_qual"].numpy(), keep["read_ins"].numpy(), keep["read_del"].numpy(), keep["hap_off"].numpy(), keep["hap_bases"].numpy())
out = torch.empty(b.n_pairs, dtype=torch.float64).pin_memory().numpy(); fl = torch.empty(b.n_pairs, dtype=torch.uint8).pin_memory().numpy()
with lk.Context(0) as ctx:
    for i in range(3):
        t = time.perf_counter(); ctx.pairhmm_batch(pb, out=out, flags=fl); print("e2e %.1f ms" % ((time.perf_counter() - t) * 1e3), flush=True)
    os.environ["LK_TRACE"] = "1"
    t = time.perf_counter(); ctx.pairhmm_batch(pb, out=out, flags=fl); print("e2e traced %.1f ms" % ((time.perf_counter() - t) * 1e3), flush=True)
    # the region driver (bench.py's e2e path) on the same pinned batch
    from lorikeet_b200.engine import PairHMMLikelihoodCalculationEngine
    del os.environ["LK_TRACE"]
    eng = PairHMMLikelihoodCalculationEngine(ctx)
    mapq = pin(np.full(b.n_reads, 60, np.uint8)).numpy()
    keepm = torch.empty(b.n_reads, dtype=torch.uint8).pin_memory().numpy()
    for i in range(3):
        t = time.perf_counter(); eng.compute_read_likelihoods(pb, mapq, values=out, keep=keepm); print("engine e2e %.1f ms" % ((time.perf_counter() - t) * 1e3), flush=True)
    os.environ["LK_TRACE"] = "1"
    t = time.perf_counter(); eng.compute_read_likelihoods(pb, mapq, values=out, keep=keepm); print("engine e2e traced %.1f ms" % ((time.perf_counter() - t) * 1e3), flush=True)
