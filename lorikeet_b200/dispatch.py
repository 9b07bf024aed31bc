"""Multi-GPU region dispatcher: active regions are independent (compute_read_likelihoods takes only one region's
haplotypes and reads, pair_hmm_likelihood_calculation_engine.rs:195-200), so region batches are dealt round-robin
to the ranks with NO data-path collective; a single all_gather of the per-rank result slabs at the end gives every
rank the full likelihood set (BASELINE.json north_star / SURVEY.md 8(e)).

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch on the GPU box; the same code runs over gloo on
CPU tensors in the tests, where the per-rank compute function is injected).
"""
import numpy as np


def shard_plan(n_regions, world, batch_regions=256):
    """Round-robin deal of region batches: returns per rank a list of (g0, g1) region ranges."""
    plan = [[] for _ in range(world)]
    for i, g0 in enumerate(range(0, n_regions, batch_regions)):
        plan[i % world].append((g0, min(n_regions, g0 + batch_regions)))
    return plan


def _concat_ranges(batch, ranges):
    """Sub-batch made of the given region ranges (in order), offsets rebased."""
    from .batch import PairHmmBatch
    if not ranges:
        z = np.zeros(1, dtype=np.int64)
        e = np.zeros(0, dtype=np.uint8)
        return PairHmmBatch(z, z.copy(), z.copy(), e, e, e, e, z.copy(), e, read_gcp=None, gcp_constant=batch.gcp_constant)
    parts = [batch.region_slice(g0, g1) for g0, g1 in ranges]
    if len(parts) == 1:
        return parts[0]
    rrs, rhs, roff, hoff = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)], [np.zeros(1, np.int64)], [np.zeros(1, np.int64)]
    nr = nh = rb = hb = 0
    for p in parts:
        rrs.append(p.region_read_start[1:] + nr); rhs.append(p.region_hap_start[1:] + nh)
        roff.append(p.read_off[1:] + rb); hoff.append(p.hap_off[1:] + hb)
        nr += p.n_reads; nh += p.n_haps; rb += int(p.read_off[-1]); hb += int(p.hap_off[-1])
    cat = lambda k: np.concatenate([getattr(p, k) for p in parts])
    gcp = None if batch.read_gcp is None else cat("read_gcp")
    return PairHmmBatch(np.concatenate(rrs), np.concatenate(rhs), np.concatenate(roff), cat("read_bases"), cat("read_qual"),
                        cat("read_ins"), cat("read_del"), np.concatenate(hoff), cat("hap_bases"), read_gcp=gcp,
                        gcp_constant=batch.gcp_constant)


class RegionDispatcher:
    """compute(batch) -> log10 likelihoods of EVERY region (layout of lk_pairhmm_batch_host) on every rank."""

    def __init__(self, ctx=None, batch_regions=256, group=None):
        self.ctx = ctx
        self.batch_regions = batch_regions
        self.group = group

    def compute(self, batch, compute_fn=None):
        import torch
        import torch.distributed as dist
        distributed = dist.is_available() and dist.is_initialized()
        world = dist.get_world_size(self.group) if distributed else 1
        rank = dist.get_rank(self.group) if distributed else 0
        plan = shard_plan(batch.n_regions, world, self.batch_regions)
        pair_counts = batch.pair_counts()
        pair_base = np.zeros(batch.n_regions + 1, dtype=np.int64)
        np.cumsum(pair_counts, out=pair_base[1:])
        sizes = [int(sum(pair_base[g1] - pair_base[g0] for g0, g1 in plan[r])) for r in range(world)]
        mine = _concat_ranges(batch, plan[rank])
        use_gpu = compute_fn is None
        if use_gpu:
            if self.ctx is None:
                raise RuntimeError("RegionDispatcher needs a lorikeet_b200.Context (there is no CPU fallback)")
            res = self.ctx.pairhmm_upload(mine)
            res.run()
            if world == 1:
                out = res.download()
                res.free()
                return out

            class _Ptr:
                def __init__(self, ptr, n):
                    self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 3}
            slab = torch.zeros(max(sizes), dtype=torch.float64, device="cuda")
            if sizes[rank]:
                slab[:sizes[rank]] = torch.as_tensor(_Ptr(res.out_device_ptr, sizes[rank]), device="cuda")
            gathered = torch.empty(world * max(sizes), dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(gathered, slab, group=self.group)  # the one collective of the path
            res.free()
            gathered = gathered.cpu().numpy()
        else:
            local = np.asarray(compute_fn(mine), dtype=np.float64) if sizes[rank] else np.zeros(0)
            assert local.size == sizes[rank]
            if world == 1:
                return local
            slab = torch.zeros(max(sizes), dtype=torch.float64)
            slab[:sizes[rank]] = torch.from_numpy(local)
            parts = [torch.empty_like(slab) for _ in range(world)]
            dist.all_gather(parts, slab, group=self.group)
            gathered = torch.cat(parts).numpy()
        # scatter the per-rank slabs back into global region order
        out = np.empty(int(pair_base[-1]), dtype=np.float64)
        stride = max(sizes)
        for r in range(world):
            pos = r * stride
            for g0, g1 in plan[r]:
                n = int(pair_base[g1] - pair_base[g0])
                out[pair_base[g0]:pair_base[g1]] = gathered[pos:pos + n]
                pos += n
        return out
