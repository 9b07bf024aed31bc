"""Multi-GPU region dispatcher: active regions are independent (compute_read_likelihoods takes only one region's
haplotypes and reads, pair_hmm_likelihood_calculation_engine.rs:195-200), so region batches are dealt round-robin
to the ranks with NO data-path collective; a single all_gather of the per-rank result slabs at the end gives every
rank the full likelihood set (BASELINE.json north_star / SURVEY.md 8(e)).

One process per GPU (torch.distributed, NCCL over NVLink/NVSwitch on the GPU box; the same code runs over gloo on
CPU tensors in the tests, where the per-rank compute function is injected).  The in-process form of the same deal
(one host thread + lk_ctx per device behind ONE handle, no torch) is lk_pool_* in the C ABI.

Layout of the gathered buffer: rank r's slab starts at r * stride (stride = the largest slab, in f64 elements) and
holds the pair blocks of r's region batches in the order of its plan; SlabLayout maps a region batch to its place.
"""
import numpy as np


def shard_plan(n_regions, world, batch_regions=256):
    """Round-robin deal of region batches: returns per rank a list of (g0, g1) region ranges."""
    plan = [[] for _ in range(world)]
    for i, g0 in enumerate(range(0, n_regions, batch_regions)):
        plan[i % world].append((g0, min(n_regions, g0 + batch_regions)))
    return plan


def concat_batches(parts, gcp_constant=10):
    """One PairHmmBatch out of several (in order), offsets rebased."""
    from .batch import PairHmmBatch
    if not parts:
        z = np.zeros(1, dtype=np.int64)
        e = np.zeros(0, dtype=np.uint8)
        return PairHmmBatch(z, z.copy(), z.copy(), e, e, e, e, z.copy(), e, read_gcp=None, gcp_constant=gcp_constant)
    if len(parts) == 1:
        return parts[0]
    rrs, rhs, roff, hoff = [np.zeros(1, np.int64)], [np.zeros(1, np.int64)], [np.zeros(1, np.int64)], [np.zeros(1, np.int64)]
    nr = nh = rb = hb = 0
    for p in parts:
        rrs.append(p.region_read_start[1:] + nr); rhs.append(p.region_hap_start[1:] + nh)
        roff.append(p.read_off[1:] + rb); hoff.append(p.hap_off[1:] + hb)
        nr += p.n_reads; nh += p.n_haps; rb += int(p.read_off[-1]); hb += int(p.hap_off[-1])
    cat = lambda k: np.concatenate([getattr(p, k) for p in parts])
    gcp = None if parts[0].read_gcp is None else cat("read_gcp")
    return PairHmmBatch(np.concatenate(rrs), np.concatenate(rhs), np.concatenate(roff), cat("read_bases"), cat("read_qual"),
                        cat("read_ins"), cat("read_del"), np.concatenate(hoff), cat("hap_bases"), read_gcp=gcp,
                        gcp_constant=parts[0].gcp_constant)


def _concat_ranges(batch, ranges):
    """Sub-batch made of the given region ranges (in order), offsets rebased."""
    return concat_batches([batch.region_slice(g0, g1) for g0, g1 in ranges], gcp_constant=batch.gcp_constant)


class SlabLayout:
    """Where every region batch of the plan lives inside the gathered buffer.
    batch_pairs[r][i] = (read, haplotype) pairs of the i-th batch of rank r."""

    def __init__(self, plan, batch_pairs):
        self.plan = plan
        self.batch_pairs = [list(map(int, bp)) for bp in batch_pairs]
        self.sizes = [sum(bp) for bp in self.batch_pairs]
        self.stride = max(self.sizes) if self.sizes else 0
        self.world = len(plan)

    def position(self, rank, i):
        """(offset, count) in f64 elements of batch i of `rank` inside the gathered buffer."""
        return rank * self.stride + sum(self.batch_pairs[rank][:i]), self.batch_pairs[rank][i]

    def reorder(self, gathered, pair_base):
        """Scatter the per-rank slabs back into global region order; pair_base[g] = first pair of region g."""
        out = np.empty(int(pair_base[-1]), dtype=np.float64)
        for r in range(self.world):
            for i, (g0, g1) in enumerate(self.plan[r]):
                pos, n = self.position(r, i)
                out[pair_base[g0]:pair_base[g1]] = gathered[pos:pos + n]
        return out


class ShardedRegions:
    """This rank's share of `n_regions` dealt round-robin in batches of `batch_regions` (cfg#5: 256), and the slab layout
    all ranks agree on.  Works without a process group (world 1)."""

    def __init__(self, n_regions, batch_regions=256, group=None):
        import torch.distributed as dist
        self.group = group
        self.distributed = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if self.distributed else 1
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.n_regions = n_regions
        self.plan = shard_plan(n_regions, self.world, batch_regions)
        self.ranges = self.plan[self.rank]
        self.layout = None

    def build(self, gen_fn):
        """gen_fn(g0, g1) -> PairHmmBatch of regions [g0, g1); returns this rank's shard (its batches concatenated)."""
        return concat_batches([gen_fn(g0, g1) for g0, g1 in self.ranges])

    def agree_layout(self, mine, device=None):
        """Exchange the per-batch pair counts (one small all_gather) so every rank can address every slab."""
        import torch
        import torch.distributed as dist
        pc = mine.pair_counts()
        mine_counts, g = [], 0
        for g0, g1 in self.ranges:
            mine_counts.append(int(pc[g:g + (g1 - g0)].sum()))
            g += g1 - g0
        if self.world == 1:
            counts = [mine_counts]
        else:
            width = max(len(p) for p in self.plan)
            t = torch.zeros(width, dtype=torch.int64, device=device)
            if mine_counts:
                t[:len(mine_counts)] = torch.tensor(mine_counts, dtype=torch.int64)
            parts = [torch.empty_like(t) for _ in range(self.world)]
            dist.all_gather(parts, t, group=self.group)
            counts = [parts[r][:len(self.plan[r])].cpu().tolist() for r in range(self.world)]
        self.layout = SlabLayout(self.plan, counts)
        return self.layout

    def all_gather(self, slab, gathered):
        """THE collective of the path: every rank's (padded) result slab -> every rank.  `slab` has layout.stride elements."""
        import torch.distributed as dist
        if self.world == 1:
            gathered[:slab.numel()].copy_(slab)
        elif slab.is_cuda:
            dist.all_gather_into_tensor(gathered, slab, group=self.group)  # NCCL
        else:  # gloo (CPU tests)
            parts = list(gathered.view(self.world, -1).unbind(0))
            dist.all_gather(parts, slab, group=self.group)
        return gathered


class RegionDispatcher:
    """compute(batch) -> log10 likelihoods of EVERY region (layout of lk_pairhmm_batch_host) on every rank."""

    def __init__(self, ctx=None, batch_regions=256, group=None):
        self.ctx = ctx
        self.batch_regions = batch_regions
        self.group = group

    def compute(self, batch, compute_fn=None):
        import torch
        sh = ShardedRegions(batch.n_regions, self.batch_regions, self.group)
        pair_base = np.zeros(batch.n_regions + 1, dtype=np.int64)
        np.cumsum(batch.pair_counts(), out=pair_base[1:])
        mine = _concat_ranges(batch, sh.ranges)
        use_gpu = compute_fn is None
        if use_gpu and self.ctx is None:
            raise RuntimeError("RegionDispatcher needs a lorikeet_b200.Context (there is no CPU fallback)")
        lay = sh.agree_layout(mine, device="cuda" if use_gpu else None)
        n_mine = lay.sizes[sh.rank]
        if use_gpu:
            res = self.ctx.pairhmm_upload(mine)
            res.run()
            if sh.world == 1:
                out = res.download()
                res.free()
                return out
            slab = torch.zeros(lay.stride, dtype=torch.float64, device="cuda")
            if n_mine:
                slab[:n_mine] = device_view(res.out_device_ptr, n_mine)
            gathered = torch.empty(sh.world * lay.stride, dtype=torch.float64, device="cuda")
            sh.all_gather(slab, gathered)
            res.free()
            gathered = gathered.cpu().numpy()
        else:
            local = np.asarray(compute_fn(mine), dtype=np.float64) if n_mine else np.zeros(0)
            assert local.size == n_mine
            if sh.world == 1:
                return local
            slab = torch.zeros(lay.stride, dtype=torch.float64)
            slab[:n_mine] = torch.from_numpy(local)
            gathered = torch.empty(sh.world * lay.stride, dtype=torch.float64)
            sh.all_gather(slab, gathered)
            gathered = gathered.numpy()
        return lay.reorder(gathered, pair_base)


def device_view(ptr, n, dtype="<f8"):
    """Zero-copy torch view of `n` elements of device memory owned by the library (e.g. lk_resident_out_device_ptr)."""
    import torch

    class _Ptr:
        def __init__(self):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": dtype, "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(_Ptr(), device="cuda")
