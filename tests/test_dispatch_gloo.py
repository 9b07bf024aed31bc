"""CPU test of the N>1 path: world_size-2 gloo run of the region dispatcher (round-robin region batches, one
all_gather of result slabs).  The per-rank compute function is the oracle here (no GPU in this container); on the GPU
box the same dispatcher runs lk kernels + NCCL (tests/test_gpu_multi.py, bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_regions, batch_regions, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from lorikeet_b200 import synth
    from lorikeet_b200.dispatch import RegionDispatcher
    from oracle import lk_oracle as O
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        batch = synth.pairhmm_regions(n_regions, n_threads=1, n_reads=5, read_len=40, n_haps=3, hap_len=120)
        out = RegionDispatcher(batch_regions=batch_regions).compute(batch, compute_fn=lambda b: O.pairhmm_batch(b, n_threads=1))
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_regions,batch_regions", [(11, 2), (3, 256), (8, 4)])
def test_two_rank_gloo_dispatch_matches_single_process(n_regions, batch_regions):
    import torch.multiprocessing as mp
    from lorikeet_b200 import synth
    from oracle import lk_oracle as O
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_regions, batch_regions, q)) for r in range(2)]
    [p.start() for p in procs]
    results = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    batch = synth.pairhmm_regions(n_regions, n_threads=1, n_reads=5, read_len=40, n_haps=3, hap_len=120)
    exp = O.pairhmm_batch(batch, n_threads=1)
    assert np.array_equal(results[0], exp) and np.array_equal(results[1], exp)


def test_shard_plan_round_robin():
    from lorikeet_b200.dispatch import shard_plan
    plan = shard_plan(1000, 4, 256)
    assert plan == [[(0, 256)], [(256, 512)], [(512, 768)], [(768, 1000)]]
    plan = shard_plan(10, 3, 2)
    assert plan == [[(0, 2), (6, 8)], [(2, 4), (8, 10)], [(4, 6)]]
    covered = sorted(x for r in plan for x in r)
    assert covered[0][0] == 0 and covered[-1][1] == 10 and all(a[1] == b[0] for a, b in zip(covered, covered[1:]))


def test_slab_layout_addresses_every_batch_of_a_ragged_deal():
    """SlabLayout: where a region batch sits in the gathered buffer when the per-rank slabs differ in size (cfg#5 at N = 8:
    782 batches over 8 ranks, the last batch short) -- what bench.py's gather check and RegionDispatcher.compute rely on."""
    from lorikeet_b200.dispatch import SlabLayout, shard_plan
    n_regions, world, B = 200000, 8, 256
    plan = shard_plan(n_regions, world, B)
    assert sum(len(p) for p in plan) == 782 and [len(p) for p in plan] == [98] * 6 + [97] * 2
    counts = [[(g1 - g0) * 1024 for g0, g1 in p] for p in plan]
    lay = SlabLayout(plan, counts)
    assert lay.stride == 98 * 256 * 1024 - (256 - 64) * 1024 or lay.stride == 98 * 256 * 1024  # the short batch sits on rank 5
    assert lay.stride == max(lay.sizes) and lay.sizes[6] == 97 * 256 * 1024
    # walk every batch: positions are disjoint, inside their rank's slab, and reorder() inverts the deal
    pair_base = np.arange(n_regions + 1, dtype=np.int64) * 4  # 4 pairs per region for a small check
    small = SlabLayout(plan, [[(g1 - g0) * 4 for g0, g1 in p] for p in plan])
    gathered = np.full(world * small.stride, -1.0)
    for r in range(world):
        for i, (g0, g1) in enumerate(plan[r]):
            pos, n = small.position(r, i)
            assert r * small.stride <= pos and pos + n <= (r + 1) * small.stride and n == (g1 - g0) * 4
            assert np.all(gathered[pos:pos + n] == -1.0)
            gathered[pos:pos + n] = np.arange(pair_base[g0], pair_base[g1])
    assert np.array_equal(small.reorder(gathered, pair_base), np.arange(n_regions * 4, dtype=np.float64))


def _layout_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from lorikeet_b200 import synth
    from lorikeet_b200.dispatch import ShardedRegions
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        sh = ShardedRegions(13, batch_regions=3)  # 5 batches over 2 ranks: 3 + 2, the last one short
        gen = lambda g0, g1: synth.pairhmm_regions(g1 - g0, first_region=g0, n_threads=1, n_reads=2 + (g0 % 3), read_len=40, n_haps=2, hap_len=120)
        mine = sh.build(gen)
        lay = sh.agree_layout(mine)
        q.put((rank, lay.batch_pairs, lay.sizes, lay.stride, mine.n_regions))
    finally:
        dist.destroy_process_group()


def test_sharded_regions_agree_on_a_ragged_layout_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_layout_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = dict((r[0], r[1:]) for r in (q.get(timeout=120) for _ in range(2)))
    [p.join(timeout=60) for p in procs]
    assert res[0][:3] == res[1][:3]                      # both ranks hold the same layout
    batch_pairs, sizes, stride, _ = res[0]
    # batches start at regions 0, 3, 6, 9, 12 with 2 + (g0 % 3) reads x 2 haplotypes per region
    exp = {0: 3 * 2 * 2, 3: 3 * 2 * 2, 6: 3 * 2 * 2, 9: 3 * 2 * 2, 12: 1 * 2 * 2}
    assert batch_pairs == [[exp[0], exp[6], exp[12]], [exp[3], exp[9]]]
    assert sizes == [sum(batch_pairs[0]), sum(batch_pairs[1])] and stride == max(sizes)
    assert res[0][3] == 7 and res[1][3] == 6
