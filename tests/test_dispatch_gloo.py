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
