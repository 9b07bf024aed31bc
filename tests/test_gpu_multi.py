"""GPU tests (-m gpu) of the multi-GPU region dispatcher: world 1 always; world 2 over NCCL when the box has >= 2 GPUs."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_dispatcher_world1_equals_host_path():
    import lorikeet_b200 as lk
    from lorikeet_b200 import synth
    from lorikeet_b200.dispatch import RegionDispatcher
    b = synth.pairhmm_regions(40, first_region=300)
    with lk.Context(0) as ctx:
        a = RegionDispatcher(ctx, batch_regions=7).compute(b)
        assert np.array_equal(a, ctx.pairhmm_batch(b))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import lorikeet_b200 as lk
    from lorikeet_b200 import synth
    from lorikeet_b200.dispatch import RegionDispatcher
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        b = synth.pairhmm_regions(50, first_region=900)
        with lk.Context(rank) as ctx:
            out = RegionDispatcher(ctx, batch_regions=8).compute(b)
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


def test_dispatcher_world2_nccl():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import lorikeet_b200 as lk
    from lorikeet_b200 import synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    results = dict(q.get(timeout=300) for _ in range(2))
    [p.join(timeout=60) for p in procs]
    b = synth.pairhmm_regions(50, first_region=900)
    with lk.Context(0) as ctx:
        exp = ctx.pairhmm_batch(b)
    assert np.array_equal(results[0], exp) and np.array_equal(results[1], exp)


def test_two_devices_in_one_process():
    # the kernels' shared-memory opt-in is per function AND per device: a second context on another device in the same
    # process must work (one host thread + lk_ctx per device is the layout SURVEY 8(e) describes)
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import lorikeet_b200 as lk
    from lorikeet_b200 import synth
    b = synth.pairhmm_regions(12, first_region=500)
    ro, refs, ao, alts = synth.sw_pairs(200)
    outs = []
    for dev in (0, 1):
        with lk.Context(dev) as ctx:
            outs.append((ctx.pairhmm_batch(b), ctx.sw_batch(ro, refs, ao, alts, [200, -150, -260, -11], "SoftClip")))
    assert np.array_equal(outs[0][0], outs[1][0])
    for a, c in zip(outs[0][1], outs[1][1]):
        assert np.array_equal(a, c)
