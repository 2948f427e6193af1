"""world_size-2 gloo test (CPU) of the host-side multi-GPU plumbing: handle exchange in rank order and the
hypercube slice arithmetic used by the sharded sum-check / MSM."""
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import halo2_lasso_b200 as hl

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([rank]) * 64
    hs = hl.exchange_handles(mine, world)
    ok = hs == [bytes([r]) * 64 for r in range(world)]
    lo, hi = hl.shard_slice(10, rank, world)
    ok = ok and (lo, hi) == (rank * (1024 // world), (rank + 1) * (1024 // world))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_handle_exchange_and_slices_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_slice_covers_the_hypercube():
    sys.path.insert(0, ROOT)
    import halo2_lasso_b200 as hl

    for world in (1, 2, 4, 8):
        spans = [hl.shard_slice(12, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 1 << 12
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
