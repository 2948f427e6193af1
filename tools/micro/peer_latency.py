"""Per-round cost of the fused peer all-gather: sharded sum-check on tiny tables (launch under torchrun)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import halo2_lasso_b200 as hl
from bench import rand_canonical, mont_one

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group(os.environ.get("PEER_BACKEND", "nccl"), **({"device_id": torch.device("cuda", local)} if os.environ.get("PEER_BACKEND", "nccl") == "nccl" else {}))
ctx = hl.Context(local)
hl.dist_init(ctx, rank, world)
g = world.bit_length() - 1
one = mont_one()
for n_loc in (8, 14, 20):
    n_tot = n_loc + g
    polys = [hl.MultilinearPolynomial.new(ctx, rand_canonical(s + 10 * rank, 1 << n_loc)) for s in (1, 2)]
    y = rand_canonical(3, n_tot)
    def run():
        hl.Keccak256Transcript(ctx)
        hl.sumcheck_prove_evals_sharded(ctx, n_tot, polys, one.reshape(1, 4), y, one)
    for _ in range(5):
        run()
    ctx.sync(); dist.barrier()
    t0 = time.perf_counter()
    reps = 30
    for _ in range(reps):
        run()
    ctx.sync()
    dt = (time.perf_counter() - t0) / reps
    # same shape without the exchange: plain sum-check on the local slice
    yl = y[:n_loc]
    def run1():
        hl.Keccak256Transcript(ctx)
        hl.ClassicSumCheck.prove_evals(ctx, n_loc, polys, one.reshape(1, 4), yl, one)
    for _ in range(5):
        run1()
    ctx.sync(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        run1()
    ctx.sync()
    d1 = (time.perf_counter() - t0) / reps
    if rank == 0:
        print(f"world {world} n_loc {n_loc}: sharded {dt*1e3:.3f} ms ({n_tot} rounds), local-only {d1*1e3:.3f} ms ({n_loc} rounds), "
              f"extra per sharded round {(dt - d1) / n_loc * 1e6:.1f} us", flush=True)
    dist.barrier()
dist.destroy_process_group()
