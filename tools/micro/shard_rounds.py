"""What a sharded sum-check round and the bulk all-gather cost (torchrun worker, one rank per GPU):
ms per sum-check (eq*a*b over n variables, 2^(n - log2 G) entries per rank) for several numbers R of sharded rounds.
(t(R2) - t(R1)) / (R2 - R1) compared with the local per-round time at the same sizes = cost of one in-kernel exchange."""
import os, sys
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
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = hl.Context(local)
hl.dist_init(ctx, rank, world)
stream = torch.cuda.ExternalStream(ctx.stream, device=local)
g = world.bit_length() - 1
one = mont_one()
for n in (int(a) for a in (sys.argv[1:] or ["14", "21"])):
    nl = n - g
    polys = [hl.MultilinearPolynomial.new(ctx, rand_canonical(seed + 10 * rank, 1 << nl)) for seed in (1, 2)]
    y = rand_canonical(3, n)
    res = {}
    for R in sorted({0, 2, 4, 8, min(12, nl), nl}):
        if R > nl:
            continue
        def run():
            hl.Keccak256Transcript(ctx)
            hl.sumcheck_prove_evals_sharded(ctx, n, polys, one.reshape(1, 4), y, one, sharded_rounds=R)
        for _ in range(3):
            run()
        ctx.sync(); dist.barrier()
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            run()
        e1.record(stream)
        ctx.sync(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[R] = round(float(t.item()), 4)
    hl.dist_check(ctx)
    if rank == 0:
        print(f"SHARD_ROUNDS world={world} n={n} local_entries=2^{nl} ms_by_sharded_rounds={res}", flush=True)
dist.barrier()
ctx.close()
dist.destroy_process_group()
