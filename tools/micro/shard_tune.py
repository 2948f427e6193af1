"""One torchrun job, several exchange variants (b200_dist_tune): cost of a sharded sum-check round = (t(R=8) - t(R=0)) / 8
for eq*a*b over n = 14 variables. Variants: small-message protocol 1 / 2, heartbeat off / NVLink stores / HBM reads."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
os.environ.setdefault("B200_HEARTBEAT", "1,1000,0")  # creates the heartbeat stream; switched per variant below
import numpy as np
import torch
import torch.distributed as dist

import halo2_lasso_b200 as hl
from bench import mont_one, rand_canonical

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = hl.Context(local)
hl.dist_init(ctx, rank, world)
stream = torch.cuda.ExternalStream(ctx.stream, device=local)
g = world.bit_length() - 1
one = mont_one()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 14
nl = n - g
polys = [hl.MultilinearPolynomial.new(ctx, rand_canonical(seed + 10 * rank, 1 << nl)) for seed in (1, 2)]
y = rand_canonical(3, n)


def tune(key, value):
    hl._chk(hl.lib().b200_dist_tune(ctx.h, C.c_int(key), C.c_int(value)), "dist_tune")


def measure(R, reps=20):
    def run():
        hl.Keccak256Transcript(ctx)
        hl.sumcheck_prove_evals_sharded(ctx, n, polys, one.reshape(1, 4), y, one, sharded_rounds=R)
    for _ in range(3):
        run()
    ctx.sync()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        run()
    e1.record(stream)
    ctx.sync()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


VARIANTS = [("proto1", 1, 0, 0, 0), ("proto3", 3, 0, 0, 0), ("proto2", 2, 0, 0, 0), ("proto3", 3, 0, 0, 0), ("proto1", 1, 0, 0, 0)]
tune(4, 40)  # a heartbeat that cannot be stopped in stream order costs at most 40 ms per run here
for name, proto, hb, sleep_ns, mode in VARIANTS:
    tune(0, proto)
    tune(2, sleep_ns)
    tune(3, mode)
    tune(1, hb)
    dist.barrier()
    t0, t8 = measure(0), measure(min(8, nl))
    hl.dist_check(ctx)
    if rank == 0:
        print(f"SHARD_TUNE world={world} n={n} {name}: R=0 {t0:.4f} ms, R={min(8, nl)} {t8:.4f} ms, "
              f"per sharded round {1e3 * (t8 - t0) / min(8, nl):.1f} us", flush=True)
dist.barrier()
ctx.close()
dist.destroy_process_group()
