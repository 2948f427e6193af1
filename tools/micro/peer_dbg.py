"""clock64 stamps inside the sharded round kernels (single-CTA rounds): where does a collective spend its time?"""
import os, sys
os.environ["B200_DEBUG_CLOCKS"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import ctypes as C
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
g = world.bit_length() - 1
n_loc = 8
n_tot = n_loc + g
one = mont_one()
polys = [hl.MultilinearPolynomial.new(ctx, rand_canonical(s + 10 * rank, 1 << n_loc)) for s in (1, 2)]
y = rand_canonical(3, n_tot)
for _ in range(6):
    hl.Keccak256Transcript(ctx)
    hl.sumcheck_prove_evals_sharded(ctx, n_tot, polys, one.reshape(1, 4), y, one)
out = (C.c_longlong * 512)()
hl._chk(hl.lib().b200_debug_clocks(ctx.h, out), "dbg")
names = ["start", "loop", "reduce1", "partial", "ticket", "sum2", "tr_in", "xchg+canon", "absorb4", "squeeze", "interp", "end"]
dist.barrier()
for r_ in range(world):
    if r_ == rank:
        for r in range(n_loc):
            st = [out[r * 16 + i] for i in range(12)]
            if st[0] == 0:
                continue
            print(f"rank {rank} round {r}:", " ".join(f"{names[i]}={(st[i]-st[i-1])/1965:.1f}" for i in range(1, 12)),
                  f"total {(st[11] - st[0])/1965:.1f} us | wall clock: xchg {(out[r*16+14]-out[r*16+13])/1e3:.1f} us, tail {(out[r*16+15]-out[r*16+14])/1e3:.1f} us, total {(out[r*16+15]-out[r*16+12])/1e3:.1f} us", flush=True)
    dist.barrier()
dist.destroy_process_group()
