"""One full Lasso proof (cfg3 by default) — the command profiled with ncu (profiles/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import halo2_lasso_b200 as hl
from bench import rand_canonical, rand_u64s

mu = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = hl.Context(0)
raw = np.concatenate([rand_canonical(7, mu), np.zeros((32 - mu, 4), dtype=np.uint64)])
p = hl.MultilinearPolynomial.new(ctx, raw)
hl._chk(hl.lib().b200_fr_convert(ctx.h, p.dev, p.dev, C.c_uint64(32), C.c_int(1)), "cv")
kzg = hl.MultilinearKzg.setup(ctx, p.evals()[:max(mu, 16)] if mu >= 16 else p.evals()[:16])
prover = hl.LassoProver(ctx, kzg, 0, 4)
xs = rand_u64s(5, 1 << mu)
for _ in range(reps):
    tr = hl.Keccak256Transcript(ctx)
    prover.prove(xs)
print("proof bytes", len(tr.into_proof()), "launches", ctx.launch_count())
