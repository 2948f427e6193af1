"""One full Lasso proof — the command profiled with ncu (profiles/).
prof_lasso.py MU [REPS] [KIND(range|and|xor)] [CHUNKS]   (default: cfg3 = range, 4 chunks)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import halo2_lasso_b200 as hl
from bench import rand_canonical, rand_u64s

mu = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kind = {"range": 0, "and": 1, "xor": 2}[sys.argv[3]] if len(sys.argv) > 3 else 0
chunks = int(sys.argv[4]) if len(sys.argv) > 4 else 4
ctx = hl.Context(0)
raw = np.concatenate([rand_canonical(7, mu), np.zeros((32 - mu, 4), dtype=np.uint64)])
p = hl.MultilinearPolynomial.new(ctx, raw)
hl._chk(hl.lib().b200_fr_convert(ctx.h, p.dev, p.dev, C.c_uint64(32), C.c_int(1)), "cv")
kzg = hl.MultilinearKzg.setup(ctx, p.evals()[:max(mu, 16)] if mu >= 16 else p.evals()[:16])
prover = hl.LassoProver(ctx, kzg, kind, chunks)
bits = (16 if kind == 0 else 8) * chunks
mask = np.uint64((1 << bits) - 1) if bits < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
xs = rand_u64s(5, 1 << mu) & mask
ys = (rand_u64s(6, 1 << mu) & mask) if kind else None
for _ in range(reps):
    tr = hl.Keccak256Transcript(ctx)
    prover.prove(xs, ys)
print("proof bytes", len(tr.into_proof()), "launches", ctx.launch_count())
