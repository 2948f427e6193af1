"""One cfg2 sum-check (n=20) on resident tables — the command profiled with ncu (profiles/)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ctypes as C
import halo2_lasso_b200 as hl
from bench import rand_canonical

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = hl.Context(0)
polys = []
for seed in (1, 2):
    p = hl.MultilinearPolynomial.new(ctx, rand_canonical(seed, 1 << n))
    hl._chk(hl.lib().b200_fr_convert(ctx.h, p.dev, p.dev, C.c_uint64(1 << n), C.c_int(1)), "cv")
    polys.append(p)
y = rand_canonical(3, n)
one = np.array([0xAC96341C4FFFFFFB, 0x36FC76959F60CD29, 0x666EA36F7879462E, 0x0E0A77C19A07DF2F], dtype=np.uint64)
for _ in range(reps):
    tr = hl.Keccak256Transcript(ctx)
    hl.ClassicSumCheck.prove_evals(ctx, n, polys, one.reshape(1, 4), y, one)
ctx.sync()
print("done", ctx.launch_count())
