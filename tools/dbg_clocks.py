import os, sys
os.environ["B200_DEBUG_CLOCKS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C, numpy as np
import halo2_lasso_b200 as hl
from bench import rand_canonical
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ctx = hl.Context(0)
polys = []
for seed in (1, 2):
    p = hl.MultilinearPolynomial.new(ctx, rand_canonical(seed, 1 << n))
    polys.append(p)
y = rand_canonical(3, n)
one = np.array([0xAC96341C4FFFFFFB, 0x36FC76959F60CD29, 0x666EA36F7879462E, 0x0E0A77C19A07DF2F], dtype=np.uint64)
for _ in range(3):
    hl.Keccak256Transcript(ctx)
    hl.ClassicSumCheck.prove_evals(ctx, n, polys, one.reshape(1, 4), y, one)
out = (C.c_longlong * 512)()
hl._chk(hl.lib().b200_debug_clocks(ctx.h, out), "dbg")
names = ["start", "loop", "reduce1", "partial", "ticket", "sum2", "tr_in", "canon", "absorb4", "squeeze", "interp", "end"]
for r in range(n):
    st = [out[r * 16 + i] for i in range(12)]
    if st[0] == 0: continue
    print(r, " ".join(f"{names[i]}={st[i]-st[i-1]}" for i in range(1, 12)), "total", st[11] - st[0])
