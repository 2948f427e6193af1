"""Per-round launch times (CUDA events on the library stream) of one cfg2 sum-check; T terms optional."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ctypes as C
import halo2_lasso_b200 as hl
from bench import rand_canonical

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ctx = hl.Context(0)
polys = []
for seed in range(2 * T):
    p = hl.MultilinearPolynomial.new(ctx, rand_canonical(seed + 1, 1 << n))
    hl._chk(hl.lib().b200_fr_convert(ctx.h, p.dev, p.dev, C.c_uint64(1 << n), C.c_int(1)), "cv")
    polys.append(p)
y = rand_canonical(99, n)
one = np.array([0xAC96341C4FFFFFFB, 0x36FC76959F60CD29, 0x666EA36F7879462E, 0x0E0A77C19A07DF2F], dtype=np.uint64)
w = np.tile(one, (T, 1))


def run():
    hl.Keccak256Transcript(ctx)
    hl.ClassicSumCheck.prove_evals(ctx, n, polys, w, y, one)


for _ in range(3):
    run()
best = None
for _ in range(5):
    r = hl.profile_rounds(ctx, run, n, 3)
    if best is None or sum(r["round_ms"]) < sum(best["round_ms"]):
        best = r
print("n", n, "T", T, "total_ms %.4f" % sum(best["round_ms"]))
print("round_us", [round(1e3 * x, 1) for x in best["round_ms"]])
