"""The reference's own `zero_check` criterion bench shape (plonkish_backend/benches/zero_check.rs:16-42):
ClassicSumCheck<EvaluationsProver> on vanilla_plonk_expression, claimed sum 0, n = 20 (..23 with --n)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
import halo2_lasso_b200 as hl
from halo2_lasso_b200.expression import vanilla_plonk_expression
from bench import rand_canonical

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
ctx = hl.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream)
expr = vanilla_plonk_expression(n)
polys = []
for i in range(13):
    p = hl.MultilinearPolynomial.new(ctx, rand_canonical(100 + i, 1 << n))
    hl._chk(hl.lib().b200_fr_convert(ctx.h, p.dev, p.dev, C.c_uint64(1 << n), C.c_int(1)), "cv")
    polys.append(p)
y = polys[0].evals()[:n].copy()
ch = [int(x) for x in rand_canonical(7, 3)[:, 0]]
zero = np.zeros(4, dtype=np.uint64)
times = []
for it in range(6):
    hl.Keccak256Transcript(ctx)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record(stream)
    hl.prove_expression(ctx, n, expr, polys, ch, [y], zero)
    s1.record(stream)
    ctx.sync(); torch.cuda.synchronize()
    if it >= 2:
        times.append(s0.elapsed_time(s1))
print(json.dumps({"bench": "zero_check (vanilla_plonk_expression, 17 tables, degree 5)", "num_vars": n,
                  "ms": sum(times) / len(times), "samples": len(times)}))
prof = hl.profile(ctx, lambda: (hl.Keccak256Transcript(ctx), hl.prove_expression(ctx, n, expr, polys, ch, [y], zero)))
print("round_ms", [round(t, 3) for tag, t in prof if tag < 1000])
