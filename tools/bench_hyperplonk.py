"""HyperPlonk prove on the reference's proof_system bench shape (benchmark/benches/proof_system.rs:47-77:
vanilla_plonk / vanilla_plonk_with_lookup circuits, MultilinearKzg<Bn256>).
Usage: bench_hyperplonk.py K [--lookup] [--oracle]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import halo2_lasso_b200 as hl
from halo2_lasso_b200 import hyperplonk as H
from bench import rand_canonical

k = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ctx = hl.Context(0)
ss = H.ints_to_mont(ctx, [int(sum(int(v[j]) << (64 * j) for j in range(4))) for v in rand_canonical(7, k)])
kzg = hl.MultilinearKzg.setup(ctx, ss)
t0 = time.time()
LOOKUP = "--lookup" in sys.argv
info, instances, w = (H.rand_vanilla_plonk_with_lookup_circuit if LOOKUP else H.rand_vanilla_plonk_circuit)(k, 1)
gen_s = time.time() - t0
hp = H.HyperPlonk(ctx, kzg, info)
wit = [H.upload_ints(ctx, c) for c in w]
stream = torch.cuda.ExternalStream(ctx.stream)
times = []
for it in range(7):
    tr = hl.Keccak256Transcript(ctx)
    ctx.sync()
    t0 = time.perf_counter()
    hp.prove(instances, witness_polys=wit)
    ctx.sync()
    if it >= 2:
        times.append(1e3 * (time.perf_counter() - t0))
out = {"bench": "HyperPlonk prove, " + ("vanilla_plonk_with_lookup (19 polys, permutation + LogUp lookup argument, degree 5)"
                if LOOKUP else "vanilla_plonk (13 polys, permutation argument, no lookups)") + ", MultilinearKzg", "k": k,
       "gpu_ms_wall": sorted(times)[len(times) // 2], "gpu_ms_all": [round(t, 2) for t in times], "proof_bytes": len(tr.into_proof()), "fixture_gen_s": round(gen_s, 1)}
if "--oracle" in sys.argv:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    from halo2_lasso_b200.expression import compose
    okzg = O.Kzg.from_eqs(ss, [kzg.eqs(i) for i in range(k + 1)])
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys, lookups=info.lookups)
    ohp = O.HyperPlonk(okzg, k, expr, len(instances), 3, [O.fr_from_ints(p) for p in info.preprocess_polys],
                       info.permutation_polys, info.permutations, nz, lookups=info.lookups)
    to = O.Transcript()
    t0 = time.perf_counter()
    ohp.prove(to, O.fr_from_ints(instances), [O.fr_from_ints(c) for c in w])
    out["oracle_ms"] = 1e3 * (time.perf_counter() - t0)
    out["oracle_note"] = "naive tree-walking generic evaluator, single-threaded sum-check: NOT a performance baseline"
    out["byte_identical"] = to.proof() == tr.into_proof()
print(json.dumps(out))
