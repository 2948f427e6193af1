"""cProfile of the host side of HyperPlonk.prove (where does wall time beyond the kernels go?)."""
import os, sys, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import halo2_lasso_b200 as hl
from halo2_lasso_b200 import hyperplonk as H
from bench import rand_canonical

k = int(sys.argv[1]) if len(sys.argv) > 1 else 18
LOOKUP = "--lookup" in sys.argv
ctx = hl.Context(0)
ss = H.ints_to_mont(ctx, [int(sum(int(v[j]) << (64 * j) for j in range(4))) for v in rand_canonical(7, k)])
kzg = hl.MultilinearKzg.setup(ctx, ss)
info, instances, w = (H.rand_vanilla_plonk_with_lookup_circuit if LOOKUP else H.rand_vanilla_plonk_circuit)(k, 1)
hp = H.HyperPlonk(ctx, kzg, info)
wit = [H.upload_ints(ctx, c) for c in w]
for _ in range(2):
    hl.Keccak256Transcript(ctx)
    hp.prove(instances, witness_polys=wit)
ctx.sync()
pr = cProfile.Profile()
pr.enable()
hl.Keccak256Transcript(ctx)
hp.prove(instances, witness_polys=wit)
ctx.sync()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
