"""In-process rank groups (b200_dist_init_local) on ONE GPU: wall time of a sharded sum-check per number of sharded
rounds, for worlds 2 / 4 / 8 — what a collective costs when the ranks are contexts of one process."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import halo2_lasso_b200 as hl
from bench import rand_canonical, mont_one

one = mont_one()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for world in (2, 4, 8):
    g = world.bit_length() - 1
    t0 = time.time()
    ctxs = [hl.Context(0) for _ in range(world)]
    hl.dist_init_local(ctxs)
    t_init = time.time() - t0
    y = rand_canonical(3, n)
    tabs = [rand_canonical(s, 1 << n) for s in (1, 2)]
    polys = [[hl.MultilinearPolynomial.new(c, hl.shard_window_slice(t, n, n - g, r, world)) for t in tabs] for r, c in enumerate(ctxs)]
    res = {}
    for R in (0, 4, 8):
        def run(rank, ctx):
            for _ in range(5):
                hl.Keccak256Transcript(ctx)
                hl.sumcheck_prove_evals_sharded(ctx, n, polys[rank], one.reshape(1, 4), y, one, sharded_rounds=R)
            hl.dist_check(ctx)
        hl.run_ranks(ctxs, run)
        t0 = time.time()
        hl.run_ranks(ctxs, run)
        res[R] = round(1e3 * (time.time() - t0) / 5, 2)
    print(f"LOCAL_RANKS world={world} n={n} init_s={t_init:.2f} wall_ms_per_sumcheck_by_sharded_rounds={res}", flush=True)
    for c in ctxs:
        c.close()
