"""Debug helper: the fully sharded Lasso prover with all ranks on ONE GPU (in-process rank group), several
configurations in a row, reporting per rank: error / proof equal to the oracle. B200_PEER_DEBUG=1 prints which wait
timed out first.   python tools/dbg_local_shard.py [world]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("B200_PEER_TIMEOUT_S", "1")
os.environ.setdefault("B200_ARENA_MB", "64")
os.environ.setdefault("B200_PEER_DEBUG", "1")

import numpy as np  # noqa: E402

import halo2_lasso_b200 as hl  # noqa: E402
import oracle as O  # noqa: E402
from test_gpu_sharded_local import lasso_operands  # noqa: E402

NV = 16
world = int(sys.argv[1]) if len(sys.argv) > 1 else 4
okzg = O.Kzg(O.rand_fr(7, NV))
srs = [okzg.eqs(k) for k in range(NV + 1)]
ctxs = [hl.Context(0) for _ in range(world)]
hl.dist_init_local(ctxs)
kzgs = [hl.MultilinearKzg(c, srs) for c in ctxs]

CASES = [(O.TABLE_XOR, 2, 13, 10, 1 << 14), (O.TABLE_XOR, 2, 13, 10, 64), (O.TABLE_XOR, 2, 12, 10, 1 << 14),
         (O.TABLE_RANGE, 2, 13, 10, 1 << 14), (O.TABLE_XOR, 4, 13, 10, 1 << 14), (O.TABLE_XOR, 2, 13, 12, 1 << 14),
         (O.TABLE_RANGE, 2, 14, 13, 256), (O.TABLE_XOR, 2, 13, 10, 1 << 14)]
REPEAT = int(sys.argv[2]) if len(sys.argv) > 2 else 1
for kind, chunks, mu, k0, min_items in CASES * REPEAT:
    xs, ys = lasso_operands(kind, chunks, mu, 8400 + mu)
    to = O.Transcript()
    assert O.lasso_prove(okzg, to, kind, chunks, mu, xs, ys)
    res = [None] * world

    def run(rank, ctx):
        hl.dist_shard_lasso(ctx, k0)
        hl.dist_shard_min_items(ctx, min_items)
        tr = hl.Keccak256Transcript(ctx)
        status = "ok"
        proof = b""
        try:
            hl.LassoProver(ctx, kzgs[rank], kind, chunks).prove(xs, ys)
            proof = tr.into_proof()
        except Exception as e:  # noqa: BLE001
            status = f"prove: {e!r}"
        try:
            hl.dist_check(ctx)
        except Exception as e:  # noqa: BLE001
            status += f" | check: {e!r}"
        hl.dist_shard_lasso(ctx, 0)
        hl.dist_shard_min_items(ctx, 1 << 16)
        res[rank] = (status, proof == to.proof())
        return None

    t0 = time.time()
    try:
        hl.run_ranks(ctxs, run)
    except Exception as e:  # noqa: BLE001
        print("run_ranks:", repr(e)[:300])
    print(f"CASE {'OK  ' if all(r is not None and r[0] == 'ok' and r[1] for r in res) else 'FAIL'} kind={kind} c={chunks} mu={mu} k0={k0} min_items={min_items} world={world} {time.time() - t0:.1f}s:",
          res, flush=True)
