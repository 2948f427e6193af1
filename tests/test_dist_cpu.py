"""gloo tests (CPU, world sizes 2 and 4) of the multi-GPU path: the host-side plumbing (handle exchange in rank order,
hypercube slice arithmetic) and the sharded sum-check / point-sharded MSM PROTOCOLS restated with the oracle's field
arithmetic and real gloo collectives, byte-identical to the single-process oracle."""
import os
import sys

import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import halo2_lasso_b200 as hl

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = bytes([rank]) * 64
    hs = hl.exchange_handles(mine, world)
    ok = hs == [bytes([r]) * 64 for r in range(world)]
    lo, hi = hl.shard_slice(10, rank, world)
    ok = ok and (lo, hi) == (rank * (1024 // world), (rank + 1) * (1024 // world))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_handle_exchange_and_slices_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, 29611, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_shard_slice_covers_the_hypercube():
    sys.path.insert(0, ROOT)
    import halo2_lasso_b200 as hl

    for world in (1, 2, 4, 8):
        spans = [hl.shard_slice(12, r, world) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == 1 << 12
        assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))


# ---- the sharded sum-check PROTOCOL (csrc/shard.cu) restated over gloo ---------------------------
# Each rank owns the slice [rank 2^n / G, (rank + 1) 2^n / G) of every table (top log2 G variables fixed to the bits
# of the rank), builds its eq slice as eq(y[:n_loc]) times the eq factor of the fixed top variables, and per round
# all-gathers only the D partial evaluations; after n_loc rounds the one remaining value per table and rank is
# all-gathered and the last log2 G rounds run redundantly. Field arithmetic = the oracle's vector ops; the collective
# is a real gloo all_gather between two processes. The transcript must equal the single-process oracle's.
#
# Index-WINDOW layout (the fully sharded Lasso prover, shard.cu): rank = index bits [p, p + g); the local tables are the
# compact slices (hl.shard_window_slice), the local eq point is y[0..p) ++ y[p+g..n) with the eq factor of the window
# bits, only the first SR <= p rounds are exchanged, then the bound tables are all-gathered in natural index order
# (full index = ((hi * G + rank) << (p - SR)) | lo) and the remaining n - SR rounds run replicated. p = n - g, SR = n - g
# is the top-variable layout above.
def _sharded_sumcheck_model(O, dist, rank, world, n, tabs, w, y, claim, NP, p=None, SR=None):
    import numpy as np
    import torch

    import halo2_lasso_b200 as hl

    R = O.R_MOD
    g = world.bit_length() - 1
    n_loc, T, D = n - g, len(tabs) // NP, NP + 1
    p = n_loc if p is None else p
    SR = p if SR is None else SR
    assert 0 <= SR <= p <= n_loc
    one = O.fr_from_ints([1])[0]

    def fsum(v):  # Σ of a (k, 4) vector of field elements by pairwise halving
        v = np.ascontiguousarray(v)
        while v.shape[0] > 1:
            if v.shape[0] & 1:
                v = np.concatenate([v, O.fr_from_ints([0])])
            h = v.shape[0] // 2
            v = O.field_op("add", v[:h], v[h:])
        return v[0]

    def bcast(s, k):
        return np.ascontiguousarray(np.broadcast_to(s, (k, 4)))

    def all_gather(v):  # (k, 4) uint64 per rank -> list over ranks
        t = torch.from_numpy(np.ascontiguousarray(v).view(np.int64).copy())
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [o.numpy().view(np.uint64).reshape(-1, 4) for o in out]

    def round_partials(eq, tables):  # evaluations at x = 1..D summed over this rank's pairs (eval.rs:102-131)
        pairs = eq.shape[0] // 2
        e0, e1 = eq[0::2], eq[1::2]
        lo_hi = [(t[0::2], t[1::2]) for t in tables]
        es = O.field_op("sub", e1, e0)
        steps = [O.field_op("sub", b, a) for a, b in lo_hi]
        cur_e, cur = e1.copy(), [b.copy() for _, b in lo_hi]
        out = []
        for x in range(D):
            acc = O.fr_from_ints([0] * pairs)
            for t in range(T):
                prod = cur[t * NP]
                for k in range(1, NP):
                    prod = O.field_op("mul", prod, cur[t * NP + k])
                acc = O.field_op("add", acc, O.field_op("mul", prod, bcast(w[t], pairs)))
            out.append(fsum(O.field_op("mul", acc, cur_e)))
            cur_e = O.field_op("add", cur_e, es)
            cur = [O.field_op("add", c, s) for c, s in zip(cur, steps)]
        return np.stack(out)

    def interpolate(msg, r):  # barycentric over the points 0..D (arithmetic.rs:108-136), on canonical ints
        vals, rr = O.fr_to_ints(msg), O.fr_to_ints(r.reshape(1, 4))[0]
        acc = 0
        for i, v in enumerate(vals):
            num, den = 1, 1
            for j in range(len(vals)):
                if j != i:
                    num, den = num * (rr - j) % R, den * (i - j) % R
            acc = (acc + v * num * pow(den, -1, R)) % R
        return O.fr_from_ints([acc])[0]

    def bind(t, r):
        return O.fix_var(t, r) if t.shape[0] > 1 else t

    tr = O.Transcript()
    # local state: eq slice = eq(y_loc) * Π_j (rank_j ? y_(p+j) : 1 - y_(p+j)),  y_loc = y[0..p) ++ y[p+g..n)
    factor = one
    for j in range(g):
        yj = y[p + j]
        factor = O.field_op("mul", factor.reshape(1, 4), (yj if (rank >> j) & 1 else O.field_op("sub", one.reshape(1, 4), yj.reshape(1, 4))[0]).reshape(1, 4))[0]
    y_loc = np.concatenate([y[:p], y[p + g:]])
    eq = O.field_op("mul", O.eq_xy(y_loc), bcast(factor, 1 << n_loc))
    tables = [hl.shard_window_slice(t, n, p, rank, world).copy() for t in tabs]
    challenges = []
    for rnd in range(n):
        if rnd == SR:  # all-gather of the (bound) local tables, natural index order: ((hi * G + rank) << q) | lo
            q = p - SR
            gathered = all_gather(np.stack(tables + [eq]).reshape(-1, 4))
            k = len(tabs) + 1
            stacked = np.stack(gathered).reshape(world, k, -1, 1 << q, 4)  # (rank, table, hi, lo, limbs)
            full = np.ascontiguousarray(stacked.transpose(1, 2, 0, 3, 4)).reshape(k, -1, 4)  # (table, hi, rank, lo)
            tables = [np.ascontiguousarray(full[i]) for i in range(len(tabs))]
            eq = np.ascontiguousarray(full[len(tabs)])
        part = round_partials(eq, tables)
        if rnd < SR:  # the only data that crosses ranks in a sharded round: D field elements
            parts = all_gather(part)
            total = parts[0]
            for other in parts[1:]:
                total = O.field_op("add", total, other)
        else:
            total = part
        p0 = O.field_op("sub", claim.reshape(1, 4), total[0:1])[0]  # p(0) = sum - p(1), eval.rs:129
        msg = np.concatenate([p0.reshape(1, 4), total])
        for m in msg:
            tr.write_fe(m)
        r = tr.squeeze()
        challenges.append(r)
        claim = interpolate(msg, r)
        eq = bind(eq, r)
        tables = [bind(t, r) for t in tables]
    return tr.proof(), np.stack(challenges), np.stack([t[0] for t in tables])


def _protocol_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist

    import oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    O.set_num_threads(1)
    ok = True
    for n, T, NP in ((6, 2, 2), (5, 3, 1)):
        tabs = [O.rand_fr(900 + n + i, 1 << n) for i in range(T * NP)]
        w, y, claim = O.rand_fr(910 + n, T), O.rand_fr(920 + n, n), O.rand_fr(930 + n, 1)[0]
        to = O.Transcript()
        terms = [(w[t], list(range(t * NP, (t + 1) * NP))) for t in range(T)]
        ch_o, ev_o = O.sumcheck_prove_evals(to, n, tabs, y, terms, claim)
        proof, ch, ev = _sharded_sumcheck_model(O, dist, rank, world, n, tabs, w, y, claim, NP)
        ok = ok and proof == to.proof() and (ch == ch_o).all() and (ev == ev_o).all()
        # the index-window layouts of the fully sharded Lasso prover: windows in the middle / at the bottom, fewer sharded
        # rounds than the window allows, and none at all (pure all-gather)
        g = world.bit_length() - 1
        for p, SR in ((2, 2), (2, 1), (1, 0), (0, 0), (n - g, 1)):
            if p > n - g:
                continue
            proof, ch, ev = _sharded_sumcheck_model(O, dist, rank, world, n, tabs, w, y, claim, NP, p=p, SR=SR)
            ok = ok and proof == to.proof() and (ch == ch_o).all() and (ev == ev_o).all()
    # point-sharded MSM (shard.cu msm_sharded / msm_batch_dist): partial commitments all-gathered and added in rank order
    import numpy as np
    import torch

    kz = O.Kzg(O.rand_fr(7, 6))
    bases, sc = kz.eqs(6), O.rand_fr(940, 64)
    size = 64 // world
    part = O.msm(sc[rank * size:(rank + 1) * size], bases[rank * size:(rank + 1) * size])
    t = torch.from_numpy(np.ascontiguousarray(part).view(np.int64).copy())
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    acc = out[0].numpy().view(np.uint64)
    for o in out[1:]:
        acc = O.g1_add(acc, o.numpy().view(np.uint64))
    ok = ok and (acc == O.msm(sc, bases)).all()
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def _run_protocol(world, port):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_protocol_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == [(r, True) for r in range(world)]


def test_sharded_sumcheck_protocol_over_gloo_world2():
    _run_protocol(2, 29612)


def test_sharded_sumcheck_protocol_over_gloo_world4():
    _run_protocol(4, 29614)
