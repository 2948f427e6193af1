"""Independent pure-Python model of `HyperPlonk::prove` (pb/backend/hyperplonk.rs:164-291, prover.rs:32-409) for the
reference's test circuits, built on pymodel.py (big ints, own Keccak, affine curve arithmetic, naive MSM). Every
expression is evaluated by walking the tree (no compiler, no evaluator cache); tables are materialised as plain lists.
Only usable at toy sizes (k <= 4). Returns the proof stream, which tests compare with the oracle's byte for byte."""
import pymodel as M
from halo2_lasso_b200 import hyperplonk as H
from halo2_lasso_b200.expression import BooleanHypercube, compose

R = M.R


def eval_tree(n, leaf, ch):
    k = n[0]
    if k == "const":
        return n[1]
    if k == "chal":
        return ch[n[1]]
    if k in ("identity", "lagrange", "eq", "poly"):
        return leaf(n)
    if k == "neg":
        return -eval_tree(n[1], leaf, ch) % R
    if k == "sum":
        return (eval_tree(n[1], leaf, ch) + eval_tree(n[2], leaf, ch)) % R
    if k == "prod":
        return eval_tree(n[1], leaf, ch) * eval_tree(n[2], leaf, ch) % R
    if k == "scaled":
        return eval_tree(n[1], leaf, ch) * n[2] % R
    base = eval_tree(n[2], leaf, ch)
    acc, pw = eval_tree(n[1][0], leaf, ch), base
    for c in n[1][1:]:
        acc, pw = (acc + pw * eval_tree(c, leaf, ch)) % R, pw * base % R
    return acc


def prove(srs, info, instances, witness, max_degree=4):
    k = info.k
    N = 1 << k
    bh = BooleanHypercube(k)
    order = bh.iter()
    tr = M.Transcript()
    # one instance column given as a flat list, or one list per column (pb/backend.rs:50-51)
    inst_cols = instances if isinstance(info.num_instances, list) else [instances]
    inst_polys = []
    for col in inst_cols:
        for v in col:
            tr.common_fe(v)
    for col in inst_cols:
        poly = [0] * N
        for i, v in enumerate(col):
            poly[order[i + 1]] = v
        inst_polys.append(poly)
    # witness phases (hyperplonk.rs:183-204): `witness` is the list of columns, or synthesize(round, challenges)
    phases = info.num_witness_polys if isinstance(info.num_witness_polys, list) else [info.num_witness_polys]
    phase_challenges = getattr(info, "num_challenges", [0] * len(phases))
    circuit_ch, witness_cols = [], []
    for rnd, (nw, nc) in enumerate(zip(phases, phase_challenges)):
        cols = witness(rnd, list(circuit_ch)) if callable(witness) else witness
        assert len(cols) == nw
        for col in cols:
            tr.write_comm(M.kzg_commit(srs, col))
        witness_cols += [list(c) for c in cols]
        circuit_ch += [tr.squeeze() for _ in range(nc)]
    polys = inst_polys + [list(p) for p in info.preprocess_polys] + witness_cols

    def row_leaf(b):
        def leaf(n):
            if n[0] == "poly":
                return polys[n[1]][bh.rotate(b, n[2])]
            if n[0] == "identity":
                return b
            if n[0] == "lagrange":
                return 1 if b == order[n[1] % N] else 0
            raise ValueError(n)
        return leaf

    # LogUp (prover.rs:50-250)
    beta = tr.squeeze()
    compressed, ms = [], []
    for lookup in info.lookups:
        ci, ct = [0] * N, [0] * N
        for b in range(N):
            pw = 1
            for inp, tab in lookup:
                ci[b] = (ci[b] + pw * eval_tree(inp.node, row_leaf(b), circuit_ch)) % R
                ct[b] = (ct[b] + pw * eval_tree(tab.node, row_leaf(b), circuit_ch)) % R
                pw = pw * beta % R
        last = {v: i for i, v in enumerate(ct)}
        m = [0] * N
        for v in ci:
            m[last[v]] += 1
        compressed.append((ci, ct))
        ms.append(m)
    for m in ms:
        tr.write_comm(M.kzg_commit(srs, m))
    gamma = tr.squeeze()
    hs = [[(pow((gamma + a) % R, -1, R) - mm * pow((gamma + t) % R, -1, R)) % R for a, t, mm in zip(ci, ct, m)]
          for (ci, ct), m in zip(compressed, ms)]
    # permutation_z_polys (prover.rs:252-345)
    sigmas = H.permutation_polys(k, info.permutation_polys, info.permutations)
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys, num_challenges=len(circuit_ch),
                       max_degree=max_degree, lookups=info.lookups)
    chunk = -(-len(info.permutation_polys) // nz) if nz else 0
    products = []
    for c in range(nz):
        cols = list(range(c * chunk, min(len(info.permutation_polys), (c + 1) * chunk)))
        prod = []
        for b in range(N):
            num = den = 1
            for i in cols:
                wv = polys[info.permutation_polys[i]][b]
                num = num * (wv + beta * ((i << k) + b) + gamma) % R
                den = den * (wv + beta * sigmas[i][b] + gamma) % R
            prod.append(num * pow(den, -1, R) % R)
        products.append(prod)
    flat = [0] * (nz * N)
    if nz:
        flat[nz] = 1
        state, pos = 1, nz + 1
        for kk in range(1, N):
            for c in range(nz):
                if pos >= len(flat):
                    break
                state = state * products[c][order[kk]] % R
                flat[pos] = state
                pos += 1
    nth = bh.nth_map()
    zs = [[flat[c + nz * nth[b]] for b in range(N)] for c in range(nz)]
    for p in hs + zs:
        tr.write_comm(M.kzg_commit(srs, p))
    alpha = tr.squeeze()
    y = [tr.squeeze() for _ in range(k)]
    polys = polys + sigmas + ms + hs + zs
    ch = circuit_ch + [beta, gamma, alpha]
    # zero check: EvaluationsProver over materialised leaf tables (eval.rs:92-131), claimed sum 0
    leaves = expr.leaves()
    tabs = {}
    for l in leaves:
        if l[0] == "poly":
            tabs[l] = [polys[l[1]][bh.rotate(b, l[2])] for b in range(N)]
        elif l[0] == "eq":
            tabs[l] = M.eq_xy(y)
        elif l[0] == "identity":
            tabs[l] = list(range(N))
        else:
            tabs[l] = [1 if b == order[l[1] % N] else 0 for b in range(N)]
    bound = {p: list(polys[p]) for p in range(len(polys))}
    d = expr.degree()
    claim, x = 0, []
    for _ in range(k):
        ev = [0] * (d + 1)
        size = len(next(iter(tabs.values())))
        for b in range(size // 2):
            for xx in range(1, d + 1):
                ev[xx] = (ev[xx] + eval_tree(expr.node, lambda l: (tabs[l][2 * b] + xx * (tabs[l][2 * b + 1] - tabs[l][2 * b])) % R, ch)) % R
        ev[0] = (claim - ev[1]) % R
        for e in ev:
            tr.write_fe(e)
        r = tr.squeeze()
        x.append(r)
        claim = M.interpolate(ev, r)
        tabs = {l: M.fix_var(t, r) for l, t in tabs.items()}
        bound = {p: M.fix_var(t, r) for p, t in bound.items()}
    # evaluations in pcs_query order (verifier.rs:147-182), rotated ones at the rotation_eval_points
    queries = sorted({(l[1], l[2]) for l in leaves if l[0] == "poly" and l[1] >= len(inst_cols)})
    rotations = sorted({r for _, r in queries})
    points, offset = [], {}
    for r in rotations:
        offset[r] = len(points)
        points += H.rotation_eval_points(x, r)
    evals = []
    for p, r in queries:
        if r == 0:
            evals.append((p, offset[0], bound[p][0]))
        else:
            for j in range(1 << abs(r)):
                evals.append((p, offset[r] + j, M.evaluate(polys[p], points[offset[r] + j])))
    for _, _, v in evals:
        tr.write_fe(v)
    M.kzg_batch_open(srs, tr, k, polys, points, evals)
    return tr.stream
