"""Independent pure-Python model of the Lasso / Surge prover specified in DESIGN.md §4, built on pymodel.py (big ints,
own Keccak, affine curve arithmetic). Written from the specification text, not from the oracle: witness, commitments,
primary sum-check, fingerprints, ONE batched multi-height grand product, leaf evaluations, two additive batch
openings. The 2^16-entry subtables make the SRS large, so its points are produced lazily (only the bases that meet a
non-zero scalar are ever computed). `make_golden_lasso.py` runs it once and stores the proofs as fixtures."""
import pymodel as M

R = M.R
SUB_VARS = 16
RANGE, AND, XOR = 0, 1, 2


def out_bits(kind):
    return 16 if kind == RANGE else 8


def subtable(kind, x):
    if kind == RANGE:
        return x
    p, q = x >> 8, x & 0xFF
    return (p & q) if kind == AND else (p ^ q)


def dim(kind, x, y, t):
    if kind == RANGE:
        return (x >> (16 * t)) & 0xFFFF
    return (((x >> (8 * t)) & 0xFF) << 8) | ((y >> (8 * t)) & 0xFF)


class LazySrs:
    """MultilinearKzg::setup (kzg.rs:166-213): eqs[k][b] = g1 * Π_j (b_j ? s_j : 1 - s_j), newest variable on the TOP bit"""

    def __init__(self, ss):
        self.scalars = [[1]]
        for s in ss:
            last = self.scalars[-1]
            hi = [s * e % R for e in last]
            self.scalars.append([(e - h) % R for e, h in zip(last, hi)] + hi)
        self.cache = {}

    def point(self, level, i):
        key = (level, i)
        if key not in self.cache:
            self.cache[key] = M.g1_mul(M.G, self.scalars[level][i])
        return self.cache[key]

    def commit(self, poly):
        level = len(poly).bit_length() - 1
        acc = None
        for i, v in enumerate(poly):
            if v:
                acc = M.g1_add(acc, M.g1_mul(self.point(level, i), v))
        return acc


def kzg_open(srs, tr, poly, point):
    """kzg.rs:276-302: quotients top variable first, commitments written in variable order"""
    rem = list(poly)
    comms = [None] * len(point)
    for nv in reversed(range(len(point))):
        half = 1 << nv
        q = [(rem[half + i] - rem[i]) % R for i in range(half)]
        rem = [(rem[i] + (rem[half + i] - rem[i]) * point[nv]) % R for i in range(half)]
        comms[nv] = srs.commit(q)
    for cm in comms:
        tr.write_comm(cm)


def kzg_batch_open(srs, tr, n, polys, points, evals):
    """additive::batch_open (pb/pcs/multilinear.rs:134-235); evals = [(poly, point, value)]"""
    ell = (len(evals) - 1).bit_length()
    if ell == 0:  # eq_xy(&[]) is the zero polynomial (multilinear.rs:92-94): the reference cannot batch a single evaluation
        raise ValueError("batch_open needs at least two evaluations")
    t = [tr.squeeze() for _ in range(ell)]
    eq_xt = M.eq_xy(t)
    merged = [None] * len(points)
    for (pi, qi, _), e in zip(evals, eq_xt):
        contrib = [e * v % R for v in polys[pi]]
        merged[qi] = contrib if merged[qi] is None else [(a + b) % R for a, b in zip(merged[qi], contrib)]
    tilde = sum(v * e for (_, _, v), e in zip(evals, eq_xt)) % R
    chal, _ = M.sumcheck_prove_coeffs(tr, n, merged, [(1, points[i], i) for i in range(len(points))], tilde)
    g = [0] * (1 << n)
    for i, mp in enumerate(merged):
        s = M.eq_xy_eval(chal, points[i])
        g = [(a + s * b) % R for a, b in zip(g, mp)]
    kzg_open(srs, tr, g, chal)


def grand_product(tr, leaves):
    """§4 step 6: ONE batched layered product argument over trees of different heights."""
    T = len(leaves)
    hs = [len(l).bit_length() - 1 for l in leaves]
    layers = []
    for t in range(T):
        lv = {hs[t]: list(leaves[t])}
        for k in range(hs[t] - 1, -1, -1):
            ch, half = lv[k + 1], 1 << k
            lv[k] = [ch[i] * ch[i + half] % R for i in range(half)]
        layers.append(lv)
    claims = [layers[t][0][0] for t in range(T)]
    for c in claims:
        tr.write_fe(c)
    points, y = {}, []
    for k in range(max(hs)):
        half = 1 << k
        act = [t for t in range(T) if hs[t] > k]
        ls = [layers[t][k + 1][:half] for t in act]
        rs = [layers[t][k + 1][half:] for t in act]
        if k == 0:
            x = []
            evals = [v for l, r in zip(ls, rs) for v in (l[0], r[0])]
        else:
            gamma = tr.squeeze()
            terms, claim, pw = [], 0, 1
            for i, t in enumerate(act):
                terms.append((pw, [2 * i, 2 * i + 1]))
                claim = (claim + pw * claims[t]) % R
                pw = pw * gamma % R
            polys = [p for l, r in zip(ls, rs) for p in (l, r)]
            x, evals = M.sumcheck_prove_evals(tr, k, polys, y, terms, claim)
        for e in evals:
            tr.write_fe(e)
        mu = tr.squeeze()
        for i, t in enumerate(act):
            claims[t] = (evals[2 * i] + mu * (evals[2 * i + 1] - evals[2 * i])) % R
        y = list(x) + [mu]
        points[k + 1] = y
    return claims, points


CUSTOM = 3


class CustomTable:
    """DESIGN.md §4 "tables as data": the 2^16 subtable values, 1 or 2 operands of operand_bits bits per chunk,
    g = Σ_t 2^(out_bits t) E_t. The table is part of the statement: (num_operands, operand_bits, out_bits) and the
    Keccak-256 digest of the values as little-endian u32 words (little-endian integer mod r) are absorbed after (3, c, mu)."""

    def __init__(self, num_operands, operand_bits, out_bits, values):
        assert len(values) == 1 << SUB_VARS and num_operands * operand_bits <= SUB_VARS
        self.num_operands, self.operand_bits, self.out_bits, self.values = num_operands, operand_bits, out_bits, values

    def dim(self, x, y, t):
        mask = (1 << self.operand_bits) - 1
        xt, yt = (x >> (self.operand_bits * t)) & mask, (y >> (self.operand_bits * t)) & mask
        return xt if self.num_operands == 1 else (xt << self.operand_bits) | yt

    def digest(self):
        data = b"".join(v.to_bytes(4, "little") for v in self.values)
        return int.from_bytes(M.keccak256(data), "little") % R


def prove(ss, kind, c, mu, xs, ys=None, table=None):
    """-> proof bytes, or None when a commitment is the identity (all-distinct addresses); kind == CUSTOM: `table`"""
    global out_bits, subtable, dim
    if kind == CUSTOM:  # the same protocol over the table's own maps
        saved = (out_bits, subtable, dim)
        out_bits, subtable, dim = (lambda k: table.out_bits), (lambda k, x: table.values[x]), (lambda k, x, y, t: table.dim(x, y, t))
        try:
            return _prove(ss, kind, c, mu, xs, ys, [table.num_operands, table.operand_bits, table.out_bits, table.digest()])
        finally:
            out_bits, subtable, dim = saved
    return _prove(ss, kind, c, mu, xs, ys, [])


def _prove(ss, kind, c, mu, xs, ys, extra_statement):
    m, S = 1 << mu, 1 << SUB_VARS
    srs = LazySrs(ss)
    tr = M.Transcript()
    for v in [kind, c, mu] + extra_statement:
        tr.common_fe(v)
    ys = ys if ys is not None else [0] * m
    dims = [[dim(kind, xs[j], ys[j], t) for j in range(m)] for t in range(c)]
    es = [[subtable(kind, d) for d in dims[t]] for t in range(c)]
    read_ts, final_cts = [], []
    for t in range(c):
        cnt, ts = {}, []
        for d in dims[t]:
            ts.append(cnt.get(d, 0))
            cnt[d] = cnt.get(d, 0) + 1
        read_ts.append(ts)
        final_cts.append([cnt.get(x, 0) for x in range(S)])
    a = [sum(es[t][j] << (out_bits(kind) * t) for t in range(c)) for j in range(m)]
    mpolys = [a] + dims + es + read_ts
    for p in mpolys + final_cts:
        cm = srs.commit(p)
        if cm is None:
            return None
        tr.write_comm(cm)
    r = [tr.squeeze() for _ in range(mu)]
    v_a = M.evaluate(a, r)
    tr.write_fe(v_a)
    x_p, e_p = M.sumcheck_prove_evals(tr, mu, es, r, [(1 << (out_bits(kind) * t), [t]) for t in range(c)], v_a)
    for e in e_p:
        tr.write_fe(e)
    gamma, tau = tr.squeeze(), tr.squeeze()
    g2 = gamma * gamma % R
    mleaves, sleaves = [], []
    for t in range(c):
        rd = [(dims[t][j] * g2 + es[t][j] * gamma + read_ts[t][j] - tau) % R for j in range(m)]
        mleaves += [rd, [(v + 1) % R for v in rd]]
        init = [(x * g2 + subtable(kind, x) * gamma - tau) % R for x in range(S)]
        sleaves += [init, [(v + f) % R for v, f in zip(init, final_cts[t])]]
    _, points = grand_product(tr, mleaves + sleaves)
    x_m, x_s = points[mu], points[SUB_VARS]
    ev_dim = [M.evaluate(p, x_m) for p in dims]
    ev_e = [M.evaluate(p, x_m) for p in es]
    ev_ts = [M.evaluate(p, x_m) for p in read_ts]
    ev_cts = [M.evaluate(p, x_s) for p in final_cts]
    for v in ev_dim + ev_e + ev_ts + ev_cts:
        tr.write_fe(v)
    evs = [(0, 0, v_a)]
    evs += [(1 + c + t, 1, e_p[t]) for t in range(c)]
    evs += [(1 + t, 2, ev_dim[t]) for t in range(c)]
    evs += [(1 + c + t, 2, ev_e[t]) for t in range(c)]
    evs += [(1 + 2 * c + t, 2, ev_ts[t]) for t in range(c)]
    kzg_batch_open(srs, tr, mu, mpolys, [r, x_p, x_m], evs)
    kzg_batch_open(srs, tr, SUB_VARS, final_cts, [x_s], [(t, 0, ev_cts[t]) for t in range(c)])
    return tr.stream
