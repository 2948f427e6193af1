"""Independent pure-Python model (big ints, own Keccak) of the reference algorithms on the hot path.

Purpose: pin the C++ oracle. The reference ships no golden vectors and cannot be run
here (Rust, no toolchain), so this third, deliberately naive implementation — written directly from
the reference sources cited below, sharing no code with oracle/ or the CUDA library — generates the
fixtures in tests/golden/*.json (see make_golden.py). Values are plain canonical integers.
"""
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
MASK64 = (1 << 64) - 1


# ---- documented synthetic-input PRNG (stateless splitmix64) ----------------------------------
def sm64(seed, i):
    z = (seed + (i + 1) * 0x9E3779B97F4A7C15) & MASK64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & MASK64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & MASK64
    return z ^ (z >> 31)


def rand_fr(seed, n):
    out = []
    for i in range(n):
        limbs = [sm64(seed, 4 * i + k) for k in range(4)]
        limbs[3] &= 0x1FFFFFFFFFFFFFFF
        out.append(sum(l << (64 * k) for k, l in enumerate(limbs)))
    return out


# ---- Keccak-256 (rate 136, pad 0x01) ----------------------------------------------------------
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B,
       0x0000000080000001, 0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088,
       0x0000000080008009, 0x000000008000000A, 0x000000008000808B, 0x800000000000008B, 0x8000000000008089,
       0x8000000000008003, 0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
       0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & MASK64 if n else x


def keccak_f(A):  # A[x][y]
    for rnd in range(24):
        Cc = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        D = [Cc[(x - 1) % 5] ^ _rol(Cc[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ D[x] for y in range(5)] for x in range(5)]
        B = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                B[y][(2 * x + 3 * y) % 5] = _rol(A[x][y], _ROT[x][y])
        A = [[B[x][y] ^ ((~B[(x + 1) % 5][y]) & MASK64 & B[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        A[0][0] ^= _RC[rnd]
    return A


def keccak256(data: bytes, pad=0x01) -> bytes:
    rate = 136
    msg = bytearray(data)
    msg.append(pad)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    A = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            A[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i: off + 8 * i + 8], "little")
        A = keccak_f(A)
    return b"".join(A[i % 5][i // 5].to_bytes(8, "little") for i in range(4))


# ---- transcript (pb/util/transcript.rs:99-238) -------------------------------------------------
class Transcript:
    def __init__(self):
        self.buf = b""  # bytes absorbed since the last squeeze
        self.stream = b""

    def common_fe(self, v):
        self.buf += v.to_bytes(32, "little")

    def write_fe(self, v):
        self.common_fe(v)
        self.stream += v.to_bytes(32, "big")

    def squeeze(self):
        h = keccak256(self.buf)
        self.buf = h
        return int.from_bytes(h, "little") % R

    def write_comm(self, pt):
        assert pt is not None, "identity has no coordinates"
        self.buf += pt[0].to_bytes(32, "little") + pt[1].to_bytes(32, "little")
        self.stream += pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")


# ---- multilinear (pb/poly/multilinear.rs) --------------------------------------------------------
def eq_xy(y):
    ev = [1]
    for yi in reversed(y):
        nxt = []
        for e in ev:
            hi = e * yi % R
            nxt += [(e - hi) % R, hi]
        ev = nxt
    return ev


def fix_var(p, r):
    return [((p[2 * b + 1] - p[2 * b]) * r + p[2 * b]) % R for b in range(len(p) // 2)]


def evaluate(p, x):
    for xi in x:
        p = fix_var(p, xi)
    return p[0]


def eq_xy_eval(x, y):
    acc = 1
    for a, b in zip(x, y):
        acc = acc * (2 * a * b + 1 - a - b) % R
    return acc


# ---- sum-check (pb/piop/sum_check/classic*.rs) ----------------------------------------------------
def interpolate(evals, r):
    """value at r of the polynomial with evals at 0..d (what barycentric_interpolate returns)."""
    d = len(evals) - 1
    acc = 0
    for i, e in enumerate(evals):
        num, den = 1, 1
        for j in range(d + 1):
            if j != i:
                num = num * (r - j) % R
                den = den * (i - j) % R
        acc = (acc + e * num * pow(den, -1, R)) % R
    return acc


def sumcheck_prove_evals(tr, n, polys, y, terms, claim):
    """F = eq(x,y) * Σ coeff * Π polys[idx]; message p(0..d) with p(0) = claim - p(1) (eval.rs:129)."""
    d = 1 + max(len(t[1]) for t in terms)
    tabs = [list(p) for p in polys]
    eq = eq_xy(y)
    chal = []
    for _ in range(n):
        ev = [0] * (d + 1)
        for b in range(len(eq) // 2):
            for x in range(1, d + 1):
                def at(t):
                    return (t[2 * b] + x * (t[2 * b + 1] - t[2 * b])) % R
                s = 0
                for coeff, idx in terms:
                    p = coeff
                    for i in idx:
                        p = p * at(tabs[i]) % R
                    s += p
                ev[x] = (ev[x] + at(eq) * s) % R
        ev[0] = (claim - ev[1]) % R
        for e in ev:
            tr.write_fe(e)
        r = tr.squeeze()
        chal.append(r)
        claim = interpolate(ev, r)
        eq = fix_var(eq, r)
        tabs = [fix_var(t, r) for t in tabs]
    return chal, [t[0] for t in tabs]


def sumcheck_prove_coeffs(tr, n, polys, prods, claim):
    """Σ scalar * eq(x, y_k) * polys[idx]; message (c0, c1, c2), c1 = claim - 2 c0 - c2 (coeff.rs:136-149)."""
    tabs = [list(p) for p in polys]
    eqs = [eq_xy(p[1]) for p in prods]
    chal = []
    for _ in range(n):
        c0 = c2 = 0
        for (scalar, _, idx), e in zip(prods, eqs):
            t = tabs[idx]
            a0 = sum(e[2 * b] * t[2 * b] for b in range(len(e) // 2)) % R
            a2 = sum((e[2 * b + 1] - e[2 * b]) * (t[2 * b + 1] - t[2 * b]) for b in range(len(e) // 2)) % R
            c0 = (c0 + scalar * a0) % R
            c2 = (c2 + scalar * a2) % R
        c1 = (claim - 2 * c0 - c2) % R
        for cc in (c0, c1, c2):
            tr.write_fe(cc)
        r = tr.squeeze()
        chal.append(r)
        claim = ((c2 * r + c1) * r + c0) % R
        eqs = [fix_var(e, r) for e in eqs]
        tabs = [fix_var(t, r) for t in tabs]
    return chal, [t[0] for t in tabs]


# ---- G1 affine (y^2 = x^3 + 3), None = identity ---------------------------------------------------
G = (1, 2)


def g1_add(P, Q2):
    if P is None:
        return Q2
    if Q2 is None:
        return P
    if P[0] == Q2[0]:
        if (P[1] + Q2[1]) % Q == 0:
            return None
        lam = 3 * P[0] * P[0] * pow(2 * P[1], -1, Q) % Q
    else:
        lam = (Q2[1] - P[1]) * pow(Q2[0] - P[0], -1, Q) % Q
    x = (lam * lam - P[0] - Q2[0]) % Q
    return x, (lam * (P[0] - x) - P[1]) % Q


def g1_mul(P, k):
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, P)
        P = g1_add(P, P)
        k >>= 1
    return acc


def msm(scalars, bases):
    acc = None
    for s, b in zip(scalars, bases):
        acc = g1_add(acc, g1_mul(b, s % R))
    return acc


# ---- MultilinearKzg (pb/pcs/multilinear/kzg.rs, pb/pcs/multilinear.rs) -----------------------------
def kzg_setup(ss):
    eqs = [[1]]
    for s in ss:  # kzg.rs:178-195: newest variable on the TOP bit
        last = eqs[-1]
        hi = [s * e % R for e in last]
        eqs.append([(e - h) % R for e, h in zip(last, hi)] + hi)
    return [[g1_mul(G, e) for e in lvl] for lvl in eqs]


def kzg_commit(srs, poly):
    return msm(poly, srs[len(poly).bit_length() - 1])


def kzg_open(srs, tr, poly, point):
    rem = list(poly)
    comms = [None] * len(point)
    for nv in reversed(range(len(point))):
        half = 1 << nv
        q = [(rem[half + i] - rem[i]) % R for i in range(half)]
        rem = [(rem[i] + (rem[half + i] - rem[i]) * point[nv]) % R for i in range(half)]
        comms[nv] = msm(q, srs[nv])
    for cm in comms:
        tr.write_comm(cm)
    return rem[0]


def kzg_batch_open(srs, tr, n, polys, points, evals):
    """additive::batch_open (pb/pcs/multilinear.rs:134-235); evals = [(poly, point, value)]."""
    ell = (len(evals) - 1).bit_length()
    t = [tr.squeeze() for _ in range(ell)]
    eq_xt = eq_xy(t)
    merged = [None] * len(points)
    for (pi, qi, _), e in zip(evals, eq_xt):
        contrib = [e * v % R for v in polys[pi]]
        merged[qi] = contrib if merged[qi] is None else [(a + b) % R for a, b in zip(merged[qi], contrib)]
    tilde = sum(v * e for (_, _, v), e in zip(evals, eq_xt)) % R
    chal, _ = sumcheck_prove_coeffs(tr, n, merged, [(1, points[i], i) for i in range(len(points))], tilde)
    g = [0] * (1 << n)
    for i, mp in enumerate(merged):
        s = eq_xy_eval(chal, points[i])
        g = [(a + s * b) % R for a, b in zip(g, mp)]
    kzg_open(srs, tr, g, chal)
