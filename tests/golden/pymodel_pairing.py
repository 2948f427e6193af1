"""Independent pure-Python model of the BN254 optimal-ate pairing, in a DIFFERENT formulation from oracle/pairing.hpp:
Fq12 is the single extension Fq[w]/(w^12 - 18 w^6 + 82) (not a 2-3-2 tower), G2 points are untwisted into E(Fq12)
and every step — point doubling / addition, the line functions, their divisions — is generic Fq12 arithmetic (the
structure of the widely used reference implementation in py_ecc, rewritten here from the definition).

    e(P, Q) = ( f_{6u+2, Q}(P) * l_{T, pi(Q)}(P) * l_{T + pi(Q), -pi^2(Q)}(P) ) ^ ((p^12 - 1) / r)

`tower_to_w` converts an element given in the oracle's tower basis ((x + y u) v^j w^i, u = w^6 - 9, v = w^2) to the
12 coefficients used here, so that the two implementations can be compared coefficient by coefficient."""
P = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
ATE_LOOP = 29793968203157093288  # 6u + 2, u = 4965661367192848881


def f12(c):
    return [x % P for x in c] + [0] * (12 - len(c))


ONE = f12([1])
ZERO = f12([])


def add(a, b):
    return [(x + y) % P for x, y in zip(a, b)]


def sub(a, b):
    return [(x - y) % P for x, y in zip(a, b)]


def mul(a, b):
    t = [0] * 23
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                t[i + j] += x * y
    for k in range(22, 11, -1):  # w^12 = 18 w^6 - 82
        c = t[k]
        if c:
            t[k - 6] += 18 * c
            t[k - 12] -= 82 * c
    return [x % P for x in t[:12]]


def scal(a, k):
    return [x * k % P for x in a]


def poly_deg(a):
    d = len(a) - 1
    while d and a[d] == 0:
        d -= 1
    return d


def inv(a):
    """extended Euclid in Fq[w] against the modulus polynomial"""
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], [82, 0, 0, 0, 0, 0, -18 % P, 0, 0, 0, 0, 0, 1]
    while poly_deg(low):
        # r = high / low (polynomial division, rounded)
        dl, dh = poly_deg(low), poly_deg(high)
        r = [0] * 13
        tmp = list(high)
        ilow = pow(low[dl], -1, P)
        for i in range(dh - dl, -1, -1):
            r[i] = tmp[dl + i] * ilow % P
            for c in range(dl + 1):
                tmp[c + i] = (tmp[c + i] - low[c] * r[i]) % P
        nm, new = list(hm), list(high)
        for i in range(13):
            for j in range(13 - i):
                nm[i + j] = (nm[i + j] - lm[i] * r[j]) % P
                new[i + j] = (new[i + j] - low[i] * r[j]) % P
        lm, low, hm, high = nm, new, lm, low
    il = pow(low[0], -1, P)
    return [x * il % P for x in lm[:12]]


def div(a, b):
    return mul(a, inv(b))


def fpow(a, e):
    res = ONE
    while e:
        if e & 1:
            res = mul(res, a)
        a = mul(a, a)
        e >>= 1
    return res


W = f12([0, 1])
W2, W3 = mul(W, W), mul(mul(W, W), W)


def twist(q):
    """E'(Fq2) -> E(Fq12): (x, y) with x = x0 + x1 u, u = w^6 - 9, then (x w^2, y w^3)"""
    (x0, x1), (y0, y1) = q
    nx = f12([x0 - 9 * x1, 0, 0, 0, 0, 0, x1])
    ny = f12([y0 - 9 * y1, 0, 0, 0, 0, 0, y1])
    return (mul(nx, W2), mul(ny, W3))


def cast_g1(p):
    return (f12([p[0]]), f12([p[1]]))


def double(pt):
    x, y = pt
    m = div(scal(mul(x, x), 3), scal(y, 2))
    nx = sub(mul(m, m), scal(x, 2))
    return (nx, sub(mul(m, sub(x, nx)), y))


def padd(p1, p2):
    if p1 is None or p2 is None:
        return p1 if p2 is None else p2
    (x1, y1), (x2, y2) = p1, p2
    if x1 == x2:
        return double(p1) if y1 == y2 else None
    m = div(sub(y2, y1), sub(x2, x1))
    nx = sub(sub(mul(m, m), x1), x2)
    return (nx, sub(mul(m, sub(x1, nx)), y1))


def linefunc(p1, p2, t):
    (x1, y1), (x2, y2), (xt, yt) = p1, p2, t
    if x1 != x2:
        m = div(sub(y2, y1), sub(x2, x1))
        return sub(mul(m, sub(xt, x1)), sub(yt, y1))
    if y1 == y2:
        m = div(scal(mul(x1, x1), 3), scal(y1, 2))
        return sub(mul(m, sub(xt, x1)), sub(yt, y1))
    return sub(xt, x1)


def miller_loop(q, p):
    r, f = q, ONE
    for i in range(ATE_LOOP.bit_length() - 2, -1, -1):
        f = mul(mul(f, f), linefunc(r, r, p))
        r = double(r)
        if (ATE_LOOP >> i) & 1:
            f = mul(f, linefunc(r, q, p))
            r = padd(r, q)
    q1 = (fpow(q[0], P), fpow(q[1], P))
    nq2 = (fpow(q1[0], P), sub(ZERO, fpow(q1[1], P)))
    f = mul(f, linefunc(r, q1, p))
    r = padd(r, q1)
    f = mul(f, linefunc(r, nq2, p))
    return fpow(f, (P ** 12 - 1) // R)


def pairing(g1_affine, g2_affine):
    """g1_affine = (x, y) ints; g2_affine = ((x0, x1), (y0, y1)) on the twist"""
    return miller_loop(twist(g2_affine), cast_g1(g1_affine))


def tower_to_w(parts):
    """12 ints in the oracle's order [c0.a0.c0, c0.a0.c1, c0.a1.c0, c0.a1.c1, c0.a2.*, c1.a0.*, c1.a1.*, c1.a2.*]
    -> w-basis coefficients"""
    out = [0] * 12
    for t in range(6):
        i, j = t // 3, t % 3
        x, y = parts[2 * t], parts[2 * t + 1]
        out[2 * j + i] = (out[2 * j + i] + x - 9 * y) % P
        out[2 * j + i + 6] = (out[2 * j + i + 6] + y) % P
    return out


G1 = (1, 2)
G2 = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
       11559732032986387107991004021392285783925812861821192530917403151452391805634),
      (8495653923123431417604973247489272438418190587263600148770280649306958101930,
       4082367875863433681332203403145435568316851327593401208105741076214120093531))
