"""Oracle self-consistency: the reference's own test shapes (prove -> verify round trips) and
its `sanity-check` invariants, plus the Lasso specification's soundness negatives."""
import numpy as np
import pytest

import oracle as O


@pytest.fixture(scope="module")
def kz16():
    return O.Kzg(O.rand_fr(7, 16))


@pytest.mark.parametrize("n", [1, 2, 7, 11])
def test_sumcheck_roundtrip_like_reference(n):
    """pb/piop/sum_check.rs:140-177: prove, re-read, verify, recompute the virtual polynomial at x."""
    a, b, y = O.rand_fr(1, 1 << n), O.rand_fr(2, 1 << n), O.rand_fr(3, n)
    s = O.sum_eq_ab(y, a, b)
    one = O.fr_from_ints([1])[0]
    tr = O.Transcript()
    ch, ev = O.sumcheck_prove_evals(tr, n, [a, b], y, [(one, [0, 1])], s)
    assert len(tr.proof()) == n * 4 * 32
    fin, chv = O.sumcheck_verify(O.Transcript(tr.proof()), n, 3, s)
    assert (chv == ch).all()
    assert (O.evaluate(a, ch) == ev[0]).all() and (O.evaluate(b, ch) == ev[1]).all()
    assert (O.field_op("mul", O.field_op("mul", O.eq_xy_eval(ch, y), ev[0]), ev[1])[0] == fin).all()
    # a wrong claimed sum is caught at round 0 ("Expect sum ... but get ...", classic.rs:185-190)?  No:
    # p(0) is derived from the claim, so the transcript stays consistent and only the final check fails.
    bad = O.field_op("add", s, one)[0]
    tr2 = O.Transcript()
    ch2, ev2 = O.sumcheck_prove_evals(tr2, n, [a, b], y, [(one, [0, 1])], bad)
    fin2, _ = O.sumcheck_verify(O.Transcript(tr2.proof()), n, 3, bad)
    assert not (O.field_op("mul", O.field_op("mul", O.eq_xy_eval(ch2, y), ev2[0]), ev2[1])[0] == fin2).all()


def test_mle_fix_var_evaluate_agree():
    """pb/poly/multilinear.rs:663-712: evaluate == iterated fix_var, incl. boolean coordinates."""
    n = 9
    p = O.rand_fr(4, 1 << n)
    x = O.rand_fr(5, n)
    x[2] = O.fr_from_ints([0])[0]
    x[5] = O.fr_from_ints([1])[0]
    cur = p
    for i in range(n):
        cur = O.fix_var(cur, x[i])
    assert (cur[0] == O.evaluate(p, x)).all()
    eq = O.eq_xy(x)
    acc = O.fr_from_ints([0])
    acc = O.fr_from_ints([sum(a * b for a, b in zip(O.fr_to_ints(p), O.fr_to_ints(eq))) % O.R_MOD])
    assert (acc[0] == O.evaluate(p, x)).all()  # <p, eq(., x)> — the identity the GPU evaluate uses


@pytest.mark.parametrize("nv", [3, 8])
def test_kzg_commit_open_verify(kz16, nv):
    """pb/pcs/multilinear.rs:293-338 harness shape + sanity-check invariants of kzg.rs:286-297."""
    poly, pt = O.rand_fr(10 + nv, 1 << nv), O.rand_fr(20 + nv, nv)
    cm = kz16.commit(poly)
    tr = O.Transcript()
    ev = kz16.open(tr, poly, pt)
    assert (ev == O.evaluate(poly, pt)).all()
    assert kz16.verify(O.Transcript(tr.proof()), cm, pt, ev)
    wrong = O.field_op("add", ev, O.fr_from_ints([1]))[0]
    assert not kz16.verify(O.Transcript(tr.proof()), cm, pt, wrong)
    # commitment homomorphism (pcs/multilinear.rs:215-226)
    q = O.rand_fr(30 + nv, 1 << nv)
    assert (O.g1_add(cm, kz16.commit(q)) == kz16.commit(O.field_op("add", poly, q))).all()


def test_msm_matches_naive(kz16):
    n = 70
    sc = O.rand_fr(40, n)
    bases = kz16.eqs(7)[:n]
    acc = np.zeros(8, dtype=np.uint64)
    for s, b in zip(sc, bases):
        acc = O.g1_add(acc, O.g1_mul(b, s))
    assert (O.msm(sc, bases) == acc).all()
    assert (O.msm(sc[:0], bases[:0]) == 0).all()


def test_batch_open_verify_roundtrip(kz16):
    """pb/pcs/multilinear.rs:340-406: 8 polys, 4 points, random (poly, point) pairs."""
    nv = 6
    polys = [O.rand_fr(50 + i, 1 << nv) for i in range(8)]
    points = [O.rand_fr(60 + i, nv) for i in range(4)]
    pairs = [(i % 8, (3 * i) % 4) for i in range(11)]
    evals = [(p, q, O.evaluate(polys[p], points[q])) for p, q in pairs]
    tr = O.Transcript()
    kz16.batch_open(tr, polys, points, evals)
    comms = [kz16.commit(p) for p in polys]
    assert kz16.batch_verify(O.Transcript(tr.proof()), comms, points, evals)
    evals[3] = (evals[3][0], evals[3][1], O.field_op("add", evals[3][2], O.fr_from_ints([1]))[0])
    assert not kz16.batch_verify(O.Transcript(tr.proof()), comms, points, evals)


def test_grand_product_claims_are_leaf_evaluations():
    """template invariant (fractional_sum_check.rs:184-187): returned claims == leaf MLEs at the point."""
    T, h = 3, 5
    leaves = [O.rand_fr(70 + t, 1 << h) for t in range(T)]
    tr = O.Transcript()
    claims, point = O.grand_product_prove(tr, leaves)
    for t in range(T):
        assert (claims[t] == O.evaluate(leaves[t], point)).all()


def _ops(kind, chunks, mu, seed):
    xs, ys = O.rand_u64s(seed, 1 << mu), O.rand_u64s(seed + 1, 1 << mu)
    bits = (16 if kind == O.TABLE_RANGE else 8) * chunks
    if bits < 64:
        xs &= np.uint64((1 << bits) - 1)
        ys &= np.uint64((1 << bits) - 1)
    xs[1::2], ys[1::2] = xs[0::2], ys[0::2]
    return xs, (None if kind == O.TABLE_RANGE else ys)


@pytest.mark.parametrize("kind,chunks,mu", [(O.TABLE_RANGE, 4, 5), (O.TABLE_AND, 8, 4), (O.TABLE_XOR, 2, 6)])
def test_lasso_prove_verify_and_negatives(kz16, kind, chunks, mu):
    xs, ys = _ops(kind, chunks, mu, 80 + mu)
    tr = O.Transcript()
    assert O.lasso_prove(kz16, tr, kind, chunks, mu, xs, ys)
    proof = tr.proof()
    assert O.lasso_verify(kz16, O.Transcript(proof), kind, chunks, mu)
    for pos in (10, len(proof) // 4, len(proof) // 2, len(proof) - 5):
        bad = bytearray(proof)
        bad[pos] ^= 0x01
        assert not O.lasso_verify(kz16, O.Transcript(bytes(bad)), kind, chunks, mu), pos
    assert not O.lasso_verify(kz16, O.Transcript(proof[:-64]), kind, chunks, mu)  # truncated
    assert not O.lasso_verify(kz16, O.Transcript(proof), kind, chunks, mu + 1) if mu < 15 else True


def test_lasso_witness_semantics():
    """read_ts = # earlier accesses, final_cts = total accesses, E = T[dim], a = g(E)."""
    mu, c = 6, 4
    xs, _ = _ops(O.TABLE_RANGE, c, mu, 90)
    mt, st = O.lasso_witness(O.TABLE_RANGE, c, mu, xs)
    a = O.fr_to_ints(mt[0])
    assert a == [int(x) for x in xs]
    for t in range(c):
        dim = O.fr_to_ints(mt[1 + t])
        assert dim == [(int(x) >> (16 * t)) & 0xFFFF for x in xs]
        assert O.fr_to_ints(mt[1 + c + t]) == dim  # identity subtable
        seen, ts = {}, []
        for d in dim:
            ts.append(seen.get(d, 0))
            seen[d] = seen.get(d, 0) + 1
        assert O.fr_to_ints(mt[1 + 2 * c + t]) == ts
        cts = O.fr_to_ints(st[t])
        assert all(cts[d] == n for d, n in seen.items()) and sum(cts) == 1 << mu


# ---- the reference's own generic sum-check test shapes (pb/piop/sum_check.rs:194-300) on the oracle ---------------
from ref_shapes import reference_lagrange_case as _reference_lagrange_case, reference_rotation_case as _reference_rotation_case  # noqa: E402


def _run_zero_check(n, expr, polys, seed):
    """run_zero_check / run_sum_check (sum_check.rs:140-192): prove with sum 0, verify, then recompute the expression at
    x from fresh evaluations of the (rotated) polynomials and compare with the verifier's final claim."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from pymodel_hyperplonk import eval_tree

    N = 1 << n
    alpha, y = O.rand_fr(seed + 1, 1), O.rand_fr(seed + 2, n)
    zero = O.fr_from_ints([0])[0]
    tr = O.Transcript()
    x, evals, deg = O.sumcheck_prove_generic(tr, n, expr, polys, alpha, [y], zero)
    assert deg == expr.degree() and len(tr.proof()) == n * (deg + 1) * 32
    fin, xv = O.sumcheck_verify(O.Transcript(tr.proof()), n, deg, zero)
    assert (xv == x).all()
    for i, p in enumerate(polys):  # ProverState::into_evals: every polynomial at x
        assert (O.evaluate(p, x) == evals[i]).all()
    order = [int(b) for b in O.bh_iter(n)]

    def leaf(node):
        if node[0] == "poly":  # evaluate_for_rotation + rotation_eval == the rotated table evaluated at x
            rotated = np.ascontiguousarray(polys[node[1]][[O.bh_rotate(n, b, node[2]) for b in range(N)]])
            return O.fr_to_ints(O.evaluate(rotated, x))[0]
        if node[0] == "lagrange":
            onehot = [0] * N
            onehot[order[node[1] % N]] = 1
            return O.fr_to_ints(O.evaluate(O.fr_from_ints(onehot), x))[0]
        if node[0] == "eq":
            return O.fr_to_ints(O.eq_xy_eval(x, y))[0]
        if node[0] == "identity":
            return O.fr_to_ints(O.evaluate(O.fr_from_ints(list(range(N))), x))[0]
        raise ValueError(node)

    assert eval_tree(expr.node, leaf, O.fr_to_ints(alpha)) == O.fr_to_ints(fin)[0]
    return tr.proof()


@pytest.mark.parametrize("n", [2, 3])
def test_reference_sum_check_lagrange_shape(n):
    expr, polys = _reference_lagrange_case(n)
    _run_zero_check(n, expr, polys, 40 + n)


@pytest.mark.parametrize("n", [2, 3, 5, 8])
def test_reference_sum_check_rotation_shape(n):
    expr, polys = _reference_rotation_case(n, 50 + n)
    _run_zero_check(n, expr, polys, 60 + n)
    # a polynomial that is NOT the rotation of its predecessor breaks the zero check: the final claim no longer matches
    bad = list(polys)
    bad[1] = O.rand_fr(99, 1 << n)
    with pytest.raises(AssertionError):
        _run_zero_check(n, expr, bad, 60 + n)


# ---- verified specification of a planned kernel optimisation: eq-factored round polynomial ------
def _eq_factored_prove(M, tr, n, polys, y, terms, claim, switch_round):
    """The same messages as sumcheck_prove_evals, computed WITHOUT binding an eq table in the big rounds:
        p_j(X) = c_j * l_j(X) * Q_j(X),  c_j = Π_{i<j} eq1(y_i, r_i),  l_j(X) = eq1(y_j, X),
        Q_j(X) = Σ_b E_j[b] * G(X, b),   E_j = eq table of (y_{j+1}, ..., y_{n-1})  (a level of the eq_xy doubling).
    Q_j has degree d - 1: it is evaluated at X = 1 .. d - 1 and its leading coefficient is accumulated as a third sum,
    Q_j(d) follows by extrapolation; p(0) stays DERIVED from the claim (eval.rs:129) so that an inconsistent claim
    gives the reference's bytes too. From `switch_round` on, the bound eq table c_{S-1} * E_{S-2} is materialised and
    the rounds run exactly as in the reference (small rounds / single-launch tail kernel)."""
    R = M.R
    d = 1 + max(len(t[1]) for t in terms)
    NP = d - 1
    assert all(len(t[1]) == NP for t in terms) and NP in (1, 2)
    tabs = [list(p) for p in polys]
    chal, c, c_prev, eq = [], 1, 1, None

    def eq1(a, b):
        return (a * b + (1 - a) * (1 - b)) % R

    for j in range(n):
        pairs = len(tabs[0]) // 2
        if j < switch_round:
            E = M.eq_xy(y[j + 1:]) if j + 1 < n else [1]
            q1 = q2 = lead = 0
            for b in range(pairs):
                for coeff, idx in terms:
                    u0, u1 = tabs[idx[0]][2 * b], tabs[idx[0]][2 * b + 1]
                    if NP == 2:
                        v0, v1 = tabs[idx[1]][2 * b], tabs[idx[1]][2 * b + 1]
                        du, dv = u1 - u0, v1 - v0
                        eu1, edu = E[b] * u1 % R, E[b] * du % R           # two reduced products
                        q1 += coeff * (eu1 * v1)                          # three unreduced products, accumulated wide
                        q2 += coeff * ((eu1 + edu) * (v1 + dv))
                        lead += coeff * (edu * dv)
                    else:
                        q1 += coeff * (E[b] * u1)
                        lead += coeff * (E[b] * (u1 - u0))                # the slope of the linear Q
            q1, q2, lead = q1 % R, q2 % R, lead % R
            s0, s1 = c * (1 - y[j]) % R, c * y[j] % R                     # c_j * l_j(0), c_j * l_j(1)
            s = [(s0 + k * (s1 - s0)) % R for k in range(d + 1)]
            if NP == 2:
                q = [None, q1, q2, (2 * q2 - q1 + 2 * lead) % R]          # Q(3) = 2 Q(2) - Q(1) + 2 * lead
            else:
                q = [None, q1, (q1 + lead) % R]
            ev = [None] + [s[k] * q[k] % R for k in range(1, d + 1)]
            ev[0] = (claim - ev[1]) % R
        else:
            if eq is None:  # the switch: bound eq table of the previous round's size, then the reference's bind
                eq = [c_prev * e % R for e in M.eq_xy(y[j - 1:])] if j else M.eq_xy(y)
                if j:
                    eq = M.fix_var(eq, chal[-1])
            ev = [0] * (d + 1)
            for b in range(pairs):
                for x in range(1, d + 1):
                    at = lambda t: (t[2 * b] + x * (t[2 * b + 1] - t[2 * b])) % R
                    sm = 0
                    for coeff, idx in terms:
                        p = coeff
                        for i in idx:
                            p = p * at(tabs[i]) % R
                        sm += p
                    ev[x] = (ev[x] + at(eq) * sm) % R
            ev[0] = (claim - ev[1]) % R
        for e in ev:
            tr.write_fe(e)
        r = tr.squeeze()
        chal.append(r)
        claim = M.interpolate(ev, r)
        c_prev, c = c, c * eq1(y[j], r) % R
        if eq is not None:
            eq = M.fix_var(eq, r)
        tabs = [M.fix_var(t, r) for t in tabs]
    return chal, [t[0] for t in tabs]


@pytest.mark.parametrize("n,T,NP,switch_round,true_claim", [(5, 1, 2, 5, True), (5, 3, 2, 3, False), (4, 4, 1, 2, False),
                                                           (4, 2, 2, 1, True), (3, 2, 1, 0, False), (6, 2, 2, 4, False)])
def test_eq_factored_round_polynomial_reproduces_the_reference_messages(n, T, NP, switch_round, true_claim):
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import pymodel as M

    seed = 700 + 10 * n + T
    polys = [M.rand_fr(seed + i, 1 << n) for i in range(T * NP)]
    w, y = M.rand_fr(seed + 50, T), M.rand_fr(seed + 51, n)
    terms = [(w[t], list(range(t * NP, (t + 1) * NP))) for t in range(T)]
    claim = M.rand_fr(seed + 52, 1)[0]
    if true_claim:
        eq = M.eq_xy(y)
        claim = 0
        for b in range(1 << n):
            for coeff, idx in terms:
                p = coeff * eq[b]
                for i in idx:
                    p = p * polys[i][b] % M.R
                claim = (claim + p) % M.R
    t_ref, t_new = M.Transcript(), M.Transcript()
    ch_ref, ev_ref = M.sumcheck_prove_evals(t_ref, n, polys, y, terms, claim)
    ch_new, ev_new = _eq_factored_prove(M, t_new, n, polys, y, terms, claim, switch_round)
    assert t_new.stream == t_ref.stream and ch_new == ch_ref and ev_new == ev_ref


def test_e_commitment_is_a_regrouping_of_the_dim_buckets(kz16):
    """The identity behind the grouped MSM jobs of the CUDA prover (msm.cu msm_group_kernel): E_t = T[dim_t] is constant
    per address and shares its bases with dim_t, so with B_d = Σ_{j: dim_t[j] = d} G_j (the bucket sums of the dim_t
    commitment) Com(dim_t) = Σ_d d B_d and Com(E_t) = Σ_v v Σ_{d: T[d] = v} B_d — no second pass over the points.
    Checked with the oracle's group arithmetic for the AND and XOR subtables (address 0 has T = 0 in both)."""
    mu = 7
    bases = kz16.eqs(mu)
    for kind in (O.TABLE_AND, O.TABLE_XOR):
        xs, ys = O.rand_u64s(60 + kind, 1 << mu) & np.uint64(0xFFFF), O.rand_u64s(70 + kind, 1 << mu) & np.uint64(0xFFFF)
        xs[1::3] = xs[0::3][: len(xs[1::3])]
        ys[1::3] = ys[0::3][: len(ys[1::3])]
        mt, _ = O.lasso_witness(kind, 2, mu, xs, ys)
        dim, e = O.fr_to_ints(mt[1]), O.fr_to_ints(mt[3])  # a | dim_0 dim_1 | e_0 e_1 | ts_0 ts_1
        table = {d: ((d >> 8) & (d & 0xFF)) if kind == O.TABLE_AND else ((d >> 8) ^ (d & 0xFF)) for d in set(dim)}
        assert all(table[d] == v for d, v in zip(dim, e))
        buckets = {}
        for j, d in enumerate(dim):
            buckets[d] = bases[j] if d not in buckets else O.g1_add(buckets[d], bases[j])
        by_value = {}
        for d, b in buckets.items():
            v = table[d]
            if v:
                by_value[v] = b if v not in by_value else O.g1_add(by_value[v], b)
        acc_dim = acc_e = None
        for d, b in buckets.items():
            if d:
                t = O.g1_mul(b, O.fr_from_ints([d])[0])
                acc_dim = t if acc_dim is None else O.g1_add(acc_dim, t)
        for v, b in by_value.items():
            t = O.g1_mul(b, O.fr_from_ints([v])[0])
            acc_e = t if acc_e is None else O.g1_add(acc_e, t)
        assert (acc_dim == O.msm(mt[1], bases)).all()
        assert (acc_e == O.msm(mt[3], bases)).all()
