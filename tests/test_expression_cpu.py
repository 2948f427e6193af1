"""CPU tests of the expression layer: the reference's structural KAT for `compose` (preprocessor.rs:216-256),
BooleanHypercube vs the oracle, the bytecode compiler vs direct tree evaluation, and two independent oracle
paths (generic ProverState restatement vs fixed-shape prover; oracle vs a pure-Python model that evaluates the
expression over MATERIALISED leaf tables — the strategy the GPU uses)."""
import random

import pytest

import numpy as np

import oracle as O
from halo2_lasso_b200.expression import (OP_ADD, OP_MUL, OP_NEG, OP_SUB, BooleanHypercube, Expression, R_MOD,
                                         compile_expression, vanilla_plonk_expression)

E = Expression


def hand_written_vanilla_plonk(num_vars):
    pi, q_l, q_r, q_m, q_o, q_c, w_l, w_r, w_o, s_1, s_2, s_3 = (E.polynomial(i) for i in range(12))
    z, z_next = E.polynomial(12), E.polynomial(12, 1)
    beta, gamma, alpha = (E.challenge(i) for i in range(3))
    id_1, id_2, id_3 = (E.constant(i << num_vars) + E.identity() for i in range(3))
    l_1, one = E.lagrange(1), E.one()
    constraints = [
        q_l * w_l + q_r * w_r + q_m * w_l * w_r + q_o * w_o + q_c + pi,
        l_1 * (z - one),
        (z * ((w_l + beta * id_1 + gamma) * (w_r + beta * id_2 + gamma) * (w_o + beta * id_3 + gamma)))
        - (z_next * ((w_l + beta * s_1 + gamma) * (w_r + beta * s_2 + gamma) * (w_o + beta * s_3 + gamma))),
    ]
    return E.distribute_powers(constraints, alpha) * E.eq_xy(0)


def test_compose_vanilla_plonk_kat():
    assert vanilla_plonk_expression(3) == hand_written_vanilla_plonk(3)
    assert vanilla_plonk_expression(3).degree() == 5


def test_boolean_hypercube_matches_oracle():
    for n in (1, 2, 5, 11):
        bh = BooleanHypercube(n)
        assert bh.iter() == [int(x) for x in O.bh_iter(n)]
        assert sorted(bh.iter()) == list(range(1 << n))  # LFSR order visits every row once
        for b in (0, 1, (1 << n) - 1, (1 << n) // 3):
            for rot in (-2, -1, 0, 1, 2):
                if abs(rot) <= n:
                    assert bh.rotate(b, rot) == O.bh_rotate(n, b, rot)
        assert all(bh.prev(bh.next(b)) == b for b in range(1, 1 << n))
        order = bh.iter()
        for i in (0, 1, 2, 5, -1, -2, (1 << n) - 1, 1 << n):
            assert bh.nth(i) == order[i % (1 << n)]


def eval_tree(n, leaf_val, ch):
    k = n[0]
    if k == "const":
        return n[1]
    if k == "chal":
        return ch[n[1]]
    if k in ("identity", "lagrange", "eq", "poly"):
        return leaf_val[n]
    if k == "neg":
        return -eval_tree(n[1], leaf_val, ch) % R_MOD
    if k == "sum":
        return (eval_tree(n[1], leaf_val, ch) + eval_tree(n[2], leaf_val, ch)) % R_MOD
    if k == "prod":
        return eval_tree(n[1], leaf_val, ch) * eval_tree(n[2], leaf_val, ch) % R_MOD
    if k == "scaled":
        return eval_tree(n[1], leaf_val, ch) * n[2] % R_MOD
    base = eval_tree(n[2], leaf_val, ch)
    acc, pw = eval_tree(n[1][0], leaf_val, ch), base
    for c in n[1][1:]:
        acc, pw = (acc + pw * eval_tree(c, leaf_val, ch)) % R_MOD, pw * base % R_MOD
    return acc


def run_program(leaves, consts, prog, leaf_val):
    slots = {i: leaf_val[l] for i, l in enumerate(leaves)}
    slots.update({len(leaves) + i: c for i, c in enumerate(consts)})
    for op, d, a, b in prog:
        x, y = slots[a], slots[b]
        slots[d] = {OP_ADD: (x + y), OP_SUB: (x - y), OP_MUL: (x * y), OP_NEG: -x}[op] % R_MOD
    return slots[prog[-1][1]]


def test_compiled_program_equals_tree_evaluation():
    rng = random.Random(1)
    expr = vanilla_plonk_expression(4)
    ch = [rng.randrange(R_MOD) for _ in range(3)]
    leaves, consts, prog = compile_expression(expr, ch)
    assert len(leaves) == 17  # 13 polys + z_next + identity + lagrange(1) + eq  (z cur and next are separate tables)
    for _ in range(20):
        lv = {l: rng.randrange(R_MOD) for l in leaves}
        assert run_program(leaves, consts, prog, lv) == eval_tree(expr.node, lv, ch)
    # CSE: w_l + beta*... sub-terms etc. are shared; the program is much shorter than the tree
    assert len(prog) < 60


def test_native_compiler_program_equals_tree_evaluation():
    """The library's own compiler (csrc/expr.hpp through b200_expression_compile, host code: runs without a GPU) on
    the vanilla-plonk and plonk-with-lookup zero-check expressions and on a small expression with every node kind:
    interpreting its program over random leaf values gives the value of the expression tree."""
    import halo2_lasso_b200 as hl
    from halo2_lasso_b200 import hyperplonk as H
    from halo2_lasso_b200.expression import compose, serialize_expression

    R = 1 << 256
    rinv = pow(R, -1, R_MOD)
    kinds = {1: "identity", 2: "lagrange", 3: "eq", 4: "poly"}

    def to_limbs(v):
        return [(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)]

    info, _, _ = H.rand_vanilla_plonk_with_lookup_circuit(4, 1)
    _, lookup_expr = compose(4, info.constraints, info.num_poly, info.permutation_polys, lookups=info.lookups)
    misc = (E.polynomial(0) * E.polynomial(1, 1) - E.identity() * E.lagrange(-1) + E.challenge(1) * 7
            + E.distribute_powers([E.polynomial(2), -E.polynomial(0), E.constant(5)], E.challenge(0)) + E.eq_xy(1))
    # two witness phases: circuit challenges 0, 1 in front of beta, gamma, alpha; an instance column at Rotation::next
    info2, _, _ = H.rand_two_phase_circuit(4, 1)
    _, phased_expr = compose(4, info2.constraints, info2.num_poly, info2.permutation_polys, num_challenges=2, lookups=info2.lookups)
    rng = random.Random(2)
    for expr, nleaves in ((vanilla_plonk_expression(4), 17), (lookup_expr, 23), (misc, 6), (E.polynomial(3), 1), (phased_expr, 22)):
        _, consts = serialize_expression(expr, [], [])
        cm = np.asarray([to_limbs(c * R % R_MOD) for c in consts] or [[0, 0, 0, 0]], dtype=np.uint64)
        leaves, cvals, cchal, ops, ntemps, degree = hl.compile_expression_native(expr, cm)
        assert degree == expr.degree() and len(leaves) == nleaves
        py_leaves = [(kinds[k],) if k == 1 else ((kinds[k], a) if k in (2, 3) else (kinds[k], a, b)) for k, a, b in leaves]
        assert py_leaves == expr.leaves()
        ch = [rng.randrange(R_MOD) for _ in range(5)]
        cints = [ch[j] if j >= 0 else sum(int(x) << (64 * i) for i, x in enumerate(v)) * rinv % R_MOD
                 for v, j in zip(cvals, cchal)]
        K, C_ = len(leaves), len(cints)
        assert max(d for _, d, _, _ in ops) - (K + C_) + 1 == ntemps and ntemps <= 8
        for _ in range(10):
            lv = {l: rng.randrange(R_MOD) for l in py_leaves}
            assert run_program(py_leaves, cints, ops, lv) == eval_tree(expr.node, lv, ch)


def test_generic_oracle_agrees_with_fixed_shape_oracle():
    n = 6
    a, b, y = O.rand_fr(1, 1 << n), O.rand_fr(2, 1 << n), O.rand_fr(3, n)
    s = O.sum_eq_ab(y, a, b)
    one = O.fr_from_ints([1])[0]
    t1, t2 = O.Transcript(), O.Transcript()
    ch1, ev1 = O.sumcheck_prove_evals(t1, n, [a, b], y, [(one, [0, 1])], s)
    expr = E.eq_xy(0) * E.polynomial(0) * E.polynomial(1)
    ch2, ev2, deg = O.sumcheck_prove_generic(t2, n, expr, [a, b], np.zeros((0, 4), dtype=np.uint64), [y], s)
    assert deg == 3 and t1.proof() == t2.proof() and (ch1 == ch2).all() and (ev1 == ev2).all()


def test_generic_oracle_vs_materialised_table_model():
    """Independent check that evaluating over materialised identity / one-hot Lagrange / rotated tables gives the
    reference's ProverState semantics (identity offsets, Lagrange (b, value) halving, bh.rotate in round 0)."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import pymodel as M

    n = 4
    expr = vanilla_plonk_expression(n)
    polys_i = [M.rand_fr(500 + i, 1 << n) for i in range(13)]
    ch_i = M.rand_fr(600, 3)
    y_i = M.rand_fr(601, n)
    claim_i = M.rand_fr(602, 1)[0]
    bh = BooleanHypercube(n)
    order = bh.iter()
    leaves, consts, prog = compile_expression(expr, ch_i)
    tabs = []
    for l in leaves:
        if l[0] == "poly":
            tabs.append([polys_i[l[1]][bh.rotate(b, l[2])] for b in range(1 << n)])
        elif l[0] == "eq":
            tabs.append(M.eq_xy(y_i))
        elif l[0] == "identity":
            tabs.append(list(range(1 << n)))
        else:
            tabs.append([1 if b == order[l[1] % (1 << n)] else 0 for b in range(1 << n)])
    # pure-Python sum-check over the tables with the compiled program
    d = expr.degree()
    tr = M.Transcript()
    claim, cur = claim_i, [list(t) for t in tabs]
    for _ in range(n):
        ev = [0] * (d + 1)
        for b in range(len(cur[0]) // 2):
            for x in range(1, d + 1):
                lv = {l: (t[2 * b] + x * (t[2 * b + 1] - t[2 * b])) % R_MOD for l, t in zip(leaves, cur)}
                ev[x] = (ev[x] + run_program(leaves, consts, prog, lv)) % R_MOD
        ev[0] = (claim - ev[1]) % R_MOD
        for e in ev:
            tr.write_fe(e)
        r = tr.squeeze()
        claim = M.interpolate(ev, r)
        cur = [M.fix_var(t, r) for t in cur]
    to = O.Transcript()
    O.sumcheck_prove_generic(to, n, expr, [O.fr_from_ints(p) for p in polys_i], O.fr_from_ints(ch_i), [O.fr_from_ints(y_i)],
                             O.fr_from_ints([claim_i])[0])
    assert to.proof() == tr.stream


def test_native_compiler_rejects_malformed_token_streams():
    """b200_expression_compile (host code, no GPU): truncated / out-of-range streams are B200_ERR_ARG, not a crash."""
    import ctypes as C

    import halo2_lasso_b200 as hl

    def compile_raw(tokens, nconsts=1):
        t = np.asarray(tokens, dtype=np.int32)
        consts = np.zeros((max(1, nconsts), 4), dtype=np.uint64)
        leaves = np.zeros((64, 3), dtype=np.int32)
        cout = np.zeros((64, 4), dtype=np.uint64)
        cchal = np.zeros(64, dtype=np.int32)
        ops = np.zeros((64, 4), dtype=np.int32)
        n = [C.c_int() for _ in range(5)]
        return hl.lib().b200_expression_compile(hl._p(t), C.c_int(len(t)), hl._p(consts), C.c_int(nconsts), hl._p(leaves),
                                                C.c_int(64), C.byref(n[0]), hl._p(cout), hl._p(cchal), C.c_int(64),
                                                C.byref(n[1]), hl._p(ops), C.c_int(64), C.byref(n[2]), C.byref(n[3]),
                                                C.byref(n[4]))

    assert compile_raw([8, 4, 0, 0, 4, 1, 0]) == hl.B200_OK          # poly0 * poly1
    assert compile_raw([8, 4, 0, 0]) == hl.B200_ERR_ARG               # product with one operand
    assert compile_raw([8, 4, 0, 0, 4, 1]) == hl.B200_ERR_ARG         # truncated query
    assert compile_raw([0, 3]) == hl.B200_ERR_ARG                     # constant index out of range
    assert compile_raw([42]) == hl.B200_ERR_ARG                       # unknown node kind
    assert compile_raw([4, 0, 0, 4, 1, 0]) == hl.B200_ERR_ARG         # trailing tokens
    assert compile_raw([10, 0, 4, 0, 0]) == hl.B200_ERR_ARG           # DistributePowers with no terms


def test_native_compose_equals_the_python_compose_token_for_token():
    """`b200_expression_compose` (csrc/expr.hpp e_compose + e_serialize, host code: runs without a GPU) against
    expression.py::compose + serialize_expression — two implementations of preprocessor.rs:25-60 — on vanilla plonk
    (one and two permutation chunks), plonk with the LogUp lookup and the two-phase circuit."""
    import halo2_lasso_b200 as hl
    from halo2_lasso_b200 import hyperplonk as H
    from halo2_lasso_b200.expression import compose, serialize_expression

    rinv = pow(1 << 256, -1, R_MOD)
    cases = [(H.rand_vanilla_plonk_circuit(4, 1)[0], 0, 4), (H.rand_vanilla_plonk_circuit(5, 2)[0], 0, 3),
             (H.rand_vanilla_plonk_with_lookup_circuit(4, 1)[0], 0, 4), (H.rand_two_phase_circuit(4, 1)[0], 2, 4),
             (H.rand_two_phase_circuit(4, 1, with_lookup=False)[0], 2, 4)]
    for info, nc, md in cases:
        nz, expr = compose(info.k, info.constraints, info.num_poly, info.permutation_polys, num_challenges=nc, max_degree=md,
                           lookups=info.lookups)
        tokens, consts = serialize_expression(expr, [], [])
        nz2, tok2, cm2 = hl.compose_native(info.k, info.constraints, info.num_poly, info.permutation_polys, nc, md, info.lookups)
        assert nz2 == nz and list(tok2) == tokens
        assert [sum(int(c[j]) << (64 * j) for j in range(4)) * rinv % R_MOD for c in cm2] == consts
    # malformed circuits are argument errors: a polynomial index beyond num_poly, a challenge the circuit does not have
    info = cases[0][0]
    for bad in (info.constraints + [E.polynomial(9)], info.constraints + [E.challenge(0) * E.polynomial(1)]):
        with pytest.raises(hl.B200Error) as e:
            hl.compose_native(info.k, bad, info.num_poly, info.permutation_polys, 0, 4, info.lookups)
        assert e.value.code == hl.B200_ERR_ARG


def test_token_parsers_survive_random_streams():
    """Untrusted token streams through the host-only entry points (b200_expression_compile, b200_expression_compose,
    b200v_hyperplonk_new): random and mutated streams either parse or are argument errors — never a crash."""
    import ctypes as C

    import halo2_lasso_b200 as hl
    from halo2_lasso_b200 import verifier as V
    from halo2_lasso_b200.expression import serialize_expression

    rng = random.Random(11)
    good, consts = serialize_expression(vanilla_plonk_expression(4), [], [])
    cm = np.zeros((max(1, len(consts)), 4), dtype=np.uint64)
    cap = 4096
    leaves, cout, cchal, ops = (np.zeros((cap, 3), dtype=np.int32), np.zeros((cap, 4), dtype=np.uint64),
                                np.zeros(cap, dtype=np.int32), np.zeros((cap, 4), dtype=np.int32))
    ints = [C.c_int() for _ in range(5)]
    vk = V.MultilinearKzgVerifier.setup(O.rand_fr(7, 4))
    one, two = np.asarray([1], dtype=np.int32), np.asarray([3], dtype=np.int32)
    outcomes = set()
    for trial in range(1500):
        if trial % 3 == 0:
            toks = [rng.randrange(-2, 14) for _ in range(rng.randrange(1, 40))]
        else:
            toks = list(good)
            for _ in range(rng.randrange(1, 4)):
                toks[rng.randrange(len(toks))] = rng.randrange(-3, 40)
            if rng.random() < 0.3:
                toks = toks[: rng.randrange(1, len(toks))]
        t = np.asarray(toks, dtype=np.int32)
        rc = hl.lib().b200_expression_compile(hl._p(t), C.c_int(len(t)), hl._p(cm), C.c_int(len(consts)), hl._p(leaves), C.c_int(cap),
                                              C.byref(ints[0]), hl._p(cout), hl._p(cchal), C.c_int(cap), C.byref(ints[1]), hl._p(ops),
                                              C.c_int(cap), C.byref(ints[2]), C.byref(ints[3]), C.byref(ints[4]))
        outcomes.add(rc)
        assert rc in (hl.B200_OK, hl.B200_ERR_ARG, hl.B200_ERR_NOMEM)
        tout, nt, nc, nz = np.zeros(1 << 14, dtype=np.int32), C.c_int(), C.c_int(), C.c_int()
        rc = hl.lib().b200_expression_compose(C.c_int(4), C.c_int(9), C.c_int(0), C.c_int(1), hl._p(t), C.c_int(len(t)), C.c_int(0),
                                              None, C.c_int(0), hl._p(cm), C.c_int(len(consts)), C.c_int(3),
                                              hl._p(np.asarray([6, 7, 8], dtype=np.int32)), C.c_int(4), hl._p(tout), C.c_int(len(tout)),
                                              C.byref(nt), hl._p(cout), C.c_int(cap), C.byref(nc), C.byref(nz))
        assert rc in (hl.B200_OK, hl.B200_ERR_ARG, hl.B200_ERR_NOMEM)
        h = C.c_void_p()
        rc = V.lib().b200v_hyperplonk_new(vk.h, C.c_int(4), C.c_int(1), hl._p(one), C.c_int(1), hl._p(two), hl._p(np.zeros(1, dtype=np.int32)),
                                          C.c_int(0), C.c_int(1), hl._p(t), C.c_int(len(t)), hl._p(cm), C.c_int(len(consts)), None,
                                          C.c_int(0), None, C.c_int(0), C.byref(h))
        assert rc in (V.ACCEPT, V.ERR_ARG)
        if rc == V.ACCEPT:
            V.lib().b200v_hyperplonk_free(h)
    assert hl.B200_ERR_ARG in outcomes and hl.B200_OK in outcomes
    # a pathologically deep expression is an argument error (bounded recursion), not a stack overflow — in the prover
    # library's parser as well as in the verifier's
    deep = np.asarray([6] * 200000 + [4, 0, 0], dtype=np.int32)
    rc = hl.lib().b200_expression_compile(hl._p(deep), C.c_int(len(deep)), hl._p(cm), C.c_int(len(consts)), hl._p(leaves), C.c_int(cap),
                                          C.byref(ints[0]), hl._p(cout), hl._p(cchal), C.c_int(cap), C.byref(ints[1]), hl._p(ops),
                                          C.c_int(cap), C.byref(ints[2]), C.byref(ints[3]), C.byref(ints[4]))
    assert rc == hl.B200_ERR_ARG
    deep = np.asarray([6] * 200000 + [4, 0, 0], dtype=np.int32)
    h = C.c_void_p()
    rc = V.lib().b200v_hyperplonk_new(vk.h, C.c_int(4), C.c_int(1), hl._p(one), C.c_int(1), hl._p(two), hl._p(np.zeros(1, dtype=np.int32)),
                                      C.c_int(0), C.c_int(1), hl._p(deep), C.c_int(len(deep)), hl._p(cm), C.c_int(len(consts)), None,
                                      C.c_int(0), None, C.c_int(0), C.byref(h))
    assert rc == V.ERR_ARG
