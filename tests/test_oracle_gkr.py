"""GKR for fractional sum-checks (pb/piop/gkr/fractional_sum_check.rs): the C++ oracle against the fixtures of the
independent pure-Python model (tests/golden/pymodel_gkr.py -> gkr_golden.json), and the reference's own test
(`fractional_sum_check`, :330-370: batch of 3, prove, verify, claims == evaluate(x)) restated on the oracle."""
import json
import os

import numpy as np
import pytest

import oracle as O

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gkr_golden.json")))


def inputs(case):
    B, n, seed = case["batch"], case["num_vars"], case["seed"]
    return ([O.rand_fr(seed + b, 1 << n) for b in range(B)], [O.rand_fr(seed + 50 + b, 1 << n) for b in range(B)])


@pytest.mark.parametrize("case", G["cases"], ids=lambda c: f"b{c['batch']}-n{c['num_vars']}")
def test_oracle_reproduces_the_python_model(case):
    ps, qs = inputs(case)
    B = case["batch"]
    cl = [0] * B if case["claimed"] else None
    tr = O.Transcript()
    p_xs, q_xs, x, p0, q0 = O.fractional_sum_check_prove(tr, ps, qs, cl, cl)
    assert tr.proof().hex() == case["proof"]
    for got, key in ((p_xs, "p_xs"), (q_xs, "q_xs"), (x, "x"), (p0, "p_0s"), (q0, "q_0s")):
        assert O.fr_to_ints(got) == [int(v) for v in case[key]]


@pytest.mark.parametrize("num_vars", [1, 2, 3, 5, 8, 11])
def test_reference_test_restated(num_vars):
    B = 3
    ps = [O.rand_fr(3000 + num_vars + b, 1 << num_vars) for b in range(B)]
    qs = [O.rand_fr(3100 + num_vars + b, 1 << num_vars) for b in range(B)]
    tr = O.Transcript()
    O.fractional_sum_check_prove(tr, ps, qs)
    proof = tr.proof()
    res = O.fractional_sum_check_verify(O.Transcript(proof), num_vars, [None] * B, [None] * B)
    assert res is not None
    p_xs, q_xs, x, p0, q0 = res
    for b in range(B):
        assert (O.evaluate(ps[b], x) == p_xs[b]).all() and (O.evaluate(qs[b], x) == q_xs[b]).all()
    # the statement: Σ_i p_i / q_i == p_0 / q_0
    for b in range(B):
        p_i, q_i = O.fr_to_ints(ps[b]), O.fr_to_ints(qs[b])
        s = sum(p * pow(q, -1, O.R_MOD) for p, q in zip(p_i, q_i)) % O.R_MOD
        assert s * O.fr_to_ints(q0[b])[0] % O.R_MOD == O.fr_to_ints(p0[b])[0]
    # a flipped proof byte is rejected (or changes the claims so that they no longer match the inputs)
    bad = bytearray(proof)
    bad[len(bad) // 2] ^= 1
    res = O.fractional_sum_check_verify(O.Transcript(bytes(bad)), num_vars, [None] * B, [None] * B)
    assert res is None or any((O.evaluate(ps[b], res[2]) != res[0][b]).any() for b in range(B))


def test_claimed_values_are_absorbed_not_written():
    B, n = 2, 5
    ps = [O.rand_fr(3300 + b, 1 << n) for b in range(B)]
    qs = [O.rand_fr(3400 + b, 1 << n) for b in range(B)]
    t_none, t_some = O.Transcript(), O.Transcript()
    *_, p0, q0 = O.fractional_sum_check_prove(t_none, ps, qs)
    O.fractional_sum_check_prove(t_some, ps, qs, [0] * B, [0] * B)
    assert len(t_none.proof()) == len(t_some.proof()) + 2 * B * 32
    assert O.fractional_sum_check_verify(O.Transcript(t_some.proof()), n, list(p0), list(q0)) is not None
    wrong = [np.array(p0[1]), np.array(p0[0])]
    assert O.fractional_sum_check_verify(O.Transcript(t_some.proof()), n, wrong, list(q0)) is None
