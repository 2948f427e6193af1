"""GKR for fractional sum-checks on the GPU (b200_fractional_sum_check_prove) against the CPU oracle and against the
committed fixtures of the independent pure-Python model: byte-identical proof streams, identical claims and point;
the reference's own test (`fractional_sum_check`, fractional_sum_check.rs:330-370) restated on the GPU prover."""
import json
import os

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "gkr_golden.json")))


@pytest.fixture(scope="module")
def hl():
    import halo2_lasso_b200 as m

    return m


@pytest.fixture(scope="module")
def ctx(hl):
    c = hl.Context(0)
    yield c
    c.close()


def gpu_prove(hl, ctx, ps, qs, cl_p=None, cl_q=None):
    tr = hl.Keccak256Transcript(ctx)
    dp = [hl.MultilinearPolynomial.new(ctx, p) for p in ps]
    dq = [hl.MultilinearPolynomial.new(ctx, q) for q in qs]
    out = hl.fractional_sum_check_prove(ctx, dp, dq, cl_p, cl_q)
    return tr.into_proof(), out


@pytest.mark.parametrize("case", G["cases"], ids=lambda c: f"b{c['batch']}-n{c['num_vars']}")
def test_golden_bytes_of_the_python_model(hl, ctx, case):
    B, n, seed = case["batch"], case["num_vars"], case["seed"]
    ps = [O.rand_fr(seed + b, 1 << n) for b in range(B)]
    qs = [O.rand_fr(seed + 50 + b, 1 << n) for b in range(B)]
    cl = [0] * B if case["claimed"] else None
    proof, (p_xs, q_xs, x, p0, q0) = gpu_prove(hl, ctx, ps, qs, cl, cl)
    assert proof.hex() == case["proof"]
    for got, key in ((p_xs, "p_xs"), (q_xs, "q_xs"), (x, "x"), (p0, "p_0s"), (q0, "q_0s")):
        assert O.fr_to_ints(got) == [int(v) for v in case[key]]


@pytest.mark.parametrize("B,n", [(3, 1), (3, 2), (3, 3), (1, 9), (3, 12), (10, 10), (2, 16), (3, 18)])
def test_parity_with_the_oracle_and_the_reference_test(hl, ctx, B, n):
    ps = [O.rand_fr(5000 + 31 * n + b, 1 << n) for b in range(B)]
    qs = [O.rand_fr(5500 + 31 * n + b, 1 << n) for b in range(B)]
    to = O.Transcript()
    want = O.fractional_sum_check_prove(to, ps, qs)
    proof, got = gpu_prove(hl, ctx, ps, qs)
    assert proof == to.proof(), "GPU transcript differs from the oracle"
    for g, w in zip(got, want):
        assert (np.asarray(g) == np.asarray(w)).all()
    # the reference's test: the verifier accepts and the claims are the inputs' evaluations at x
    res = O.fractional_sum_check_verify(O.Transcript(proof), n, [None] * B, [None] * B)
    assert res is not None
    # ... and the product's own CPU verifier (libb200verify.so) returns the same claims and point
    from halo2_lasso_b200 import verifier as V

    vt = V.ProofTranscript(proof)
    vres = V.fractional_sum_check_verify(vt, n, [None] * B, [None] * B)
    assert vres is not None and vt.done()
    for g, w in zip(vres, got):
        assert (np.asarray(g) == np.asarray(w)).all()
    p_xs, q_xs, x = got[0], got[1], got[2]
    for b in range(B):
        assert (O.evaluate(ps[b], x) == p_xs[b]).all() and (O.evaluate(qs[b], x) == q_xs[b]).all()


def test_claimed_values_and_argument_errors(hl, ctx):
    B, n = 2, 7
    ps = [O.rand_fr(5900 + b, 1 << n) for b in range(B)]
    qs = [O.rand_fr(5950 + b, 1 << n) for b in range(B)]
    to = O.Transcript()
    O.fractional_sum_check_prove(to, ps, qs, [0, None], [None, 0])
    proof, _ = gpu_prove(hl, ctx, ps, qs, [0, None], [None, 0])
    assert proof == to.proof()
    dp = [hl.MultilinearPolynomial.new(ctx, p) for p in ps]
    with pytest.raises(hl.B200Error):
        hl.fractional_sum_check_prove(ctx, dp * 6, dp * 6)  # more than 10 batch elements
