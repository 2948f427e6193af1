"""GPU parity: transcript, MLE and sum-check kernels through the C ABI vs the CPU oracle, byte for byte."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def hl():
    import halo2_lasso_b200 as m

    return m


@pytest.fixture(scope="module")
def ctx(hl):
    c = hl.Context(0)
    yield c
    c.close()


ONE = None


def one():
    global ONE
    if ONE is None:
        ONE = O.fr_from_ints([1])[0]
    return ONE


def test_transcript_parity(hl, ctx):
    tr = hl.Keccak256Transcript(ctx)
    to = O.Transcript()
    fes = O.rand_fr(11, 9)
    tr.common_field_elements(fes[:2])
    tr.write_field_elements(fes[2:])
    for f in fes[:2]:
        to.common_fe(f)
    for f in fes[2:]:
        to.write_fe(f)
    ch = tr.squeeze_challenges(3)
    assert (ch == to.squeeze_n(3)).all()
    g = O.g1_generator()
    pts = np.stack([O.g1_mul(g, O.fr_from_ints([k])[0]) for k in (1, 2, 12345)])
    tr.write_commitments(pts)
    for p in pts:
        to.write_comm(p)
    assert (tr.squeeze_challenge() == to.squeeze()).all()
    assert tr.into_proof() == to.proof()


def test_transcript_rejects_identity_commitment(hl, ctx):
    tr = hl.Keccak256Transcript(ctx)
    with pytest.raises(hl.B200Error) as e:
        tr.write_commitments(np.zeros((1, 8), dtype=np.uint64))
    assert e.value.code == hl.B200_ERR_TRANSCRIPT
    hl.Keccak256Transcript(ctx)  # reset clears the sticky error


@pytest.mark.parametrize("n", [1, 2, 5, 12, 13, 17])
def test_eq_xy_parity(hl, ctx, n):
    y = O.rand_fr(100 + n, n)
    got = hl.MultilinearPolynomial.eq_xy(ctx, y).evals()
    assert (got == O.eq_xy(y)).all()


@pytest.mark.parametrize("n", [1, 4, 14])
def test_fix_var_and_evaluate_parity(hl, ctx, n):
    p = O.rand_fr(200 + n, 1 << n)
    x = O.rand_fr(300 + n, n)
    dp = hl.MultilinearPolynomial.new(ctx, p)
    assert (dp.fix_var(x[0]).evals() == O.fix_var(p, x[0])).all()
    assert (dp.evaluate(x) == O.evaluate(p, x)).all()
    # boolean coordinates (the reference short-cuts them; the value must agree)
    xb = x.copy()
    xb[0] = O.fr_from_ints([1])[0]
    if n > 1:
        xb[n - 1] = O.fr_from_ints([0])[0]
    assert (dp.evaluate(xb) == O.evaluate(p, xb)).all()


def _run_evals(hl, ctx, n, T, NP, seed):
    tabs = [O.rand_fr(seed + i, 1 << n) for i in range(T * NP)]
    w = O.rand_fr(seed + 100, T) if T > 1 else one().reshape(1, 4)
    y = O.rand_fr(seed + 101, n)
    claim = O.rand_fr(seed + 102, 1)[0]  # p(0) is derived from the claim, any claim gives a transcript
    to = O.Transcript()
    terms = [(w[t], list(range(t * NP, (t + 1) * NP))) for t in range(T)]
    ch_o, ev_o = O.sumcheck_prove_evals(to, n, tabs, y, terms, claim)
    tr = hl.Keccak256Transcript(ctx)
    polys = [hl.MultilinearPolynomial.new(ctx, t) for t in tabs]
    ch, ev = hl.ClassicSumCheck.prove_evals(ctx, n, polys, w, y, claim, np_per_term=NP)
    proof = tr.into_proof()
    assert len(proof) == n * (NP + 2) * 32
    assert proof == to.proof()
    assert (ch == ch_o).all()
    assert (ev == ev_o).all()
    # the inputs are borrowed, never modified
    assert (polys[0].evals() == tabs[0]).all()


@pytest.mark.parametrize("n", [1, 2, 3, 8, 13])
def test_sumcheck_eq_a_b_parity(hl, ctx, n):
    _run_evals(hl, ctx, n, 1, 2, 1000 + n)


@pytest.mark.parametrize("n,T", [(4, 2), (9, 8), (12, 16)])
def test_sumcheck_batched_products_parity(hl, ctx, n, T):
    """grand-product layer shape: eq * Σ_t gamma^t l_t r_t"""
    _run_evals(hl, ctx, n, T, 2, 2000 + n)


@pytest.mark.parametrize("n,T", [(5, 4), (11, 8)])
def test_sumcheck_linear_g_parity(hl, ctx, n, T):
    """Surge primary shape: eq * Σ_t c_t E_t (degree 2)"""
    _run_evals(hl, ctx, n, T, 1, 3000 + n)


@pytest.mark.parametrize("n,T,NP", [(3, 1, 2), (12, 1, 2), (13, 1, 2), (14, 3, 2), (16, 2, 1), (17, 4, 2)])
def test_sumcheck_eq_factored_and_plain_kernels_agree_with_the_oracle(hl, ctx, n, T, NP):
    """The default round kernel factors eq out of the round polynomial (suffix eq tables: all-direct levels up to n = 12,
    product-form levels above); the plain kernel binds a materialised eq table. Same bytes from both (random, i.e.
    inconsistent, claims included: p(0) is derived from the claim exactly as the reference does)."""
    _run_evals(hl, ctx, n, T, NP, 5000 + n)
    hl.sumcheck_eq_factored(ctx, False)
    try:
        _run_evals(hl, ctx, n, T, NP, 5000 + n)
    finally:
        hl.sumcheck_eq_factored(ctx, True)


@pytest.mark.parametrize("n,K", [(1, 1), (6, 3), (12, 2)])
def test_sumcheck_coefficients_parity(hl, ctx, n, K):
    tabs = [O.rand_fr(4000 + n + i, 1 << n) for i in range(K)]
    sc = O.rand_fr(4100 + n, K)
    ys = np.stack([O.rand_fr(4200 + n + k, n) for k in range(K)])
    claim = O.rand_fr(4300 + n, 1)[0]
    to = O.Transcript()
    ch_o, ev_o = O.sumcheck_prove_coeffs(to, n, tabs, [(sc[k], ys[k], k) for k in range(K)], claim)
    tr = hl.Keccak256Transcript(ctx)
    polys = [hl.MultilinearPolynomial.new(ctx, t) for t in tabs]
    ch, ev = hl.ClassicSumCheck.prove_coeffs(ctx, n, polys, sc, ys, claim)
    assert tr.into_proof() == to.proof()
    assert (ch == ch_o).all() and (ev == ev_o).all()


def test_sumcheck_cfg2_full_size_and_verifies(hl, ctx):
    """BASELINE cfg2: degree-3 eq*a*b over 20 variables: 2560 proof bytes, byte-identical, and the
    proof verifies (reference test shape pb/piop/sum_check.rs:140-177)."""
    n = 20
    a, b, y = O.rand_fr(1, 1 << n), O.rand_fr(2, 1 << n), O.rand_fr(3, n)
    s = O.sum_eq_ab(y, a, b)
    to = O.Transcript()
    ch_o, ev_o = O.sumcheck_prove_evals(to, n, [a, b], y, [(one(), [0, 1])], s)
    tr = hl.Keccak256Transcript(ctx)
    ch, ev = hl.ClassicSumCheck.prove_evals_host(ctx, n, [a, b], one().reshape(1, 4), y, s)
    proof = tr.into_proof()
    assert len(proof) == 2560
    assert proof == to.proof()
    assert (ch == ch_o).all() and (ev == ev_o).all()
    fin, ch_v = O.sumcheck_verify(O.Transcript(proof), n, 3, s)
    assert (ch_v == ch).all()
    exp = O.field_op("mul", O.field_op("mul", O.eq_xy_eval(ch, y), ev[0]), ev[1])[0]
    assert (exp == fin).all()
