"""GPU parity of the generic-expression sum-check (bytecode round kernel) vs the oracle's restatement of
ProverState + SumCheckEvaluator, on the reference's own fixture expression (vanilla plonk zero check with
permutation argument: rotation next, Lagrange(1), identity polynomial, eq_xy, challenges)."""
import numpy as np
import pytest

import oracle as O
from halo2_lasso_b200.expression import Expression as E, vanilla_plonk_expression

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


def _run(hl, ctx, n, expr, npolys, nch, seed):
    polys = [O.rand_fr(seed + i, 1 << n) for i in range(npolys)]
    ch = O.rand_fr(seed + 100, max(nch, 1))[:nch]
    y = O.rand_fr(seed + 101, n)
    claim = O.rand_fr(seed + 102, 1)[0]
    to = O.Transcript()
    ch_o, ev_o, deg = O.sumcheck_prove_generic(to, n, expr, polys, ch, [y], claim)
    assert deg == expr.degree()
    dps = [hl.MultilinearPolynomial.new(ctx, p) for p in polys]
    # both entry points: the expression compiled inside the library (tokens cross the C ABI) and the bytecode-level
    # call fed by the Python mirror of the compiler
    for prove in (hl.prove_expression_native, hl.prove_expression):
        tr = hl.Keccak256Transcript(ctx)
        got_ch, got_ev = prove(ctx, n, expr, dps, O.fr_to_ints(ch) if nch else [], [y], claim)
        proof = tr.into_proof()
        assert len(proof) == n * (deg + 1) * 32
        assert proof == to.proof(), prove.__name__
        assert (got_ch == ch_o).all() and (got_ev == ev_o).all(), prove.__name__


@pytest.mark.parametrize("n", [2, 3, 6, 11])
def test_vanilla_plonk_zero_check_parity(hl, ctx, n):
    _run(hl, ctx, n, vanilla_plonk_expression(n), 13, 3, 5000 + n)


def test_rotation_prev_and_lagrange_last_row(hl, ctx):
    """sum_check_rotation / sum_check_lagrange shapes of pb/piop/sum_check.rs:194-260: negative rotations and
    negative Lagrange indices (rem_euclid)."""
    n = 5
    expr = (E.polynomial(0, -1) * E.polynomial(1, 2) + E.lagrange(-1) * E.polynomial(0) + E.lagrange(0) * E.polynomial(1) * 7
            - E.identity() * E.polynomial(1, 1)) * E.eq_xy(0)
    _run(hl, ctx, n, expr, 2, 0, 6000)


def test_unqueried_polynomial_is_still_bound(hl, ctx):
    """ProverState::into_evals returns EVERY polynomial bound at the challenges (classic.rs:143-149)."""
    n = 4
    expr = E.eq_xy(0) * E.polynomial(0) * E.polynomial(2, 1)
    _run(hl, ctx, n, expr, 3, 0, 7000)


def _run_given(hl, ctx, n, expr, polys, seed):
    """the reference's run_zero_check shape: claimed sum 0, one challenge (alpha), one eq point"""
    alpha, y = O.rand_fr(seed + 1, 1), O.rand_fr(seed + 2, n)
    zero = O.fr_from_ints([0])[0]
    to = O.Transcript()
    ch_o, ev_o, deg = O.sumcheck_prove_generic(to, n, expr, polys, alpha, [y], zero)
    dps = [hl.MultilinearPolynomial.new(ctx, p) for p in polys]
    for prove in (hl.prove_expression_native, hl.prove_expression):
        tr = hl.Keccak256Transcript(ctx)
        got_ch, got_ev = prove(ctx, n, expr, dps, O.fr_to_ints(alpha), [y], zero)
        assert tr.into_proof() == to.proof(), prove.__name__
        assert (got_ch == ch_o).all() and (got_ev == ev_o).all(), prove.__name__


@pytest.mark.parametrize("n", [2, 3])
def test_reference_sum_check_lagrange_shape(hl, ctx, n):
    """sum_check_lagrange (pb/piop/sum_check.rs:197-243): 2^n one-hot polynomials against Lagrange(i)"""
    from ref_shapes import reference_lagrange_case

    expr, polys = reference_lagrange_case(n)
    _run_given(hl, ctx, n, expr, polys, 8000 + n)


@pytest.mark.parametrize("n", [2, 4, 7])
def test_reference_sum_check_rotation_shape(hl, ctx, n):
    """sum_check_rotation (pb/piop/sum_check.rs:245-297): 2n - 1 polynomials queried at rotations n-1 .. -(n-1)"""
    from ref_shapes import reference_rotation_case

    expr, polys = reference_rotation_case(n, 8100 + n)
    _run_given(hl, ctx, n, expr, polys, 8200 + n)
