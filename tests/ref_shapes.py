"""The reference's own generic sum-check test shapes (pb/piop/sum_check.rs:194-300) as (expression, polynomials)
builders shared by the oracle tests and the GPU parity tests."""
import numpy as np

import oracle as O


def reference_lagrange_case(n):
    """sum_check_lagrange: 2^n one-hot polynomials in BooleanHypercube order, gates Lagrange(i) - poly_i."""
    from halo2_lasso_b200.expression import Expression as E

    N = 1 << n
    gates = [E.lagrange(i) - E.polynomial(i) for i in range(N)]
    expr = E.distribute_powers(gates, E.challenge(0)) * E.eq_xy(0)
    order = [int(b) for b in O.bh_iter(n)]
    polys = []
    for b in order:
        p = [0] * N
        p[b] = 1
        polys.append(O.fr_from_ints(p))
    return expr, polys


def reference_rotation_case(n, seed):
    """sum_check_rotation: polynomial idx is the (idx)-fold Rotation::next image of a random one and is queried at
    rotation n - 1 - idx, so that consecutive queries agree on every row."""
    from halo2_lasso_b200.expression import Expression as E

    N = 1 << n
    qs = [E.polynomial(idx, rot) for idx, rot in enumerate(reversed(range(-n + 1, n)))]
    gates = [qs[i + 1] - qs[i] for i in range(len(qs) - 1)]
    expr = E.distribute_powers(gates, E.challenge(0)) * E.eq_xy(0)
    base = O.rand_fr(seed, N)
    polys = [base]
    for _ in range(2 * n - 2):
        prev = polys[-1]
        polys.append(np.ascontiguousarray(prev[[O.bh_rotate(n, b, 1) for b in range(N)]]))
    return expr, polys
