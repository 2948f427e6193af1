"""Oracle HyperPlonk (lookup-free vanilla plonk): the reference's e2e test shape (pb/backend.rs:202-241,
hyperplonk.rs:398-426): preprocess -> prove -> verify accepts; tampering / wrong instances are rejected. Also the
product's host-side helpers (fixture, permutation_polys, rotation_eval_points) against the oracle."""
import numpy as np
import pytest

import oracle as O
from halo2_lasso_b200.expression import BooleanHypercube, R_MOD
from halo2_lasso_b200 import hyperplonk as H


@pytest.fixture(scope="module")
def kz():
    return O.Kzg(O.rand_fr(7, 8))


def build(kz, k, seed):
    info, instances, w = H.rand_vanilla_plonk_circuit(k, seed)
    from halo2_lasso_b200.expression import compose

    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys)
    hp = O.HyperPlonk(kz, k, expr, len(instances), 3, [O.fr_from_ints(p) for p in info.preprocess_polys],
                      info.permutation_polys, info.permutations, nz)
    return info, instances, w, hp


def test_fixture_is_satisfiable_and_permutation_consistent():
    k = 5
    info, instances, w = H.rand_vanilla_plonk_circuit(k, 11)
    q = info.preprocess_polys
    order = BooleanHypercube(k).iter()
    pi = [0] * (1 << k)
    for i, v in enumerate(instances):
        pi[order[i + 1]] = v
    for b in range(1 << k):
        g = q[0][b] * w[0][b] + q[1][b] * w[1][b] + q[2][b] * w[0][b] * w[1][b] + q[3][b] * w[2][b] + q[4][b] + pi[b]
        assert g % R_MOD == 0, b
    cols = {6: w[0], 7: w[1], 8: w[2]}
    assert info.permutations, "fixture should contain copy constraints"
    for cyc in info.permutations:
        assert len({cols[p][r] for p, r in cyc}) == 1


@pytest.mark.parametrize("k", [3, 4, 6])
def test_prove_verify_roundtrip_and_negatives(kz, k):
    info, instances, w, hp = build(kz, k, 20 + k)
    # host permutation_polys == oracle's
    mine = H.permutation_polys(k, info.permutation_polys, info.permutations)
    for i in range(3):
        assert O.fr_to_ints(hp.permutation_poly(i)) == mine[i]
    inst = O.fr_from_ints(instances)
    tr = O.Transcript()
    assert hp.prove(tr, inst, [O.fr_from_ints(c) for c in w])
    proof = tr.proof()
    assert hp.verify(O.Transcript(proof), inst)
    for pos in (5, len(proof) // 3, len(proof) - 7):
        bad = bytearray(proof)
        bad[pos] ^= 2
        assert not hp.verify(O.Transcript(bytes(bad)), inst)
    wrong = inst.copy()
    wrong[0] = O.field_op("add", wrong[0], O.fr_from_ints([1]))[0]
    assert not hp.verify(O.Transcript(proof), wrong)
    # an unsatisfied gate must not verify
    w_bad = [list(c) for c in w]
    w_bad[2][(1 << k) - 1] = (w_bad[2][(1 << k) - 1] + 1) % R_MOD
    tr2 = O.Transcript()
    hp.prove(tr2, inst, [O.fr_from_ints(c) for c in w_bad])
    assert not hp.verify(O.Transcript(tr2.proof()), inst)


def test_rotation_eval_points_define_the_rotated_evaluation():
    """z(next)(x) recombined from the two evaluations at rotation_eval_points equals the MLE of the rotated
    table (verifier-side `rotation_eval`, multilinear.rs:433-473, for Rotation::next)."""
    k = 6
    table = O.rand_fr(5, 1 << k)
    x_m = O.rand_fr(6, k)
    x = O.fr_to_ints(x_m)
    pts = H.rotation_eval_points(x, 1)
    assert len(pts) == 2 and pts[0][0] == 0 and pts[1][0] == 1
    e0, e1 = (O.fr_to_ints(O.evaluate(table, O.fr_from_ints(p)))[0] for p in pts)
    combined = (e0 + x[k - 1] * (e1 - e0)) % R_MOD
    bh = BooleanHypercube(k)
    rotated = table[[bh.rotate(b, 1) for b in range(1 << k)]]
    assert combined == O.fr_to_ints(O.evaluate(rotated, x_m))[0]
