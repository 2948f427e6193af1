"""Oracle HyperPlonk (vanilla plonk, without and with the LogUp lookup argument): the reference's e2e test shape (pb/backend.rs:202-241,
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


# ---- vanilla plonk WITH the LogUp lookup argument (util.rs:216-316, prover.rs:50-250) ---------------------------
def build_lookup(kz, k, seed):
    from halo2_lasso_b200.expression import compose

    info, instances, w = H.rand_vanilla_plonk_with_lookup_circuit(k, seed)
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys, lookups=info.lookups)
    hp = O.HyperPlonk(kz, k, expr, len(instances), 3, [O.fr_from_ints(p) for p in info.preprocess_polys],
                      info.permutation_polys, info.permutations, nz, lookups=info.lookups)
    return info, instances, w, hp, expr


def test_lookup_fixture_rows_are_in_the_table():
    k = 6
    info, instances, w = H.rand_vanilla_plonk_with_lookup_circuit(k, 5)
    q = info.preprocess_polys
    rows = {(q[6][b], q[7][b], q[8][b]) for b in range(1 << k)}
    n_lookup = 0
    for b in range(1 << k):
        tup = tuple(q[5][b] * w[i][b] % R_MOD for i in range(3))
        assert tup in rows
        n_lookup += q[5][b]
    assert n_lookup > 4


def test_lookup_expression_shape():
    """util.rs:88-98: 13 + 3 sigma + m + h + z = 19 polynomials, degree 5 with the eq factor (BASELINE cfg5 shape)."""
    from halo2_lasso_b200.expression import compose

    info, _, _ = H.rand_vanilla_plonk_with_lookup_circuit(4, 1)
    nz, expr = compose(4, info.constraints, info.num_poly, info.permutation_polys, lookups=info.lookups)
    assert nz == 1 and expr.degree() == 5
    polys = {l[1] for l in expr.leaves() if l[0] == "poly"}
    assert polys == set(range(19))
    assert ("poly", 18, 1) in expr.leaves()  # z(next)


def test_lookup_m_counts_on_the_last_duplicate_row_and_h_sums_to_zero():
    k = 5
    N = 1 << k
    table = O.rand_fr(3, N)
    table[0] = table[1] = 0  # the zero tuple twice, as in the reference fixture
    table[7] = table[20]     # another duplicate: row 20 wins
    idx = [1, 1, 20, 20, 20, 5] + [int(x) % N for x in O.rand_u64s(9, N - 6)]
    inp = table[idx]
    m = O.fr_to_ints(O.lookup_m(inp, table))
    exp = [0] * N
    last = {}
    for i in range(N):
        last[tuple(int(x) for x in table[i])] = i
    for i in idx:
        exp[last[tuple(int(x) for x in table[i])]] += 1
    assert m == exp and m[0] == 0 and m[7] == 0 and m[20] >= 3
    gamma = O.rand_fr(4, 1)[0]
    h = O.fr_to_ints(O.lookup_h(inp, table, O.fr_from_ints(m), gamma))
    assert sum(h) % R_MOD == 0
    # an input that is not in the table is rejected
    bad = inp.copy()
    bad[3] = O.rand_fr(77, 1)[0]
    assert O.lookup_m(bad, table) is None


def test_expression_rows_matches_python_evaluation():
    from halo2_lasso_b200.expression import Expression as E

    k = 4
    N = 1 << k
    polys = [O.rand_fr(30 + i, N) for i in range(3)]
    ints = [O.fr_to_ints(p) for p in polys]
    ch = [12345, 678]
    expr = E.polynomial(0) * E.polynomial(1, 1) + E.challenge(1) * E.polynomial(2, -1) - E.identity() * E.lagrange(2) + E.constant(9)
    got = O.fr_to_ints(O.expression_rows(k, expr, polys, O.fr_from_ints(ch)))
    bh = BooleanHypercube(k)
    for b in range(N):
        want = (ints[0][b] * ints[1][bh.rotate(b, 1)] + ch[1] * ints[2][bh.rotate(b, -1)] - b * (1 if b == bh.nth(2) else 0) + 9) % R_MOD
        assert got[b] == want, b


@pytest.mark.parametrize("k", [3, 5])
def test_lookup_prove_verify_roundtrip_and_negatives(kz, k):
    info, instances, w, hp, expr = build_lookup(kz, k, 60 + k)
    inst = O.fr_from_ints(instances)
    tr = O.Transcript()
    assert hp.prove(tr, inst, [O.fr_from_ints(c) for c in w])
    proof = tr.proof()
    assert hp.verify(O.Transcript(proof), inst)
    for pos in (9, len(proof) // 2, len(proof) - 3):
        bad = bytearray(proof)
        bad[pos] ^= 4
        assert not hp.verify(O.Transcript(bytes(bad)), inst)
    # a lookup row whose tuple is not a table row: the prover fails with "Invalid lookup input"
    q = info.preprocess_polys
    row = next(b for b in range(1 << k) if q[5][b])
    w_bad = [list(c) for c in w]
    w_bad[1][row] = (w_bad[1][row] + 1) % R_MOD
    assert not hp.prove(O.Transcript(), inst, [O.fr_from_ints(c) for c in w_bad])


@pytest.mark.parametrize("lookup,max_degree", [(False, 4), (True, 4), (False, 3)])
def test_hyperplonk_proof_matches_independent_python_model(lookup, max_degree):
    """Whole proofs (vanilla plonk; with the LogUp lookup argument; with two permutation chunks) against the pure-Python
    model tests/golden/pymodel_hyperplonk.py — tree-walking expression evaluation, big-int field arithmetic, affine
    curve arithmetic, naive MSM — byte for byte."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import pymodel as M
    import pymodel_hyperplonk as MH
    from halo2_lasso_b200.expression import compose

    k = 3
    ss = M.rand_fr(7, k)
    info, instances, w = (H.rand_vanilla_plonk_with_lookup_circuit if lookup else H.rand_vanilla_plonk_circuit)(k, 31, num_instances=2)
    want = MH.prove(M.kzg_setup(ss), info, instances, w, max_degree=max_degree)
    kz = O.Kzg(O.fr_from_ints(ss))
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys, max_degree=max_degree, lookups=info.lookups)
    hp = O.HyperPlonk(kz, k, expr, len(instances), 3, [O.fr_from_ints(p) for p in info.preprocess_polys],
                      info.permutation_polys, info.permutations, nz, lookups=info.lookups)
    tr = O.Transcript()
    assert hp.prove(tr, O.fr_from_ints(instances), [O.fr_from_ints(c) for c in w])
    assert tr.proof() == want
    assert hp.verify(O.Transcript(want), O.fr_from_ints(instances))


# ---- several instance columns, several witness phases (pb/backend.rs:50-60, hyperplonk.rs:183-204) -----------
def two_phase_oracle(kz, k, seed, with_lookup=True):
    from halo2_lasso_b200.expression import compose

    info, inst_cols, synth = H.rand_two_phase_circuit(k, seed, with_lookup)
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys,
                       num_challenges=sum(info.num_challenges), lookups=info.lookups)
    hp = O.HyperPlonk(kz, k, expr, info.num_instances, info.num_witness_polys, [O.fr_from_ints(p) for p in info.preprocess_polys],
                      info.permutation_polys, info.permutations, nz, lookups=info.lookups, num_challenges=info.num_challenges)
    rounds = []

    def synth_fr(rnd, challenges):
        rounds.append((rnd, len(challenges)))
        return [O.fr_from_ints(c) for c in synth(rnd, O.fr_to_ints(challenges) if len(challenges) else [])]

    inst = O.fr_from_ints([v for col in inst_cols for v in col])
    return info, inst_cols, synth, hp, synth_fr, inst, rounds


def test_two_phase_fixture_is_satisfiable():
    k = 5
    info, inst_cols, synth = H.rand_two_phase_circuit(k, 3)
    N, bh = 1 << k, BooleanHypercube(k)
    order = bh.iter()
    r0, r1 = 1234567, 7654321
    (a, b), (c, d) = synth(0, []), synth(1, [r0, r1])
    assert synth(1, [r0, r1]) == [c, d], "synthesize must be deterministic"
    q_mix, q_io, q_io2, q_mul, q_lk, t = info.preprocess_polys
    pi = [[0] * N, [0] * N]
    for col, vals in enumerate(inst_cols):
        for i, v in enumerate(vals):
            pi[col][order[i + 1]] = v
    table = set(t)
    for row in range(N):
        assert q_mix[row] * (a[row] + r0 * b[row] - c[row]) % R_MOD == 0
        assert q_io[row] * (a[row] - pi[0][row]) % R_MOD == 0
        assert q_io2[row] * (b[row] - pi[1][bh.rotate(row, 1)]) % R_MOD == 0
        assert q_mul[row] * (r1 * a[row] * b[row] - d[row]) % R_MOD == 0
        assert q_lk[row] * (c[row] - r0 * b[row]) % R_MOD in table
    cols = {8: a, 9: b, 10: c}
    assert sum(q_io) == 2 and sum(q_io2) == 3 and sum(q_lk) > 0 and len(info.permutations) >= 2
    for cyc in info.permutations:
        assert len({cols[p][r] for p, r in cyc}) == 1


@pytest.mark.parametrize("k,with_lookup", [(4, True), (4, False), (6, True)])
def test_two_phase_prove_verify_roundtrip_and_negatives(kz, k, with_lookup):
    info, inst_cols, synth, hp, synth_fr, inst, rounds = two_phase_oracle(kz, k, 50 + k, with_lookup)
    tr = O.Transcript()
    assert hp.prove_phased(tr, inst, synth_fr)
    assert rounds == [(0, 0), (1, 2)]  # phase 1 sees the two challenges squeezed after the phase-0 commitments
    proof = tr.proof()
    assert hp.verify(O.Transcript(proof), inst)
    # proof layout: 4 witness commitments (2 + 2), then m (if any), h + z, sum-check, evaluations, opening
    for pos in (5, 64 * 3 + 9, len(proof) // 2, len(proof) - 7):
        bad = bytearray(proof)
        bad[pos] ^= 4
        assert not hp.verify(O.Transcript(bytes(bad)), inst)
    for j in range(inst.shape[0]):  # every instance of BOTH columns is bound (pi_b through its rotated query)
        wrong = inst.copy()
        wrong[j] = O.rand_fr(99, 1)[0]
        assert not hp.verify(O.Transcript(proof), wrong)
    # instance columns of the wrong shape (hyperplonk.rs:171-173 assert / :299-305 Error::InvalidSnark)
    assert not hp.verify(O.Transcript(proof), inst[:-1])
    # a phase-1 witness that ignores the challenges does not verify
    stale = lambda rnd, ch: synth_fr(rnd, O.fr_from_ints([1, 2]) if rnd else ch)
    tr2 = O.Transcript()
    ok = hp.prove_phased(tr2, inst, stale)
    assert not (ok and hp.verify(O.Transcript(tr2.proof()), inst))
    # a synthesize callback returning the wrong number of polynomials fails the prover
    assert not hp.prove_phased(O.Transcript(), inst, lambda rnd, ch: synth_fr(rnd, ch)[:1])


@pytest.mark.parametrize("with_lookup", [True, False])
def test_two_phase_proof_matches_independent_python_model(with_lookup):
    """Two instance columns (one queried at Rotation::next), two witness phases with circuit challenges (also inside a
    lookup input): the oracle's proof equals the pure-Python model's byte for byte."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import pymodel as M
    import pymodel_hyperplonk as MH

    k = 4
    ss = M.rand_fr(7, k)
    kz = O.Kzg(O.fr_from_ints(ss))
    info, inst_cols, synth, hp, synth_fr, inst, rounds = two_phase_oracle(kz, k, 61, with_lookup)
    want = MH.prove(M.kzg_setup(ss), info, inst_cols, synth)
    tr = O.Transcript()
    assert hp.prove_phased(tr, inst, synth_fr)
    assert tr.proof() == want
    assert hp.verify(O.Transcript(want), inst)


def test_hyperplonk_proofs_match_committed_golden_bytes():
    """tests/golden/hyperplonk_golden.json (written by make_golden_hyperplonk.py from the pure-Python model): the oracle
    reproduces the committed proof bytes of all five circuits and its verifier accepts them."""
    import json
    import os

    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import pymodel as M
    from halo2_lasso_b200.expression import compose

    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hyperplonk_golden.json")))
    assert len(gold["cases"]) == 5
    for c in gold["cases"]:
        k, want = c["k"], bytes.fromhex(c["proof"])
        kz = O.Kzg(O.fr_from_ints(M.rand_fr(c["srs_seed"], k)))
        tr = O.Transcript()
        if c["circuit"] == "two_phase":
            info, inst_cols, synth, hp, synth_fr, inst, _ = two_phase_oracle(kz, k, c["seed"], c["with_lookup"])
            assert hp.prove_phased(tr, inst, synth_fr)
        else:
            fixture = H.rand_vanilla_plonk_with_lookup_circuit if c["circuit"].endswith("lookup") else H.rand_vanilla_plonk_circuit
            info, instances, w = fixture(k, c["seed"], num_instances=2)
            nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys, max_degree=c["max_degree"], lookups=info.lookups)
            hp = O.HyperPlonk(kz, k, expr, len(instances), 3, [O.fr_from_ints(p) for p in info.preprocess_polys],
                              info.permutation_polys, info.permutations, nz, lookups=info.lookups)
            inst = O.fr_from_ints(instances)
            assert hp.prove(tr, inst, [O.fr_from_ints(col) for col in w])
        assert tr.proof() == want, c["circuit"]
        assert hp.verify(O.Transcript(want), inst)
