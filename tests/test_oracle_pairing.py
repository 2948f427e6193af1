"""The oracle's BN254 pairing (oracle/pairing.hpp) — the verifier side of MultilinearKzg (kzg.rs:330-361 →
`pairings_product_is_identity`, pb/util/arithmetic.rs:25-32). halo2curves' bn256 pairing is third-party code that
is not under /root/reference, so it is pinned by what characterises a pairing: G2 is on the twist and has order r,
the map is bilinear and non-degenerate, and the pairing form of `verify` accepts / rejects exactly what the
trapdoor form does (and what the provers produce)."""
import numpy as np
import pytest

import oracle as O

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
ONE = [1] + [0] * 11


def fr(v):
    return O.fr_from_ints([v % R_MOD])[0]


def test_g2_generator_on_twist_and_of_order_r():
    assert O.g2_checks(fr(0x1234567890ABCDEF1234567)) == 0


def test_bilinear_and_non_degenerate():
    a, b = 0xDEADBEEFCAFEBABE1234, 0xFEEDFACE0BADF00D5678
    e_ab = O.pairing_gen_multiples(fr(a), fr(b))
    assert e_ab == O.pairing_gen_pow(fr(a * b))           # e(aP, bQ) = e(P, Q)^(ab)
    assert e_ab == O.pairing_gen_multiples(fr(b), fr(a))  # = e(bP, aQ)
    assert e_ab == O.pairing_gen_multiples(fr(a * b), fr(1))
    g = O.pairing_gen_multiples(fr(1), fr(1))
    assert g != ONE                                        # non-degenerate
    assert O.pairing_gen_pow(fr(0)) == ONE and O.pairing_gen_pow(fr(R_MOD - 1)) != ONE
    # g^(r-1) * g = 1: the target group has order r
    assert O.pairing_gen_multiples(fr(R_MOD - 1), fr(1)) == O.pairing_gen_pow(fr(R_MOD - 1))
    assert O.pairing_gen_multiples(fr(0), fr(5)) == ONE    # identity input


def test_pairings_product_is_identity():
    a, b, c = 111111111111, 222222222222222, 3333333
    # e(aP, bQ) e(cP, Q) e(-(ab + c) P, Q) = 1
    assert O.pairing_product_is_identity([fr(a), fr(c), fr(-(a * b + c))], [fr(b), fr(1), fr(1)])
    assert not O.pairing_product_is_identity([fr(a), fr(c), fr(-(a * b + c) + 1)], [fr(b), fr(1), fr(1)])
    assert O.pairing_product_is_identity([fr(7)], [fr(0)])


@pytest.mark.parametrize("nv", [1, 3, 5])
def test_kzg_verify_pairing_form_agrees_with_trapdoor_form(nv):
    kz = O.Kzg(O.rand_fr(7, nv))
    poly, point = O.rand_fr(10 + nv, 1 << nv), O.rand_fr(20 + nv, nv)
    comm = kz.commit(poly)
    tr = O.Transcript()
    ev = kz.open(tr, poly, point)
    proof = tr.proof()
    bad_ev = O.field_op("add", ev, O.fr_from_ints([1]))[0]
    bad_proof = bytearray(proof)
    bad_proof[40] ^= 1
    other = kz.commit(O.rand_fr(99, 1 << nv))
    for pairing in (False, True):
        kz.set_pairing_check(pairing)
        assert kz.verify(O.Transcript(proof), comm, point, ev), pairing
        assert not kz.verify(O.Transcript(proof), comm, point, bad_ev), pairing
        assert not kz.verify(O.Transcript(proof), other, point, ev), pairing
        assert not kz.verify(O.Transcript(bytes(bad_proof)), comm, point, ev), pairing


def test_whole_provers_verify_with_the_pairing_check():
    """Lasso (2^4 lookups, 2^16 subtables) and HyperPlonk with lookups, verified as the reference would: SRS in G2,
    pairing product, no trapdoor."""
    from halo2_lasso_b200 import hyperplonk as H
    from halo2_lasso_b200.expression import compose

    kz = O.Kzg(O.rand_fr(7, 16))
    kz.set_pairing_check(True)
    xs = O.rand_u64s(5, 16)
    xs[8:] = xs[:8]  # repeated addresses: an all-distinct pattern has read_ts = 0, whose commitment is the identity
    tr = O.Transcript()
    assert O.lasso_prove(kz, tr, O.TABLE_RANGE, 4, 4, xs, None)
    proof = tr.proof()
    assert O.lasso_verify(kz, O.Transcript(proof), O.TABLE_RANGE, 4, 4)
    bad = bytearray(proof)
    bad[len(bad) - 40] ^= 1  # inside the last opening's quotient commitments
    assert not O.lasso_verify(kz, O.Transcript(bytes(bad)), O.TABLE_RANGE, 4, 4)

    k = 4
    info, instances, w = H.rand_vanilla_plonk_with_lookup_circuit(k, 3)
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys, lookups=info.lookups)
    hp = O.HyperPlonk(kz, k, expr, len(instances), 3, [O.fr_from_ints(p) for p in info.preprocess_polys],
                      info.permutation_polys, info.permutations, nz, lookups=info.lookups)
    inst = O.fr_from_ints(instances)
    tr = O.Transcript()
    assert hp.prove(tr, inst, [O.fr_from_ints(c) for c in w])
    assert hp.verify(O.Transcript(tr.proof()), inst)


def test_pairing_value_matches_independent_python_model():
    """oracle/pairing.hpp (2-3-2 tower, twist coordinates, sparse lines) against tests/golden/pymodel_pairing.py (single
    degree-12 extension, untwisted points, generic Fq12 line functions): e(G1, G2) and e(3 G1, 5 G2) agree coefficient
    by coefficient after the change of basis."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import pymodel_pairing as M

    want = M.pairing(M.G1, M.G2)
    got = M.tower_to_w(O.pairing_gen_multiples(fr(1), fr(1)))
    assert got == want
    assert M.tower_to_w(O.pairing_gen_multiples(fr(3), fr(5))) == M.fpow(want, 15)
    assert M.fpow(want, R_MOD) == M.ONE and want != M.ONE
