"""The product's CPU verifier (libb200verify.so, include/b200_verify.h; host code, runs without a GPU): it accepts the
proofs of the oracle provers and the committed golden proofs of the pure-Python models, rejects tampered proofs and
wrong statements, and reports malformed parameters as argument errors. Mirrors the accept / reject round trips of the
reference's own tests (pb/piop/sum_check.rs:140-177, pb/pcs/multilinear.rs:293-406, pb/backend.rs:202-241)."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import oracle as O
from halo2_lasso_b200 import hyperplonk as H
from halo2_lasso_b200 import verifier as V
from halo2_lasso_b200.expression import compose

HERE = os.path.dirname(os.path.abspath(__file__))
SS = O.rand_fr(7, 16)


@pytest.fixture(scope="module")
def okzg():
    return O.Kzg(SS)


@pytest.fixture(scope="module")
def vkzg():
    return V.MultilinearKzgVerifier.setup(SS)


def tampered(proof, pos, bit=1):
    bad = bytearray(proof)
    bad[pos] ^= bit
    return bytes(bad)


def test_library_exports_exactly_the_declared_symbols():
    out = subprocess.run(["nm", "-D", "--defined-only", V.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = sorted(set(re.findall(r"\b(b200v_[a-z0-9_]+)\b", out)))
    assert exported == V.declared_symbols() and len(exported) == 21


def test_transcript_reading_side_matches_the_oracle():
    to = O.Transcript()
    fes = O.rand_fr(11, 5)
    for f in fes[:4]:
        to.write_fe(f)
    c1 = to.squeeze()
    to.common_fe(fes[4])
    pt = O.g1_mul(O.g1_generator(), O.fr_from_ints([5])[0])
    to.write_comm(pt)
    c2 = to.squeeze()
    tr = V.ProofTranscript(to.proof())
    assert (tr.read_field_elements(4) == fes[:4]).all()
    assert (tr.squeeze_challenges(1)[0] == c1).all()
    tr.common_field_elements(fes[4:5])
    assert (tr.read_commitments(1)[0] == pt).all()
    assert (tr.squeeze_challenges(1)[0] == c2).all() and tr.done()
    # non-canonical field element (>= r) and a point off the curve are rejected (transcript.rs:146-152, 198-208)
    with pytest.raises(ValueError):
        V.ProofTranscript(b"\xff" * 32).read_field_elements(1)
    with pytest.raises(ValueError):
        V.ProofTranscript(tampered(to.proof()[4 * 32:], 40)).read_commitments(1)
    with pytest.raises(ValueError):
        V.ProofTranscript(b"\x01" * 31).read_field_elements(1)  # truncated


@pytest.mark.parametrize("n", [1, 5, 9])
def test_sumcheck_verify_evaluations_and_coefficients(n):
    a, b, y = O.rand_fr(1, 1 << n), O.rand_fr(2, 1 << n), O.rand_fr(3, n)
    s = O.sum_eq_ab(y, a, b)
    one = O.fr_from_ints([1])[0]
    to = O.Transcript()
    ch, ev = O.sumcheck_prove_evals(to, n, [a, b], y, [(one, [0, 1])], s)
    fin_o, _ = O.sumcheck_verify(O.Transcript(to.proof()), n, 3, s)
    tr = V.ProofTranscript(to.proof())
    fin, x = V.sumcheck_verify(tr, n, 3, s)
    assert tr.done() and (x == ch).all() and (fin == fin_o).all()
    # the caller's final check (sum_check.rs:166-176): claim == eq(x, y) * a(x) * b(x)
    assert (O.field_op("mul", O.field_op("mul", O.eq_xy_eval(x, y), ev[0]), ev[1])[0] == fin).all()
    assert V.sumcheck_verify(V.ProofTranscript(tampered(to.proof(), 33)), n, 3, s) is None
    assert V.sumcheck_verify(V.ProofTranscript(to.proof()), n, 3, O.field_op("add", s, one)[0]) is None
    # CoefficientsProver messages
    scal, ys = O.rand_fr(5, 2), [O.rand_fr(6, n), O.rand_fr(7, n)]
    claim = O.rand_fr(8, 1)[0]
    tc = O.Transcript()
    O.sumcheck_prove_coeffs(tc, n, [a, b], [(scal[0], ys[0], 0), (scal[1], ys[1], 1)], claim)
    want = O.sumcheck_verify(O.Transcript(tc.proof()), n, 2, claim, coeffs=True)
    got = V.sumcheck_verify(V.ProofTranscript(tc.proof()), n, 2, claim, coefficients_form=True)
    assert (want is None) == (got is None)
    if got is not None:
        assert (got[0] == want[0]).all() and (got[1] == want[1]).all()
    with pytest.raises(V.VerifierArgError):
        V.sumcheck_verify(V.ProofTranscript(to.proof()), 0, 3, s)


def test_kzg_open_and_batch_open_verify(okzg, vkzg):
    nv = 6
    poly, point = O.rand_fr(600, 1 << nv), O.rand_fr(601, nv)
    comm = okzg.commit(poly)
    to = O.Transcript()
    ev = okzg.open(to, poly, point)
    assert vkzg.verify(V.ProofTranscript(to.proof()), comm, point, ev)
    assert not vkzg.verify(V.ProofTranscript(to.proof()), comm, point, O.field_op("add", ev, O.fr_from_ints([1]))[0])
    assert not vkzg.verify(V.ProofTranscript(tampered(to.proof(), 70)), comm, point, ev)
    assert not vkzg.verify(V.ProofTranscript(to.proof()), okzg.commit(O.rand_fr(602, 1 << nv)), point, ev)
    # batch: 4 polynomials, 2 points, 5 (poly, point) pairs (pb/pcs/multilinear.rs:345-406 shape)
    polys = [O.rand_fr(900 + i, 1 << nv) for i in range(4)]
    points = [O.rand_fr(950 + i, nv) for i in range(2)]
    evals = [(p, q, O.evaluate(polys[p], points[q])) for p, q in [(0, 0), (1, 1), (2, 1), (3, 0), (0, 1)]]
    comms = [okzg.commit(p) for p in polys]
    tb = O.Transcript()
    okzg.batch_open(tb, polys, points, evals)
    assert vkzg.batch_verify(V.ProofTranscript(tb.proof()), comms, points, evals)
    bad = list(evals)
    bad[2] = (2, 1, O.rand_fr(1, 1)[0])
    assert not vkzg.batch_verify(V.ProofTranscript(tb.proof()), comms, points, bad)
    assert not vkzg.batch_verify(V.ProofTranscript(tampered(tb.proof(), len(tb.proof()) - 5)), comms, points, evals)
    with pytest.raises(V.VerifierArgError):
        vkzg.batch_verify(V.ProofTranscript(tb.proof()), comms, points, [(7, 0, evals[0][2]), evals[1]])
    # parameters travel as G2 points: export / import round trip verifies the same opening
    again = V.MultilinearKzgVerifier.from_g2_powers(vkzg.g2_powers())
    assert again.verify(V.ProofTranscript(to.proof()), comm, point, ev)
    broken = vkzg.g2_powers()
    broken[0, 0] ^= 1
    with pytest.raises(V.VerifierArgError):
        V.MultilinearKzgVerifier.from_g2_powers(broken)


@pytest.mark.parametrize("kind,chunks,mu", [(O.TABLE_RANGE, 4, 5), (O.TABLE_AND, 8, 4), (O.TABLE_XOR, 2, 6)])
def test_lasso_verify_accepts_oracle_proofs_and_rejects_tampering(okzg, vkzg, kind, chunks, mu):
    xs, ys = O.rand_u64s(70 + mu, 1 << mu), O.rand_u64s(80 + mu, 1 << mu)
    if kind == O.TABLE_RANGE:
        ys = None
    else:
        xs &= np.uint64((1 << (8 * chunks)) - 1)
        ys &= np.uint64((1 << (8 * chunks)) - 1)
        ys[1::2] = ys[0::2]
    xs[1::2] = xs[0::2]  # repeated addresses: read_ts != 0 (an all-zero polynomial commits to the identity, which the
    to = O.Transcript()  # transcript cannot encode, transcript.rs:174-179)
    assert O.lasso_prove(okzg, to, kind, chunks, mu, xs, ys)
    proof = to.proof()
    tr = V.ProofTranscript(proof)
    assert vkzg.lasso_verify(tr, kind, chunks, mu) and tr.done()
    for pos in (7, len(proof) // 3, len(proof) // 2, len(proof) - 9):
        assert not vkzg.lasso_verify(V.ProofTranscript(tampered(proof, pos, 4)), kind, chunks, mu)
    assert not vkzg.lasso_verify(V.ProofTranscript(proof), kind, chunks, mu + 1)
    other = O.TABLE_XOR if kind == O.TABLE_AND else (O.TABLE_AND if kind == O.TABLE_XOR else None)
    if other is not None:
        assert not vkzg.lasso_verify(V.ProofTranscript(proof), other, chunks, mu)
    with pytest.raises(V.VerifierArgError):
        vkzg.lasso_verify(V.ProofTranscript(proof), 3, chunks, mu)


def test_lasso_verify_binds_the_proof_to_the_callers_statement(okzg, vkzg):
    """ACCEPT without a statement only says that SOME committed a decomposes into table entries (any prover can do that,
    e.g. with other lookups): with the expected commitments a proof for different lookups is rejected."""
    kind, chunks, mu = O.TABLE_AND, 4, 5
    mask = np.uint64((1 << (8 * chunks)) - 1)

    def instance(seed):
        xs, ys = O.rand_u64s(seed, 1 << mu) & mask, O.rand_u64s(seed + 1, 1 << mu) & mask
        xs[1::2], ys[1::2] = xs[0::2], ys[0::2]
        to = O.Transcript()
        assert O.lasso_prove(okzg, to, kind, chunks, mu, xs, ys)
        mt, _ = O.lasso_witness(kind, chunks, mu, xs, ys)
        return to.proof(), okzg.commit(mt[0]), [okzg.commit(mt[1 + t]) for t in range(chunks)]

    proof, com_a, com_dims = instance(900)
    other_proof, other_a, other_dims = instance(950)
    tr = V.ProofTranscript(proof)
    ok, comms = vkzg.lasso_verify(tr, kind, chunks, mu, expect_a=com_a, expect_dims=com_dims, want_commitments=True)
    assert ok and tr.done()
    assert (comms[0] == com_a).all() and all((comms[1 + t] == com_dims[t]).all() for t in range(chunks))
    # a perfectly valid proof — of other lookups — does not satisfy this statement
    assert vkzg.lasso_verify(V.ProofTranscript(other_proof), kind, chunks, mu)
    assert not vkzg.lasso_verify(V.ProofTranscript(other_proof), kind, chunks, mu, expect_a=com_a)
    assert not vkzg.lasso_verify(V.ProofTranscript(other_proof), kind, chunks, mu, expect_dims=com_dims)
    assert not vkzg.lasso_verify(V.ProofTranscript(proof), kind, chunks, mu, expect_a=other_a)
    assert vkzg.lasso_verify(V.ProofTranscript(other_proof), kind, chunks, mu, expect_a=other_a, expect_dims=other_dims)
    off_curve = com_a.copy()
    off_curve[0] ^= 1
    with pytest.raises(V.VerifierArgError):
        vkzg.lasso_verify(V.ProofTranscript(proof), kind, chunks, mu, expect_a=off_curve)


def test_lasso_verify_accepts_the_committed_golden_proofs(vkzg):
    gold = json.load(open(os.path.join(HERE, "golden", "lasso_golden.json")))
    for c in gold["cases"]:
        assert c["srs_seed"] == 7
        tr = V.ProofTranscript(bytes.fromhex(c["proof"]))
        if c["kind"] == 3:
            t = c["table"]
            tab = _or8_table(c["chunks"], t["out_bits"])
            assert vkzg.lasso_verify_table(tr, tab, c["mu"]) and tr.done()
            continue
        assert vkzg.lasso_verify(tr, c["kind"], c["chunks"], c["mu"]) and tr.done()


class _Tab:  # the fields lasso_verify_table reads (halo2_lasso_b200.LassoTable needs the CUDA library to upload)
    def __init__(self, chunks, num_operands, operand_bits, out_bits, values):
        self.chunks, self.num_operands, self.operand_bits, self.out_bits = chunks, num_operands, operand_bits, out_bits
        self.values = np.ascontiguousarray(values, dtype=np.uint32)


def _or8_table(chunks, out_bits=8):
    return _Tab(chunks, 2, 8, out_bits, [(x >> 8) | (x & 0xFF) for x in range(1 << 16)])


def test_lasso_tables_given_as_data_are_part_of_the_statement(okzg, vkzg):
    """b200v_lasso_verify_table: accepts the oracle's proof for the same descriptor; another subtable value, operand
    layout or output stride -> REJECT; a one-operand 16-bit table (x -> popcount) works the same way."""
    mu = 5
    for tab, two in ((_or8_table(3), True), (_Tab(2, 1, 16, 5, [bin(x).count("1") for x in range(1 << 16)]), False)):
        bits = tab.operand_bits * tab.chunks
        xs = O.rand_u64s(91, 1 << mu) & np.uint64((1 << bits) - 1)
        ys = (O.rand_u64s(92, 1 << mu) & np.uint64((1 << bits) - 1)) if two else None
        xs[1::2] = xs[0::2]
        if two:
            ys[1::2] = ys[0::2]
        otab = O.CustomTable(tab.chunks, tab.num_operands, tab.operand_bits, tab.out_bits, tab.values)
        tr = O.Transcript()
        assert O.lasso_prove_custom(okzg, tr, otab, mu, xs, ys) == 0
        proof = tr.proof()
        t = V.ProofTranscript(proof)
        assert vkzg.lasso_verify_table(t, tab, mu) and t.done()
        wrong_vals = tab.values.copy()
        wrong_vals[12345] += 1
        for bad in (_Tab(tab.chunks, tab.num_operands, tab.operand_bits, tab.out_bits, wrong_vals),
                    _Tab(tab.chunks, tab.num_operands, tab.operand_bits, tab.out_bits + 1, tab.values),
                    _Tab(tab.chunks, tab.num_operands, tab.operand_bits - 1, tab.out_bits, tab.values)):
            assert not vkzg.lasso_verify_table(V.ProofTranscript(proof), bad, mu)
        # an operand outside the table is refused by the prover (oracle rc 2)
        big = xs.copy()
        if bits < 64:
            big[3] = np.uint64(1 << bits)
            assert O.lasso_prove_custom(okzg, O.Transcript(), otab, mu, big, ys) == 2
    with pytest.raises(V.VerifierArgError):
        vkzg.lasso_verify_table(V.ProofTranscript(proof), _Tab(2, 2, 9, 8, tab.values), mu)  # 2 x 9 bits > 16


def _hyperplonk_verifier(okzg, vkzg, info, expr, nz):
    pre = [okzg.commit(O.fr_from_ints(p)) for p in info.preprocess_polys]
    sig = [okzg.commit(O.fr_from_ints(p)) for p in H.permutation_polys(info.k, info.permutation_polys, info.permutations)]
    return V.HyperPlonkVerifier(vkzg, info.k, info.num_instances, info.num_witness_polys, getattr(info, "num_challenges", None),
                                len(info.lookups), nz, expr, pre, sig)


def test_hyperplonk_verify_accepts_the_committed_golden_proofs(okzg, vkzg):
    gold = json.load(open(os.path.join(HERE, "golden", "hyperplonk_golden.json")))
    for c in gold["cases"]:
        k, proof = c["k"], bytes.fromhex(c["proof"])
        if c["circuit"] == "two_phase":
            info, inst_cols, _ = H.rand_two_phase_circuit(k, c["seed"], c["with_lookup"])
            nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys,
                               num_challenges=sum(info.num_challenges), lookups=info.lookups)
            instances = [v for col in inst_cols for v in col]
        else:
            fixture = H.rand_vanilla_plonk_with_lookup_circuit if c["circuit"].endswith("lookup") else H.rand_vanilla_plonk_circuit
            info, instances, _ = fixture(k, c["seed"], num_instances=2)
            nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys, max_degree=c["max_degree"], lookups=info.lookups)
        hv = _hyperplonk_verifier(okzg, vkzg, info, expr, nz)
        inst = O.fr_from_ints(instances)
        tr = V.ProofTranscript(proof)
        assert hv.verify(tr, inst) and tr.done(), c["circuit"]
        # the same with the expression composed inside the library (b200_expression_compose): C ABI on both sides
        import halo2_lasso_b200 as hl

        nz2, tok, cm = hl.compose_native(k, info.constraints, info.num_poly, info.permutation_polys,
                                         sum(getattr(info, "num_challenges", [0])), c.get("max_degree", 4), info.lookups)
        assert nz2 == nz
        assert _hyperplonk_verifier(okzg, vkzg, info, (tok, cm), nz).verify(V.ProofTranscript(proof), inst)
        for pos in (3, len(proof) // 2, len(proof) - 11):
            assert not hv.verify(V.ProofTranscript(tampered(proof, pos, 2)), inst)
        for j in range(inst.shape[0]):
            wrong = inst.copy()
            wrong[j] = O.rand_fr(99, 1)[0]
            assert not hv.verify(V.ProofTranscript(proof), wrong)
        assert not hv.verify(V.ProofTranscript(proof), inst[:-1])


@pytest.mark.parametrize("k,mu", [(4, 4), (10, 10)])
def test_cfg1_hyperplonk_then_lasso_on_one_transcript(okzg, vkzg, k, mu):
    """BASELINE cfg1 (HyperPlonk + Lasso 64-bit range check, 2^10 lookups into 2^16 subtables, bn256 MultilinearKzg, CPU
    plumbing) at toy size and at its stated size: the HyperPlonk section followed by the Lasso section on one proof
    stream, proved by the oracle, verified by the product's verifier."""
    chunks = 4
    info, instances, w = H.rand_vanilla_plonk_circuit(k, 77)
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys)
    ohp = O.HyperPlonk(okzg, k, expr, len(instances), 3, [O.fr_from_ints(p) for p in info.preprocess_polys],
                       info.permutation_polys, info.permutations, nz)
    inst = O.fr_from_ints(instances)
    xs = O.rand_u64s(9, 1 << mu)
    xs[1::2] = xs[0::2]
    to = O.Transcript()
    assert ohp.prove(to, inst, [O.fr_from_ints(c) for c in w])
    assert O.lasso_prove(okzg, to, O.TABLE_RANGE, chunks, mu, xs, None)
    hv = _hyperplonk_verifier(okzg, vkzg, info, expr, nz)
    tr = V.ProofTranscript(to.proof())
    assert hv.verify(tr, inst) and not tr.done()
    assert vkzg.lasso_verify(tr, O.TABLE_RANGE, chunks, mu) and tr.done()


def test_fractional_sum_check_verify_on_oracle_and_golden_proofs():
    """b200v_fractional_sum_check_verify (fractional_sum_check.rs:192-265): accepts the oracle prover's proofs and the
    committed golden proofs of the pure-Python model with the right claims and point, binds public claims (Some), rejects
    tampered bytes, truncated proofs and wrong public claims."""
    gold = json.load(open(os.path.join(HERE, "golden", "gkr_golden.json")))
    for c in gold["cases"]:
        B, n = c["batch"], c["num_vars"]
        proof = bytes.fromhex(c["proof"])
        p0 = O.fr_from_ints([int(v) for v in c["p_0s"]])
        q0 = O.fr_from_ints([int(v) for v in c["q_0s"]])
        cl_p = list(p0) if c["claimed"] else [None] * B
        cl_q = list(q0) if c["claimed"] else [None] * B
        tr = V.ProofTranscript(proof)
        res = V.fractional_sum_check_verify(tr, n, cl_p, cl_q)
        assert res is not None and tr.done()
        for got, key in zip(res, ("p_xs", "q_xs", "x", "p_0s", "q_0s")):
            assert O.fr_to_ints(got) == [int(v) for v in c[key]]
        if n > 1:
            assert V.fractional_sum_check_verify(V.ProofTranscript(tampered(proof, len(proof) // 2)), n, cl_p, cl_q) is None
        assert V.fractional_sum_check_verify(V.ProofTranscript(proof[:-32]), n, cl_p, cl_q) is None
        if c["claimed"]:
            wrong = list(p0)
            wrong[0] = O.rand_fr(5, 1)[0]
            assert V.fractional_sum_check_verify(V.ProofTranscript(proof), n, wrong, cl_q) is None
    # the reference's own test shape (batch of 3, nothing claimed) on a fresh oracle proof, mixed claims on another
    B, n = 3, 9
    ps = [O.rand_fr(7700 + b, 1 << n) for b in range(B)]
    qs = [O.rand_fr(7800 + b, 1 << n) for b in range(B)]
    to = O.Transcript()
    want = O.fractional_sum_check_prove(to, ps, qs, [0, None, None], [None, None, 0])
    res = V.fractional_sum_check_verify(V.ProofTranscript(to.proof()), n, [want[3][0], None, None], [None, None, want[4][2]])
    assert res is not None
    for b in range(B):
        assert (O.evaluate(ps[b], res[2]) == res[0][b]).all() and (O.evaluate(qs[b], res[2]) == res[1][b]).all()
    assert V.fractional_sum_check_verify(V.ProofTranscript(to.proof()), n, [None] * B, [None] * B) is None  # claims not bound
    with pytest.raises(V.VerifierArgError):
        V.fractional_sum_check_verify(V.ProofTranscript(to.proof()), 0, [None], [None])


def test_non_reduced_field_limbs_are_argument_errors(okzg, vkzg):
    """Fr inputs cross as Montgomery limbs; a residue >= r is malformed, not something to compute with"""
    bad = np.array([0xFFFFFFFFFFFFFFFF] * 4, dtype=np.uint64)
    modulus = np.array([(O.R_MOD >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    for limbs in (bad, modulus):
        with pytest.raises(V.VerifierArgError):
            V.fractional_sum_check_verify(V.ProofTranscript(b"\0" * 64), 1, [limbs], [None])
        with pytest.raises(V.VerifierArgError):
            V.MultilinearKzgVerifier.setup(np.stack([limbs] * 3))
        tr = V.ProofTranscript(b"")
        with pytest.raises(V.VerifierArgError):
            tr.common_field_elements(np.stack([limbs]))


def test_lasso_as_the_lookup_argument_of_a_hyperplonk_circuit(okzg, vkzg):
    """BASELINE cfg1 at its stated size, LINKED: a k = 10 vanilla-plonk circuit whose output wire holds 32-bit values in
    every row, the Lasso range check (c = 2 x 16 bit, 2^10 lookups = one per row, 2^16 subtables) over that whole column
    on the same transcript, and the verifier requires the Lasso section's commitment to `a` to BE the HyperPlonk
    section's commitment to w_o. A Lasso section about other values — even in-range ones — is rejected."""
    k, chunks, w_o = 10, 2, 2
    info, instances, w = H.range_checked_plonk_circuit(k, 4242)
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys)
    ohp = O.HyperPlonk(okzg, k, expr, len(instances), 3, [O.fr_from_ints(p) for p in info.preprocess_polys],
                       info.permutation_polys, info.permutations, nz)
    inst = O.fr_from_ints(instances)
    xs = np.asarray(w[w_o], dtype=np.uint64)

    def prove(lookups):
        tr = O.Transcript()
        assert ohp.prove(tr, inst, [O.fr_from_ints(c) for c in w])
        assert O.lasso_prove(okzg, tr, O.TABLE_RANGE, chunks, k, lookups, None)
        return tr.proof()

    hv = V.HyperPlonkLassoVerifier(_hyperplonk_verifier(okzg, vkzg, info, expr, nz), O.TABLE_RANGE, chunks, w_o)
    proof = prove(xs)
    assert (hv.witness_commitment(proof) == okzg.commit(O.fr_from_ints(w[w_o]))).all()
    assert hv.verify(proof, inst, k)
    other = xs.copy()
    other[5] ^= np.uint64(1)  # still a 32-bit value, but not the circuit's
    bad = prove(other)
    assert not hv.verify(bad, inst, k)
    # ... although each section of `bad` is valid on its own (the unlinked check accepts it)
    tr = V.ProofTranscript(bad)
    assert hv.hpv.verify(tr, inst) and vkzg.lasso_verify(tr, O.TABLE_RANGE, chunks, k) and tr.done()
    # an out-of-range witness cannot be proven at all: the range table with c = 2 has no entry for 2^32
    assert not hv.verify(proof[:-1], inst, k) and not hv.verify(tampered(proof, 64 * w_o + 5), inst, k)


def test_hyperplonk_verifier_rejects_malformed_parameters(okzg, vkzg):
    from halo2_lasso_b200.expression import Expression as E

    info, instances, _ = H.rand_vanilla_plonk_circuit(3, 5)
    nz, expr = compose(3, info.constraints, info.num_poly, info.permutation_polys)
    pre = [okzg.commit(O.fr_from_ints(p)) for p in info.preprocess_polys]
    sig = [okzg.commit(O.fr_from_ints(p)) for p in H.permutation_polys(3, info.permutation_polys, info.permutations)]
    V.HyperPlonkVerifier(vkzg, 3, info.num_instances, 3, None, 0, nz, expr, pre, sig)
    for bad_expr in (expr * E.polynomial(40), expr + E.challenge(3), expr * E.eq_xy(1), E.polynomial(0, 9) * expr, E.constant(5)):
        with pytest.raises(V.VerifierArgError):
            V.HyperPlonkVerifier(vkzg, 3, info.num_instances, 3, None, 0, nz, bad_expr, pre, sig)
    # a rotation of INT_MIN must not slip through an abs() (undefined there, and still negative): raw prefix tokens
    # PROD [EQXY 0] [POLY 0 rot]
    for rot in (-(2 ** 31), 2 ** 31 - 1, -17, 17):
        with pytest.raises(V.VerifierArgError):
            V.HyperPlonkVerifier(vkzg, 3, info.num_instances, 3, None, 0, nz,
                                 (np.array([8, 3, 0, 4, 0, rot], dtype=np.int64).astype(np.int32), np.zeros((0, 4), dtype=np.uint64)),
                                 pre, sig)
    with pytest.raises(V.VerifierArgError):
        V.HyperPlonkVerifier(vkzg, 3, info.num_instances, [2, 1], [0, 0], 0, nz, expr, pre, sig)  # phase 0 without challenges
    off_curve = [p.copy() for p in pre]
    off_curve[0][0] ^= 1
    with pytest.raises(V.VerifierArgError):
        V.HyperPlonkVerifier(vkzg, 3, info.num_instances, 3, None, 0, nz, expr, off_curve, sig)
    with pytest.raises(V.VerifierArgError):
        V.HyperPlonkVerifier(vkzg, 17, info.num_instances, 3, None, 0, nz, expr, pre, sig)  # k beyond the parameters


def test_plain_c_host_program_verifies_a_proof_file(okzg, tmp_path):
    """examples/verify_demo.c: gcc -std=c99 against include/b200_verify.h only — what a cgo / Rust-FFI binding links"""
    root = os.path.dirname(HERE)
    lib_dir = os.path.join(root, "halo2-lasso_b200")
    exe = str(tmp_path / "verify_demo")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                           os.path.join(root, "examples", "verify_demo.c"), "-L", lib_dir, "-lb200verify",
                           f"-Wl,-rpath,{lib_dir}", "-o", exe])
    mu = 6
    xs = O.rand_u64s(5, 1 << mu)
    xs[(1 << mu) // 2:] = xs[: (1 << mu) // 2]
    to = O.Transcript()
    assert O.lasso_prove(okzg, to, O.TABLE_RANGE, 4, mu, xs, None)
    (tmp_path / "proof.bin").write_bytes(to.proof())
    (tmp_path / "bad.bin").write_bytes(tampered(to.proof(), 500))
    (tmp_path / "ss.bin").write_bytes(np.ascontiguousarray(SS).tobytes())
    run = lambda name, *stmt: subprocess.run([exe, str(tmp_path / name), str(tmp_path / "ss.bin"), *map(str, stmt)],
                                             capture_output=True, text=True, timeout=120)
    ok = run("proof.bin", 0, 4, mu)
    assert ok.returncode == 0 and "accepted" in ok.stdout, ok.stdout + ok.stderr
    assert run("bad.bin", 0, 4, mu).returncode == 1
    assert run("proof.bin", 0, 4, mu + 1).returncode == 1
    assert run("proof.bin", 5, 4, mu).returncode == 2


@pytest.mark.parametrize("circuit", ["vanilla", "lookup", "two_phase"])
def test_verifier_parameters_derived_from_prover_parameters(okzg, vkzg, circuit):
    """`HyperPlonk.verifier()` (the vp half of preprocess) without a GPU: the prover-parameter object is stubbed with
    oracle commitments; the derived verifier accepts the oracle's proof of the same circuit."""
    k = 4
    if circuit == "two_phase":
        info, inst_cols, synth = H.rand_two_phase_circuit(k, 300)
        instances = [v for col in inst_cols for v in col]
    else:
        info, instances, w = (H.rand_vanilla_plonk_with_lookup_circuit if circuit == "lookup" else H.rand_vanilla_plonk_circuit)(k, 301)
    phases = info.num_witness_polys if isinstance(info.num_witness_polys, list) else [info.num_witness_polys]
    chals = list(getattr(info, "num_challenges", [0] * len(phases)))
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys, num_challenges=sum(chals), lookups=info.lookups)
    ohp = O.HyperPlonk(okzg, k, expr, info.num_instances, info.num_witness_polys, [O.fr_from_ints(p) for p in info.preprocess_polys],
                       info.permutation_polys, info.permutations, nz, lookups=info.lookups, num_challenges=chals)
    inst = O.fr_from_ints(instances)
    to = O.Transcript()
    if circuit == "two_phase":
        assert ohp.prove_phased(to, inst, lambda r, ch: [O.fr_from_ints(c) for c in synth(r, O.fr_to_ints(ch) if len(ch) else [])])
    else:
        assert ohp.prove(to, inst, [O.fr_from_ints(c) for c in w])
    stub = H.HyperPlonk.__new__(H.HyperPlonk)  # no context, no device: only what verifier() reads
    stub.info, stub.num_z = info, nz
    stub.instance_cols = list(info.num_instances) if isinstance(info.num_instances, list) else [info.num_instances]
    stub.phase_witness, stub.phase_challenges = phases, chals
    pre = np.stack([okzg.commit(O.fr_from_ints(p)) for p in info.preprocess_polys])
    sig = np.stack([okzg.commit(O.fr_from_ints(p)) for p in H.permutation_polys(k, info.permutation_polys, info.permutations)])
    stub.commitments = lambda: (pre, sig)
    hv = stub.verifier(vkzg)
    vt = V.ProofTranscript(to.proof())
    assert hv.verify(vt, inst) and vt.done()
    assert not hv.verify(V.ProofTranscript(tampered(to.proof(), 100)), inst)


def test_verifier_survives_garbage_and_truncated_proofs(okzg, vkzg):
    """Untrusted input: truncations at every kind of boundary, random byte flips and plain garbage are rejected (never
    accepted, never a crash) by the Lasso and HyperPlonk verifiers."""
    import random

    rng = random.Random(7)
    mu = 4
    xs = O.rand_u64s(3, 1 << mu)
    xs[1::2] = xs[0::2]
    to = O.Transcript()
    assert O.lasso_prove(okzg, to, O.TABLE_RANGE, 2, mu, xs, None)
    lasso = to.proof()
    gold = json.load(open(os.path.join(HERE, "golden", "hyperplonk_golden.json")))["cases"][1]
    info, instances, _ = H.rand_vanilla_plonk_with_lookup_circuit(gold["k"], gold["seed"], num_instances=2)
    nz, expr = compose(gold["k"], info.constraints, info.num_poly, info.permutation_polys, lookups=info.lookups)
    hv = _hyperplonk_verifier(okzg, vkzg, info, expr, nz)
    inst = O.fr_from_ints(instances)
    hp_proof = bytes.fromhex(gold["proof"])
    checks = [(lasso, lambda p: vkzg.lasso_verify(V.ProofTranscript(p), O.TABLE_RANGE, 2, mu)),
              (hp_proof, lambda p: hv.verify(V.ProofTranscript(p), inst))]
    for proof, verify in checks:
        assert verify(proof)
        cuts = {0, 1, 31, 32, 63, 64, 65, len(proof) - 1, len(proof) - 32, len(proof) // 2} | {rng.randrange(len(proof)) for _ in range(25)}
        for cut in cuts:
            assert not verify(proof[:cut])
        for _ in range(40):
            bad = bytearray(proof)
            for _ in range(rng.randrange(1, 4)):
                bad[rng.randrange(len(bad))] ^= 1 << rng.randrange(8)
            assert not verify(bytes(bad))
        for n in (0, 7, 64, len(proof), 2 * len(proof)):
            assert not verify(bytes(rng.randrange(256) for _ in range(n)))
        # trailing bytes: the section verifies, the end-of-proof check fails
        tr = V.ProofTranscript(proof + b"\\0" * 32)
        ok = vkzg.lasso_verify(tr, O.TABLE_RANGE, 2, mu) if proof is lasso else hv.verify(tr, inst)
        assert ok and not tr.done()


def test_pairing_selftest(tmp_path):
    """tests/host_shim/pairing_selftest.cpp over verifier/pairing.hpp: split final exponentiation == plain power by
    (p^12 - 1) / r, Fq12 inverse and p^2-Frobenius identities, bilinearity and non-degeneracy."""
    exe = str(tmp_path / "pairing_selftest")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(HERE, "host_shim", "pairing_selftest.cpp"), "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "0 failure(s)" in out.stdout, out.stdout + out.stderr
