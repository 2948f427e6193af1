"""GPU parity of the HyperPlonk prover (vanilla plonk, without and with LogUp lookups) vs the oracle: permutation grand product, lookup polynomials and
whole proofs byte-for-byte; the oracle verifier accepts the GPU proofs."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
NV = 10


@pytest.fixture(scope="module")
def hl():
    import halo2_lasso_b200 as m

    return m


@pytest.fixture(scope="module")
def env(hl):
    ctx = hl.Context(0)
    okzg = O.Kzg(O.rand_fr(7, NV))
    kzg = hl.MultilinearKzg(ctx, [okzg.eqs(k) for k in range(NV + 1)])
    yield ctx, okzg, kzg
    ctx.close()


@pytest.mark.parametrize("k", [1, 4, 9, 13])
def test_permutation_z_parity(hl, env, k):
    import ctypes as C

    ctx, okzg, kzg = env
    wires = [O.rand_fr(100 + k + i, 1 << k) for i in range(3)]
    sigmas = [O.rand_fr(200 + k + i, 1 << k) for i in range(3)]
    bg = O.rand_fr(300 + k, 2)
    exp = O.permutation_z(sigmas, wires, bg[0], bg[1])
    dw = [hl.MultilinearPolynomial.new(ctx, t) for t in wires]
    ds = [hl.MultilinearPolynomial.new(ctx, t) for t in sigmas]
    z = hl.MultilinearPolynomial.alloc(ctx, k)
    wp = (C.c_void_p * 3)(*[p.dev for p in dw])
    sp = (C.c_void_p * 3)(*[p.dev for p in ds])
    offs = (C.c_uint64 * 3)(*[i << k for i in range(3)])
    hl._chk(hl.lib().b200_permutation_z(ctx.h, C.c_int(k), C.c_int(3), wp, sp, offs, hl._p(np.ascontiguousarray(bg)), z.dev), "pz")
    assert (z.evals() == exp).all()


@pytest.mark.parametrize("k", [3, 5, 9])
def test_hyperplonk_proof_parity_and_verifies(hl, env, k):
    from halo2_lasso_b200 import hyperplonk as H
    from halo2_lasso_b200.expression import compose

    ctx, okzg, kzg = env
    info, instances, w = H.rand_vanilla_plonk_circuit(k, 40 + k)
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys)
    ohp = O.HyperPlonk(okzg, k, expr, len(instances), 3, [O.fr_from_ints(p) for p in info.preprocess_polys],
                       info.permutation_polys, info.permutations, nz)
    inst = O.fr_from_ints(instances)
    to = O.Transcript()
    assert ohp.prove(to, inst, [O.fr_from_ints(c) for c in w])
    hp = H.HyperPlonk(ctx, kzg, info)
    # preprocess: permutation polynomials and the verifier's commitments (hyperplonk.rs:127-150)
    assert (hp.num_z, hp.degree, hp.num_polys) == (1, 5, 13)
    for i in range(3):
        assert (hp.permutation_poly(i) == ohp.permutation_poly(i)).all()
    pre_c, perm_c = hp.commitments()
    assert (pre_c == np.stack([okzg.commit(O.fr_from_ints(p)) for p in info.preprocess_polys])).all()
    assert (perm_c == np.stack([okzg.commit(ohp.permutation_poly(i)) for i in range(3)])).all()
    tr = hl.Keccak256Transcript(ctx)
    hp.prove(instances, witness_ints=w)
    proof = tr.into_proof()
    ref = to.proof()
    if proof != ref:
        first = next(i for i in range(min(len(proof), len(ref))) if proof[i] != ref[i])
        pytest.fail(f"HyperPlonk proof differs from the oracle at byte {first} (lengths {len(proof)} vs {len(ref)})")
    assert ohp.verify(O.Transcript(proof), inst)


def test_cfg1_hyperplonk_plus_lasso_on_one_transcript(hl):
    """BASELINE cfg1 shape: a HyperPlonk proof (k = 10 vanilla plonk) followed by a Lasso range-check proof for
    2^10 lookups into 2^16 subtables on the SAME Fiat-Shamir transcript (the Lasso section sits behind the
    HyperPlonk section); byte-identical to the oracle running the same composition, and both
    oracle verifiers accept when replayed in order."""
    from halo2_lasso_b200 import hyperplonk as H
    from halo2_lasso_b200.expression import compose

    k, mu, chunks = 10, 10, 4
    ctx = hl.Context(0)
    ss = O.rand_fr(7, 16)
    okzg = O.Kzg(ss)
    kzg = hl.MultilinearKzg(ctx, [okzg.eqs(i) for i in range(17)])
    info, instances, w = H.rand_vanilla_plonk_circuit(k, 77)
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys)
    ohp = O.HyperPlonk(okzg, k, expr, len(instances), 3, [O.fr_from_ints(p) for p in info.preprocess_polys],
                       info.permutation_polys, info.permutations, nz)
    xs = O.rand_u64s(9, 1 << mu)
    xs[1::2] = xs[0::2]
    inst = O.fr_from_ints(instances)
    to = O.Transcript()
    assert ohp.prove(to, inst, [O.fr_from_ints(c) for c in w])
    assert O.lasso_prove(okzg, to, O.TABLE_RANGE, chunks, mu, xs, None)
    hp = H.HyperPlonk(ctx, kzg, info)
    tr = hl.Keccak256Transcript(ctx)
    hp.prove(instances, witness_ints=w)
    hl.LassoProver(ctx, kzg, O.TABLE_RANGE, chunks).prove(xs)
    assert tr.into_proof() == to.proof()
    ctx.close()


def test_cfg1_lasso_as_the_lookup_argument_linked_by_commitment(hl):
    """BASELINE cfg1, linked: k = 10 circuit whose output wire is 32-bit in every row + the Lasso range check over that
    whole column (2^10 lookups into 2^16 subtables) proved on the GPU on one transcript (`HyperPlonkLasso`); bytes equal
    the oracle composition, and the product's CPU verifier accepts only because the Lasso section's commitment to `a` IS
    the HyperPlonk section's commitment to w_o (`HyperPlonkLassoVerifier`)."""
    from halo2_lasso_b200 import hyperplonk as H
    from halo2_lasso_b200 import verifier as V
    from halo2_lasso_b200.expression import compose

    k, chunks, w_o = 10, 2, 2
    ctx = hl.Context(0)
    ss = O.rand_fr(7, 16)
    okzg = O.Kzg(ss)
    kzg = hl.MultilinearKzg(ctx, [okzg.eqs(i) for i in range(17)])
    info, instances, w = H.range_checked_plonk_circuit(k, 4242)
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys)
    ohp = O.HyperPlonk(okzg, k, expr, len(instances), 3, [O.fr_from_ints(p) for p in info.preprocess_polys],
                       info.permutation_polys, info.permutations, nz)
    inst = O.fr_from_ints(instances)
    to = O.Transcript()
    assert ohp.prove(to, inst, [O.fr_from_ints(c) for c in w])
    assert O.lasso_prove(okzg, to, O.TABLE_RANGE, chunks, k, np.asarray(w[w_o], dtype=np.uint64), None)
    tr = hl.Keccak256Transcript(ctx)
    H.HyperPlonkLasso(ctx, kzg, info, O.TABLE_RANGE, chunks, w_o).prove(instances, w)
    proof = tr.into_proof()
    assert proof == to.proof()
    vk = V.MultilinearKzgVerifier.setup(ss)
    pre = [okzg.commit(O.fr_from_ints(p)) for p in info.preprocess_polys]
    sig = [okzg.commit(O.fr_from_ints(p)) for p in H.permutation_polys(k, info.permutation_polys, info.permutations)]
    hpv = V.HyperPlonkVerifier(vk, k, info.num_instances, 3, None, 0, nz, expr, pre, sig)
    assert V.HyperPlonkLassoVerifier(hpv, O.TABLE_RANGE, chunks, w_o).verify(proof, inst, k)
    assert not V.HyperPlonkLassoVerifier(hpv, O.TABLE_RANGE, chunks, 1).verify(proof, inst, k)  # linked to the wrong column
    ctx.close()


# ---- LogUp lookup argument (prover.rs:50-250) --------------------------------------------------------------
@pytest.mark.parametrize("k", [2, 5, 11])
def test_expression_rows_parity(hl, env, k):
    from halo2_lasso_b200.expression import Expression as E

    ctx, okzg, kzg = env
    N = 1 << k
    polys = [O.rand_fr(30 + k + i, N) for i in range(4)]
    ch = [O.fr_to_ints(O.rand_fr(50 + k, 2))[i] for i in range(2)]
    expr = (E.polynomial(0) * E.polynomial(1, 1) + E.challenge(1) * E.polynomial(2, -1) - E.identity() * E.lagrange(2)
            + E.constant(9) + E.distribute_powers([E.polynomial(3), E.polynomial(0) * E.polynomial(3), E.polynomial(2)], E.challenge(0)))
    want = O.expression_rows(k, expr, polys, O.fr_from_ints(ch))
    dev = [hl.MultilinearPolynomial.new(ctx, p) for p in polys]
    got = hl.expression_rows(ctx, k, expr, dev, ch).evals()
    assert (got == want).all()


@pytest.mark.parametrize("k", [1, 6, 12])
def test_lookup_m_and_h_parity(hl, env, k):
    ctx, okzg, kzg = env
    N = 1 << k
    table = O.rand_fr(3 + k, N)
    table[0] = table[1] = 0
    if N > 32:
        table[7] = table[20]
        table[N - 1] = table[20]
    idx = np.asarray([int(x) % N for x in O.rand_u64s(9 + k, N)])
    idx[: N // 2] = 1 if N > 2 else 0  # half of the rows gated off -> the zero tuple (heavy contention on one row)
    if N > 32:
        idx[-3:] = [7, 20, N - 1]
    inp = table[idx]
    want_m = O.lookup_m(inp, table)
    d_in, d_tab = hl.MultilinearPolynomial.new(ctx, inp), hl.MultilinearPolynomial.new(ctx, table)
    m = hl.lookup_m_poly(ctx, k, d_in, d_tab)
    assert (m.evals() == want_m).all()
    gamma = O.rand_fr(4, 1)[0]
    h = hl.lookup_h_poly(ctx, k, d_in, d_tab, m, gamma)
    assert (h.evals() == O.lookup_h(inp, table, want_m, gamma)).all()
    bad = inp.copy()
    bad[N - 1] = O.rand_fr(77, 1)[0]
    with pytest.raises(hl.B200Error) as e:
        hl.lookup_m_poly(ctx, k, hl.MultilinearPolynomial.new(ctx, bad), d_tab)
    assert e.value.code == hl.B200_ERR_LOOKUP


@pytest.mark.parametrize("k", [3, 5, 9])
def test_hyperplonk_with_lookup_proof_parity_and_verifies(hl, env, k):
    """vanilla_plonk_with_lookup (util.rs:63-98, 216-316): 19 polynomials, degree-5 zero check, LogUp m / h polys."""
    from halo2_lasso_b200 import hyperplonk as H
    from halo2_lasso_b200.expression import compose

    ctx, okzg, kzg = env
    info, instances, w = H.rand_vanilla_plonk_with_lookup_circuit(k, 70 + k)
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys, lookups=info.lookups)
    ohp = O.HyperPlonk(okzg, k, expr, len(instances), 3, [O.fr_from_ints(p) for p in info.preprocess_polys],
                       info.permutation_polys, info.permutations, nz, lookups=info.lookups)
    inst = O.fr_from_ints(instances)
    to = O.Transcript()
    assert ohp.prove(to, inst, [O.fr_from_ints(c) for c in w])
    hp = H.HyperPlonk(ctx, kzg, info)
    tr = hl.Keccak256Transcript(ctx)
    hp.prove(instances, witness_ints=w)
    proof = tr.into_proof()
    ref = to.proof()
    if proof != ref:
        first = next(i for i in range(min(len(proof), len(ref))) if proof[i] != ref[i])
        pytest.fail(f"HyperPlonk+lookup proof differs from the oracle at byte {first} (lengths {len(proof)} vs {len(ref)})")
    assert ohp.verify(O.Transcript(proof), inst)
    # a lookup row outside the table: Error::InvalidSnark("Invalid lookup input")
    row = next(b for b in range(1 << k) if info.preprocess_polys[5][b])
    w_bad = [list(c) for c in w]
    w_bad[1][row] = (w_bad[1][row] + 1) % H.R_MOD
    hl.Keccak256Transcript(ctx)
    with pytest.raises(hl.B200Error) as e:
        hp.prove(instances, witness_ints=w_bad)
    assert e.value.code == hl.B200_ERR_LOOKUP


@pytest.mark.parametrize("k", [4, 8])
def test_hyperplonk_two_permutation_chunks(hl, env, k):
    """max_degree = 3 splits the three permuted wire columns into chunks of two (preprocessor.rs:111-170): two z
    polynomials whose running product interleaves (row, chunk) (prover.rs:308-344)."""
    from halo2_lasso_b200 import hyperplonk as H
    from halo2_lasso_b200.expression import compose

    ctx, okzg, kzg = env
    info, instances, w = H.rand_vanilla_plonk_circuit(k, 90 + k)
    info.max_degree = 3
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys, max_degree=3)
    assert nz == 2 and expr.degree() == 4
    ohp = O.HyperPlonk(okzg, k, expr, len(instances), 3, [O.fr_from_ints(p) for p in info.preprocess_polys],
                       info.permutation_polys, info.permutations, nz)
    inst = O.fr_from_ints(instances)
    to = O.Transcript()
    assert ohp.prove(to, inst, [O.fr_from_ints(c) for c in w])
    hp = H.HyperPlonk(ctx, kzg, info)
    assert (hp.num_z, hp.degree, hp.num_polys) == (2, 4, 14)
    tr = hl.Keccak256Transcript(ctx)
    hp.prove(instances, witness_ints=w)
    proof = tr.into_proof()
    assert proof == to.proof()
    assert ohp.verify(O.Transcript(proof), inst)


@pytest.mark.parametrize("k,with_lookup", [(4, True), (6, False), (9, True)])
def test_hyperplonk_two_phases_two_instance_columns(hl, env, k, with_lookup):
    """PlonkishCircuitInfo with two instance columns (one queried at Rotation::next) and two witness phases
    (pb/backend.rs:50-60, hyperplonk.rs:183-204): the phase-1 witness is synthesized by a host callback from the
    challenges squeezed after the phase-0 commitments; a circuit challenge also sits inside the lookup input. The
    proof is byte-identical to the oracle's (which is pinned by the pure-Python model) and the oracle verifies it."""
    from halo2_lasso_b200 import hyperplonk as H
    from halo2_lasso_b200.expression import compose

    ctx, okzg, kzg = env
    info, inst_cols, synth = H.rand_two_phase_circuit(k, 120 + k, with_lookup)
    nz, expr = compose(k, info.constraints, info.num_poly, info.permutation_polys,
                       num_challenges=sum(info.num_challenges), lookups=info.lookups)
    ohp = O.HyperPlonk(okzg, k, expr, info.num_instances, info.num_witness_polys, [O.fr_from_ints(p) for p in info.preprocess_polys],
                       info.permutation_polys, info.permutations, nz, lookups=info.lookups, num_challenges=info.num_challenges)
    inst = O.fr_from_ints([v for col in inst_cols for v in col])
    seen_o, seen_g = [], []

    def synth_o(rnd, ch):
        ints = O.fr_to_ints(ch) if len(ch) else []
        seen_o.append((rnd, ints))
        return [O.fr_from_ints(c) for c in synth(rnd, ints)]

    def synth_g(rnd, ch):
        seen_g.append((rnd, list(ch)))
        return synth(rnd, ch)

    to = O.Transcript()
    assert ohp.prove_phased(to, inst, synth_o)
    hp = H.HyperPlonk(ctx, kzg, info)
    assert (hp.num_z, hp.degree, hp.num_polys) == (1, 5, 12 + 3 + 2 * len(info.lookups) + 1)
    tr = hl.Keccak256Transcript(ctx)
    hp.prove_phased(inst_cols, synth_g)
    proof = tr.into_proof()
    assert seen_g == seen_o, "the synthesize callback must see the same challenges as the oracle's"
    ref = to.proof()
    if proof != ref:
        first = next(i for i in range(min(len(proof), len(ref))) if proof[i] != ref[i])
        pytest.fail(f"two-phase HyperPlonk proof differs from the oracle at byte {first} (lengths {len(proof)} vs {len(ref)})")
    assert ohp.verify(O.Transcript(proof), inst)
    # error behaviour: wrong number of instances, a failing synthesize callback, the single-phase entry point
    hl.Keccak256Transcript(ctx)
    with pytest.raises(hl.B200Error) as e:
        hp.prove_phased([inst_cols[0], inst_cols[1][:-1]], synth_g)
    assert e.value.code == hl.B200_ERR_ARG
    hl.Keccak256Transcript(ctx)
    with pytest.raises(hl.B200Error) as e:
        hp.prove_phased(inst_cols, lambda rnd, ch: synth(rnd, ch)[:1])
    assert e.value.code == hl.B200_ERR_ARG
    hl.Keccak256Transcript(ctx)
    with pytest.raises(hl.B200Error) as e:
        hp.prove([v for col in inst_cols for v in col], witness_ints=synth(0, []))
    assert e.value.code == hl.B200_ERR_ARG


def test_hyperplonk_preprocess_phased_rejects_malformed_phases(hl, env):
    """is_well_formed (pb/backend.rs:76-105): a phase without witness polynomials, a non-final phase without
    challenges and an out-of-range challenge index are rejected with B200_ERR_ARG."""
    from halo2_lasso_b200 import hyperplonk as H
    from halo2_lasso_b200.expression import Expression as E

    ctx, okzg, kzg = env
    info, inst_cols, synth = H.rand_two_phase_circuit(4, 5, with_lookup=False)
    for attr, value in (("num_witness_polys", [4, 0]), ("num_challenges", [0, 0])):
        bad, _, _ = H.rand_two_phase_circuit(4, 5, with_lookup=False)
        setattr(bad, attr, value)
        with pytest.raises(hl.B200Error) as e:
            H.HyperPlonk(ctx, kzg, bad)
        assert e.value.code == hl.B200_ERR_ARG
    bad, _, _ = H.rand_two_phase_circuit(4, 5, with_lookup=False)
    bad.constraints = bad.constraints + [E.polynomial(8) * E.challenge(2)]  # only challenges 0 and 1 exist
    with pytest.raises(hl.B200Error) as e:
        H.HyperPlonk(ctx, kzg, bad)
    assert e.value.code == hl.B200_ERR_ARG
    H.HyperPlonk(ctx, kzg, info)


def test_hyperplonk_matches_committed_golden_bytes(hl, env):
    """The GPU prover against tests/golden/hyperplonk_golden.json directly (proof bytes produced by the pure-Python model,
    committed): vanilla plonk with one and two permutation chunks, with the LogUp lookup, and the two-phase circuit.
    The SRS of `env` uses the same seeded trapdoor stream (seed 7) as the fixtures."""
    import json
    import os

    from halo2_lasso_b200 import hyperplonk as H

    ctx, okzg, kzg = env
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hyperplonk_golden.json")))
    for c in gold["cases"]:
        assert c["srs_seed"] == 7
        k, want = c["k"], bytes.fromhex(c["proof"])
        tr = hl.Keccak256Transcript(ctx)
        if c["circuit"] == "two_phase":
            info, inst_cols, synth = H.rand_two_phase_circuit(k, c["seed"], c["with_lookup"])
            H.HyperPlonk(ctx, kzg, info).prove_phased(inst_cols, synth)
        else:
            fixture = H.rand_vanilla_plonk_with_lookup_circuit if c["circuit"].endswith("lookup") else H.rand_vanilla_plonk_circuit
            info, instances, w = fixture(k, c["seed"], num_instances=2)
            info.max_degree = c["max_degree"]
            H.HyperPlonk(ctx, kzg, info).prove(instances, witness_ints=w)
        assert tr.into_proof() == want, c


@pytest.mark.parametrize("circuit", ["vanilla", "lookup", "two_phase"])
def test_preprocess_prove_on_the_gpu_verify_with_the_products_cpu_verifier(hl, env, circuit):
    """`preprocess -> (pp, vp)`, `prove(pp)`, `verify(vp)` (pb/backend.rs:202-241 run_plonkish_backend) entirely inside
    the product: prover parameters and proof on the GPU, verifier parameters derived from them (commitments + the
    library-composed expression), verification by libb200verify.so — the oracle only supplies the SRS trapdoor stream."""
    from halo2_lasso_b200 import hyperplonk as H
    from halo2_lasso_b200 import verifier as V

    ctx, okzg, kzg = env
    k = 6
    vk = V.MultilinearKzgVerifier.setup(O.rand_fr(7, NV))
    tr = hl.Keccak256Transcript(ctx)
    if circuit == "two_phase":
        info, inst_cols, synth = H.rand_two_phase_circuit(k, 300)
        hp = H.HyperPlonk(ctx, kzg, info)
        hp.prove_phased(inst_cols, synth)
        instances = [v for col in inst_cols for v in col]
    else:
        fixture = H.rand_vanilla_plonk_with_lookup_circuit if circuit == "lookup" else H.rand_vanilla_plonk_circuit
        info, instances, w = fixture(k, 301)
        hp = H.HyperPlonk(ctx, kzg, info)
        hp.prove(instances, witness_ints=w)
    proof = tr.into_proof()
    hv = hp.verifier(vk)
    inst = O.fr_from_ints(instances)
    vt = V.ProofTranscript(proof)
    assert hv.verify(vt, inst) and vt.done()
    bad = bytearray(proof)
    bad[len(bad) // 2] ^= 1
    assert not hv.verify(V.ProofTranscript(bytes(bad)), inst)
    wrong = inst.copy()
    wrong[0] = O.rand_fr(5, 1)[0]
    assert not hv.verify(V.ProofTranscript(proof), wrong)
