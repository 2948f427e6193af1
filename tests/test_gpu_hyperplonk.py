"""GPU parity of the HyperPlonk prover (lookup-free vanilla plonk) vs the oracle: permutation grand product and
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
