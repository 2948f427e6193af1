"""GPU parity: Pippenger MSM, MultilinearKzg commit / open / batch_open vs the CPU oracle."""
import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu

NV = 12


@pytest.fixture(scope="module")
def hl():
    import halo2_lasso_b200 as m

    return m


@pytest.fixture(scope="module")
def env(hl):
    ctx = hl.Context(0)
    okzg = O.Kzg(O.rand_fr(7, NV))
    levels = [okzg.eqs(k) for k in range(NV + 1)]
    kzg = hl.MultilinearKzg(ctx, levels)
    yield ctx, okzg, kzg, levels
    ctx.close()


@pytest.mark.parametrize("n", [1, 2, 7, 33, 300, 4096])
def test_msm_parity(hl, env, n):
    ctx, okzg, kzg, levels = env
    bases = levels[NV][:n]
    scalars = O.rand_fr(500 + n, n)
    assert (hl.variable_base_msm(ctx, scalars, bases) == O.msm(scalars, bases)).all()


def test_msm_edge_scalars(hl, env):
    """zeros, ones, -1, repeated scalars (one heavy bucket), and a sum that is the identity."""
    ctx, okzg, kzg, levels = env
    n = 2048
    bases = levels[NV][:n]
    sc = O.rand_fr(9, n)
    sc[:100] = 0
    sc[100:200] = O.fr_from_ints([1])[0]
    sc[200:300] = O.fr_from_ints([O.R_MOD - 1])[0]
    sc[300:1500] = O.fr_from_ints([12345])[0]  # heavy bucket -> task splitting + warp kernel
    assert (hl.variable_base_msm(ctx, sc, bases) == O.msm(sc, bases)).all()
    # P - P: same base twice with s and -s
    b2 = np.stack([bases[5], bases[5]])
    s2 = np.stack([O.fr_from_ints([77])[0], O.fr_from_ints([O.R_MOD - 77])[0]])
    assert (hl.variable_base_msm(ctx, s2, b2) == 0).all()
    # all-zero scalars -> identity
    assert (hl.variable_base_msm(ctx, np.zeros((64, 4), dtype=np.uint64), bases[:64]) == 0).all()
    # duplicate bases with equal scalars (bucket sees P + P -> doubling branch)
    b3 = np.stack([bases[9]] * 8)
    s3 = np.stack([O.fr_from_ints([3])[0]] * 8)
    assert (hl.variable_base_msm(ctx, s3, b3) == O.msm(s3, b3)).all()


def test_small_value_scalars(hl, env):
    """Lasso-style witness polynomials: tiny integers embedded with F::from(u64)."""
    ctx, okzg, kzg, levels = env
    n = 1 << NV
    vals = (O.rand_u64s(3, n) % np.uint64(23)).astype(np.uint64)
    sc = O.fr_from_ints([int(v) for v in vals])
    assert (hl.variable_base_msm(ctx, sc, levels[NV]) == O.msm(sc, levels[NV])).all()


@pytest.mark.parametrize("nv", [1, 3, 8, NV])
def test_commit_and_open_parity(hl, env, nv):
    ctx, okzg, kzg, levels = env
    poly = O.rand_fr(600 + nv, 1 << nv)
    point = O.rand_fr(700 + nv, nv)
    dp = hl.MultilinearPolynomial.new(ctx, poly)
    comm = kzg.commit(dp)
    assert (comm == okzg.commit(poly)).all()
    tr = hl.Keccak256Transcript(ctx)
    kzg.open(dp, point)
    to = O.Transcript()
    ev = okzg.open(to, poly, point)
    proof = tr.into_proof()
    assert proof == to.proof() and len(proof) == 64 * nv
    assert okzg.verify(O.Transcript(proof), comm, point, ev)


def test_batch_commit_and_write(hl, env):
    ctx, okzg, kzg, levels = env
    polys = [O.rand_fr(800 + i, 1 << nv) for i, nv in enumerate((NV, 5, NV, 9))]
    dps = [hl.MultilinearPolynomial.new(ctx, p) for p in polys]
    tr = hl.Keccak256Transcript(ctx)
    comms = kzg.batch_commit_and_write(dps)
    to = O.Transcript()
    for p in polys:
        to.write_comm(okzg.commit(p))
    assert tr.into_proof() == to.proof()
    assert all((comms[i] == okzg.commit(p)).all() for i, p in enumerate(polys))


def test_commit_zero_polynomial_is_transcript_error(hl, env):
    """reference: identity commitments cannot be written (transcript.rs:174-179)."""
    ctx, okzg, kzg, levels = env
    z = hl.MultilinearPolynomial.new(ctx, np.zeros((16, 4), dtype=np.uint64))
    hl.Keccak256Transcript(ctx)
    with pytest.raises(hl.B200Error) as e:
        kzg.batch_commit_and_write([z])
    assert e.value.code == hl.B200_ERR_TRANSCRIPT
    hl.Keccak256Transcript(ctx)


def test_batch_open_parity_and_verifies(hl, env):
    """reference harness shape (pb/pcs/multilinear.rs:340-406): several polys, several points."""
    ctx, okzg, kzg, levels = env
    nv = 10
    polys = [O.rand_fr(900 + i, 1 << nv) for i in range(5)]
    points = [O.rand_fr(950 + i, nv) for i in range(3)]
    pairs = [(0, 0), (1, 1), (2, 1), (3, 2), (4, 2), (0, 2), (1, 2)]
    evals = [(p, q, O.evaluate(polys[p], points[q])) for p, q in pairs]
    dps = [hl.MultilinearPolynomial.new(ctx, p) for p in polys]
    tr = hl.Keccak256Transcript(ctx)
    seed = O.rand_fr(999, 2)
    tr.common_field_elements(seed)
    kzg.batch_open(dps, points, evals)
    to = O.Transcript()
    for f in seed:
        to.common_fe(f)
    okzg.batch_open(to, polys, points, evals)
    proof = tr.into_proof()
    assert proof == to.proof()
    tv = O.Transcript(proof)
    for f in seed:
        tv.common_fe(f)
    comms = [okzg.commit(p) for p in polys]
    assert okzg.batch_verify(tv, comms, points, evals)
