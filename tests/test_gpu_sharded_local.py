"""Sharded provers with ALL ranks on one GPU (several contexts of this process, one host thread each,
b200_dist_init_local): the same kernels, mailboxes and bulk arenas as the one-process-per-GPU deployment, only the
peer pointers are plain device pointers instead of CUDA-IPC mappings over NVLink. Runs on a single-GPU box, so the
sharded paths are parity-checked wherever the GPU tests run; tests/test_gpu_sharded.py repeats them over NVLink.

Everything is compared byte-for-byte with the single-process CPU oracle."""
import os

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
os.environ.setdefault("B200_PEER_TIMEOUT_S", "10")  # a missing peer fails the test instead of hanging the GPU
os.environ.setdefault("B200_ARENA_MB", "64")
NV = 16


@pytest.fixture(scope="module")
def hl():
    import halo2_lasso_b200 as m

    return m


@pytest.fixture(scope="module")
def groups(hl):
    """world -> list of contexts forming a local group (created lazily, shared by the tests of the module)"""
    made = {}
    okzg = O.Kzg(O.rand_fr(7, NV))
    srs = [okzg.eqs(k) for k in range(NV + 1)]

    def get(world):
        if world not in made:
            ctxs = [hl.Context(0) for _ in range(world)]
            hl.dist_init_local(ctxs)
            kzgs = [hl.MultilinearKzg(c, srs) for c in ctxs]
            made[world] = (ctxs, kzgs)
        return made[world]

    yield get, okzg
    for ctxs, _ in made.values():
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("n,T,NP,p,R", [(10, 1, 2, -1, -1), (12, 1, 2, -1, 9), (11, 5, 2, 4, 3), (10, 4, 1, 3, 3),
                                        (12, 16, 2, 5, 0), (13, 3, 2, 6, 2)])
def test_windowed_sumcheck_parity(hl, groups, world, n, T, NP, p, R):
    """EVAL-shape sum-check on tables sharded on an index window: top-variable layout (p = -1), middle windows,
    R = 0 (pure all-gather), R < p (block-interleaved gather) and R = p; 33 tables exercise the multi-table gather."""
    get, _ = groups
    ctxs, _ = get(world)
    g = world.bit_length() - 1
    if p >= 0 and p > n - g:
        pytest.skip("window does not fit")
    if p < 0 and R > n - g:
        R = n - g
    one = O.fr_from_ints([1])[0]
    tabs = [O.rand_fr(7000 + n + i, 1 << n) for i in range(T * NP)]
    w = O.rand_fr(7100 + n, T) if T > 1 else one.reshape(1, 4)
    y = O.rand_fr(7200 + n, n)
    claim = O.rand_fr(7300 + n, 1)[0]
    to = O.Transcript()
    terms = [(w[t], list(range(t * NP, (t + 1) * NP))) for t in range(T)]
    ch_o, ev_o = O.sumcheck_prove_evals(to, n, tabs, y, terms, claim)
    pp = p if p >= 0 else n - g

    def run(rank, ctx):
        tr = hl.Keccak256Transcript(ctx)
        polys = [hl.MultilinearPolynomial.new(ctx, hl.shard_window_slice(t, n, pp, rank, world)) for t in tabs]
        ch, ev = hl.sumcheck_prove_evals_sharded(ctx, n, polys, w, y, claim, np_per_term=NP, window_pos=p, sharded_rounds=R)
        proof = tr.into_proof()
        hl.dist_check(ctx)
        return proof, ch, ev

    for rank, (proof, ch, ev) in enumerate(hl.run_ranks(ctxs, run)):
        assert proof == to.proof(), f"rank {rank}: sharded sum-check transcript differs"
        assert (ch == ch_o).all() and (ev == ev_o).all(), f"rank {rank}: outputs differ"


def lasso_operands(kind, chunks, mu, seed):
    xs, ys = O.rand_u64s(seed, 1 << mu), O.rand_u64s(seed + 1, 1 << mu)
    bits = (16 if kind == O.TABLE_RANGE else 8) * chunks
    if bits < 64:
        xs &= np.uint64((1 << bits) - 1)
        ys &= np.uint64((1 << bits) - 1)
    xs[1::4] = xs[0::4]
    ys[1::4] = ys[0::4]
    return xs, (None if kind == O.TABLE_RANGE else ys)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("kind,chunks,mu,k0,min_items", [(O.TABLE_RANGE, 4, 12, 8, 64), (O.TABLE_AND, 8, 11, 9, 16),
                                                        (O.TABLE_XOR, 2, 13, 10, 1 << 14), (O.TABLE_RANGE, 2, 14, 13, 256)])
def test_fully_sharded_lasso_proof_parity(hl, groups, world, kind, chunks, mu, k0, min_items):
    """b200_dist_shard_lasso: witness tables, fingerprints, tree layers >= k0, every sum-check (primary, grand-product
    layers, batch-open) and the quotient commitments on the rank's slice; small k0 also shards the 2^16 subtable trees.
    The proof of every rank equals the oracle's single-process proof."""
    get, okzg = groups
    ctxs, kzgs = get(world)
    g = world.bit_length() - 1
    if k0 - g < 1:
        pytest.skip("window does not fit")
    xs, ys = lasso_operands(kind, chunks, mu, 8400 + mu)
    to = O.Transcript()
    assert O.lasso_prove(okzg, to, kind, chunks, mu, xs, ys)

    def run(rank, ctx):
        hl.dist_shard_lasso(ctx, k0)
        hl.dist_shard_min_items(ctx, min_items)
        tr = hl.Keccak256Transcript(ctx)
        try:
            hl.LassoProver(ctx, kzgs[rank], kind, chunks).prove(xs, ys)
            proof = tr.into_proof()
            hl.dist_check(ctx)
        finally:
            hl.dist_shard_lasso(ctx, 0)
            hl.dist_shard_min_items(ctx, 1 << 16)
        return proof

    for rank, proof in enumerate(hl.run_ranks(ctxs, run)):
        assert proof == to.proof(), f"rank {rank}: sharded Lasso proof differs (kind {kind}, c {chunks}, mu {mu}, k0 {k0})"


@pytest.mark.parametrize("world", [2, 4])
def test_commit_and_sumcheck_sharding_of_the_replicated_prover(hl, groups, world):
    """the round-1 modes (replicated tables, point-sharded commitments, top-variable sum-check slices) still agree"""
    get, okzg = groups
    ctxs, kzgs = get(world)
    kind, chunks, mu = O.TABLE_AND, 4, 12
    xs, ys = lasso_operands(kind, chunks, mu, 8600)
    to = O.Transcript()
    assert O.lasso_prove(okzg, to, kind, chunks, mu, xs, ys)

    def run(rank, ctx):
        hl.dist_shard_commits(ctx, True)
        hl.dist_shard_sumchecks(ctx, 6)
        hl.dist_shard_min_items(ctx, 32)
        tr = hl.Keccak256Transcript(ctx)
        try:
            hl.LassoProver(ctx, kzgs[rank], kind, chunks).prove(xs, ys)
            proof = tr.into_proof()
            hl.dist_check(ctx)
        finally:
            hl.dist_shard_commits(ctx, False)
            hl.dist_shard_sumchecks(ctx, 0)
            hl.dist_shard_min_items(ctx, 1 << 16)
        return proof

    for rank, proof in enumerate(hl.run_ranks(ctxs, run)):
        assert proof == to.proof(), f"rank {rank}: proof differs"


def test_a_missing_peer_times_out_instead_of_hanging(hl):
    """bounded waits (peer.cuh): one rank of a 2-rank group calls a collective alone -> B200_ERR_PEER, no hang"""
    old = os.environ.get("B200_PEER_TIMEOUT_S")
    os.environ["B200_PEER_TIMEOUT_S"] = "0.5"
    try:
        ctxs = [hl.Context(0) for _ in range(2)]
        hl.dist_init_local(ctxs)
    finally:
        os.environ["B200_PEER_TIMEOUT_S"] = old or "10"
    n = 8
    one = O.fr_from_ints([1])[0]
    tabs = [O.rand_fr(1 + i, 1 << n) for i in range(2)]
    polys = [hl.MultilinearPolynomial.new(ctxs[0], hl.shard_window_slice(t, n, n - 1, 0, 2)) for t in tabs]
    hl.Keccak256Transcript(ctxs[0])
    hl.sumcheck_prove_evals_sharded(ctxs[0], n, polys, one.reshape(1, 4), O.rand_fr(3, n), one, sharded_rounds=2)
    with pytest.raises(hl.B200Error) as e:
        hl.dist_check(ctxs[0])
    assert e.value.code == hl.B200_ERR_PEER
    for c in ctxs:
        c.close()
