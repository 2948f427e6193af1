"""Multi-GPU worker (launched by torchrun, one rank per GPU): hypercube-sharded sum-check and point-sharded
MSM through the C ABI, checked byte-for-byte against the single-process CPU oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import numpy as np
import torch
import torch.distributed as dist

import halo2_lasso_b200 as hl
import oracle as O


def check_sharded_sumchecks_in_lasso(ctx, kzg, okzg, rank, world, cases):
    """One Lasso proof on all GPUs with BOTH the commitment MSMs point-sharded and the prover's sum-checks
    hypercube-sharded (b200_dist_shard_sumchecks): byte-identical to the single-process oracle proof on every rank.
    min_vars = 6 shards nearly every grand-product layer, including the pre-bound-eq round kernels (pairs >= 2048,
    T >= 4) and, with 8 chunks, the 33-value final gather that needs two mailbox messages."""
    hl.dist_shard_commits(ctx, True)
    hl.dist_shard_sumchecks(ctx, 6)
    hl.dist_shard_min_items(ctx, 32)
    for kind, chunks, mu in cases:
        xs, ys = O.rand_u64s(8400 + mu, 1 << mu), O.rand_u64s(8500 + mu, 1 << mu)
        if kind == O.TABLE_RANGE:
            ys = None
        else:
            xs &= np.uint64((1 << (8 * chunks)) - 1)
            ys &= np.uint64((1 << (8 * chunks)) - 1)
        to = O.Transcript()
        assert O.lasso_prove(okzg, to, kind, chunks, mu, xs, ys)
        tr = hl.Keccak256Transcript(ctx)
        hl.LassoProver(ctx, kzg, kind, chunks).prove(xs, ys)
        assert tr.into_proof() == to.proof(), f"rank {rank}: sum-check-sharded Lasso proof differs (kind {kind}, c {chunks}, mu {mu})"
    hl.dist_check(ctx)
    hl.dist_shard_min_items(ctx, 1 << 16)
    hl.dist_shard_sumchecks(ctx, 0)
    hl.dist_shard_commits(ctx, False)


def check_fully_sharded_lasso(ctx, kzg, okzg, rank, world, cases):
    """b200_dist_shard_lasso over real CUDA-IPC / NVLink peers: tables, trees, sum-checks and quotient commitments on the
    rank's index-window slice; byte-identical to the single-process oracle proof on every rank."""
    g = world.bit_length() - 1
    for kind, chunks, mu, k0, min_items in cases:
        if k0 - g < 1:
            continue
        xs, ys = O.rand_u64s(8700 + mu, 1 << mu), O.rand_u64s(8800 + mu, 1 << mu)
        if kind == O.TABLE_RANGE:
            ys = None
            if chunks < 4:
                xs &= np.uint64((1 << (16 * chunks)) - 1)
        else:
            xs &= np.uint64((1 << (8 * chunks)) - 1)
            ys &= np.uint64((1 << (8 * chunks)) - 1)
        to = O.Transcript()
        assert O.lasso_prove(okzg, to, kind, chunks, mu, xs, ys)
        hl.dist_shard_lasso(ctx, k0)
        hl.dist_shard_min_items(ctx, min_items)
        tr = hl.Keccak256Transcript(ctx)
        hl.LassoProver(ctx, kzg, kind, chunks).prove(xs, ys)
        proof = tr.into_proof()
        hl.dist_check(ctx)
        hl.dist_shard_lasso(ctx, 0)
        hl.dist_shard_min_items(ctx, 1 << 16)
        assert proof == to.proof(), f"rank {rank}: fully sharded Lasso proof differs (kind {kind}, c {chunks}, mu {mu}, k0 {k0})"


def time_cooperative_lasso(ctx, kzg, rank, world, mu=20):
    """ms per 2^mu-lookup range proof on all GPUs: replicated / commitments sharded / commitments + sum-checks sharded"""
    prover = hl.LassoProver(ctx, kzg, O.TABLE_RANGE, 4)
    xs = torch.from_numpy(O.rand_u64s(5, 1 << mu).view(np.int64)).cuda()
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.cuda.current_device())
    out = {}
    proofs = []
    for name, commits, min_vars in (("replicated", False, 0), ("commits", True, 0), ("commits+sumchecks", True, 14)):
        hl.dist_shard_commits(ctx, commits)
        hl.dist_shard_sumchecks(ctx, min_vars)
        tr = hl.Keccak256Transcript(ctx)
        prover.prove_dev(mu, xs.data_ptr())
        proofs.append(tr.into_proof())
        for _ in range(2):
            hl.Keccak256Transcript(ctx)
            prover.prove_dev(mu, xs.data_ptr())
        ctx.sync()
        dist.barrier()
        steps = 5
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            hl.Keccak256Transcript(ctx)
            a.record(stream)
            prover.prove_dev(mu, xs.data_ptr())
            b.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / steps], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = round(float(t.item()), 3)
    hl.dist_shard_sumchecks(ctx, 0)
    hl.dist_shard_commits(ctx, False)
    assert proofs[0] == proofs[1] == proofs[2], f"rank {rank}: the 2^{mu} proofs of the three modes differ"
    if rank == 0:
        print(f"COOPERATIVE_LASSO world={world} mu={mu} ms={out} proof_bytes={len(proofs[0])}")


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl" if os.environ.get("DIST_BACKEND", "nccl") == "nccl" else "gloo",
                            device_id=torch.device("cuda", local))
    ctx = hl.Context(local)
    hl.dist_init(ctx, rank, world)
    if os.environ.get("DIST_QUICK"):  # only the sum-check-sharded whole prover (+ timing): a short GPU session
        timing = os.environ.get("DIST_QUICK") == "time"
        okzg = O.Kzg(O.rand_fr(7, 16))
        kzg = hl.MultilinearKzg.setup(ctx, O.rand_fr(7, 20 if timing else 16))  # same trapdoor prefix as the oracle's
        check_sharded_sumchecks_in_lasso(ctx, kzg, okzg, rank, world,
                                         ((O.TABLE_RANGE, 4, 15), (O.TABLE_AND, 8, 12), (O.TABLE_XOR, 2, 13)))
        if rank == 0:
            print(f"SHARDED_SUMCHECKS_OK world={world}", flush=True)
        if timing:
            time_cooperative_lasso(ctx, kzg, rank, world)
        dist.barrier()
        ctx.close()
        dist.destroy_process_group()
        return
    one = O.fr_from_ints([1])[0]
    for n, T, NP in ((6, 1, 2), (12, 1, 2), (11, 5, 2), (10, 4, 1)):
        tabs = [O.rand_fr(7000 + n + i, 1 << n) for i in range(T * NP)]
        w = O.rand_fr(7100 + n, T) if T > 1 else one.reshape(1, 4)
        y = O.rand_fr(7200 + n, n)
        claim = O.rand_fr(7300 + n, 1)[0]
        to = O.Transcript()
        terms = [(w[t], list(range(t * NP, (t + 1) * NP))) for t in range(T)]
        ch_o, ev_o = O.sumcheck_prove_evals(to, n, tabs, y, terms, claim)
        lo, hi = hl.shard_slice(n, rank, world)
        tr = hl.Keccak256Transcript(ctx)
        polys = [hl.MultilinearPolynomial.new(ctx, t[lo:hi]) for t in tabs]
        g = world.bit_length() - 1
        ch, ev = hl.sumcheck_prove_evals_sharded(ctx, n, polys, w, y, claim, np_per_term=NP, sharded_rounds=max(0, n - g - (n % 3)))
        proof = tr.into_proof()
        assert proof == to.proof(), f"rank {rank}: sharded sum-check transcript differs (n={n}, T={T}, NP={NP})"
        assert (ch == ch_o).all() and (ev == ev_o).all(), f"rank {rank}: outputs differ"
    # point-sharded MSM
    kz = O.Kzg(O.rand_fr(7, 10))
    bases = kz.eqs(10)
    sc = O.rand_fr(99, 1 << 10)
    lo, hi = hl.shard_slice(10, rank, world)
    got = hl.variable_base_msm_sharded(ctx, sc[lo:hi], bases[lo:hi])
    assert (got == O.msm(sc, bases)).all(), f"rank {rank}: sharded MSM differs"
    # point-sharded commitments inside whole provers (every rank runs the same prover; the commitment MSMs are split
    # by point range and summed over NVLink): proofs byte-identical to the single-process oracle on every rank
    NVK = 16
    okzg = O.Kzg(O.rand_fr(7, NVK))
    kzg = hl.MultilinearKzg(ctx, [okzg.eqs(k) for k in range(NVK + 1)])
    hl.dist_shard_commits(ctx, True)
    poly = O.rand_fr(8100, 1 << 15)
    point = O.rand_fr(8101, 15)
    to = O.Transcript()
    okzg.open(to, poly, point)
    tr = hl.Keccak256Transcript(ctx)
    kzg.open(hl.MultilinearPolynomial.new(ctx, poly), point)
    assert tr.into_proof() == to.proof(), f"rank {rank}: commit-sharded KZG opening differs"
    for kind, chunks, mu in ((O.TABLE_RANGE, 4, 14), (O.TABLE_AND, 4, 15)):
        xs, ys = O.rand_u64s(8200 + mu, 1 << mu), O.rand_u64s(8300 + mu, 1 << mu)
        if kind == O.TABLE_AND:
            xs &= np.uint64((1 << (8 * chunks)) - 1)
            ys &= np.uint64((1 << (8 * chunks)) - 1)
        else:
            ys = None
        to = O.Transcript()
        assert O.lasso_prove(okzg, to, kind, chunks, mu, xs, ys)
        tr = hl.Keccak256Transcript(ctx)
        hl.LassoProver(ctx, kzg, kind, chunks).prove(xs, ys)
        assert tr.into_proof() == to.proof(), f"rank {rank}: commit-sharded Lasso proof differs (kind {kind})"
    hl.dist_shard_commits(ctx, False)
    check_sharded_sumchecks_in_lasso(ctx, kzg, okzg, rank, world, ((O.TABLE_RANGE, 4, 15), (O.TABLE_AND, 8, 12)))
    check_fully_sharded_lasso(ctx, kzg, okzg, rank, world, ((O.TABLE_RANGE, 4, 14, 12, 64), (O.TABLE_AND, 8, 12, 9, 16),
                                                             (O.TABLE_XOR, 2, 15, 13, 1 << 14)))
    dist.barrier()
    if rank == 0:
        print(f"SHARDED_OK world={world}")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
