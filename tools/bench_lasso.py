"""Lasso prove timing for any table: bench_lasso.py KIND(range|and|xor) CHUNKS MU"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import halo2_lasso_b200 as hl
from halo2_lasso_b200 import hyperplonk as H
from bench import rand_canonical, rand_u64s

kind = {"range": 0, "and": 1, "xor": 2}[sys.argv[1]]
chunks, mu = int(sys.argv[2]), int(sys.argv[3])
ctx = hl.Context(0)
nv = max(mu, 16)
ss = H.ints_to_mont(ctx, [int(sum(int(v[j]) << (64 * j) for j in range(4))) for v in rand_canonical(7, nv)])
t0 = time.time()
kzg = hl.MultilinearKzg.setup(ctx, ss)
ctx.sync()
setup_s = time.time() - t0
bits = (16 if kind == 0 else 8) * chunks
mask = np.uint64((1 << bits) - 1) if bits < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
xs = torch.from_numpy((rand_u64s(5, 1 << mu) & mask).view(np.int64)).cuda()
ys = torch.from_numpy((rand_u64s(6, 1 << mu) & mask).view(np.int64)).cuda()
prover = hl.LassoProver(ctx, kzg, kind, chunks)
stream = torch.cuda.ExternalStream(ctx.stream)
times = []
for it in range(6):
    tr = hl.Keccak256Transcript(ctx)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    prover.prove_dev(mu, xs.data_ptr(), ys.data_ptr() if kind else None)
    e1.record(stream)
    ctx.sync(); torch.cuda.synchronize()
    if it >= 2:
        times.append(e0.elapsed_time(e1))
phases = {}
for tag, t in hl.profile(ctx, lambda: (hl.Keccak256Transcript(ctx), prover.prove_dev(mu, xs.data_ptr(), ys.data_ptr() if kind else None))):
    if tag >= 1000:
        nm = hl.PHASE_NAMES.get(tag, str(tag)); phases[nm] = round(phases.get(nm, 0.0) + t, 3)
print(json.dumps({"bench": f"Lasso prove {sys.argv[1]} c={chunks} 2^{mu} lookups", "ms": sum(times) / len(times),
                  "proof_bytes": len(hl.Keccak256Transcript.into_proof(tr)), "srs_setup_s": round(setup_s, 2),
                  "mem_GiB": round(torch.cuda.max_memory_allocated() / 2**30, 2), "phases_ms": phases}))
