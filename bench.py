#!/usr/bin/env python
"""bench.py — one JSON line per run (contract in the task statement).

Headline workload (BASELINE.json configs[2], "cfg3", the configuration the metric "Lasso prove ms
@2^20 lookups" is quoted on; it fits one GPU): 64-bit range check via Surge, c = 4 chunks of 16 bits,
one 2^16 identity subtable, m = 2^20 synthetic lookups, bn256 MultilinearKzg; one step = one FULL
Lasso proof (commitments, primary sum-check, memory-checking grand products, two batch openings).
  value : prove time in ms with the operands already resident in HBM (lower is better)
  e2e   : the same proof through b200_lasso_prove with HOST operands (pinned) + proof bytes read back
The same run also measures BASELINE configs[1] ("cfg2": ClassicSumCheck eq*a*b, n = 20) because the
metric's second half is "sumcheck GB/s vs HBM peak": reported under "sumcheck" and used for
`roofline` (dominant sum-check launch, timed live with CUDA events on the library stream).
N > 1: every rank proves an independent instance (weak scaling; no collective on the data path yet).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MU = 20
CHUNKS = 4
KIND_RANGE = 0
SC_VARS = 20
SC_TABLES = 3
SC_ALGO_BYTES = 32 * SC_TABLES * (4 * (1 << SC_VARS) - 3)  # 402,652,896
METRIC = "Lasso prove time @2^20 lookups (64-bit range via Surge, c=4x16-bit, bn256 MultilinearKzg)"
UNIT = "ms"
WORKLOAD = ("cfg3: 64-bit range check via Surge, C=4 x 16-bit subtables, 2^20 lookups, full Lasso proof; "
            "plus cfg2 sum-check (deg-3 eq*a*b, n=20) for the roofline")
SRS_SEED, X_SEED = 7, 5


# ---- synthetic inputs without the oracle: the documented splitmix64 stream (oracle/capi.cpp) -------
def sm64(seed, idx):
    import numpy as np

    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def rand_u64s(seed, n):
    import numpy as np

    return sm64(seed, np.arange(n, dtype=np.uint64))


def rand_canonical(seed, n):
    """n canonical 253-bit integers as (n,4) uint64 — element i uses sm64(seed, 4i..4i+3)."""
    import numpy as np

    raw = sm64(seed, np.arange(4 * n, dtype=np.uint64)).reshape(n, 4)
    raw[:, 3] &= np.uint64(0x1FFFFFFFFFFFFFFF)
    return raw


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc, self.lines, self.index = None, [], index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def mont_one():
    import numpy as np

    return np.array([0xAC96341C4FFFFFFB, 0x36FC76959F60CD29, 0x666EA36F7879462E, 0x0E0A77C19A07DF2F], dtype=np.uint64)


def cpu_lasso(steps, warmup, srs=None):
    """Restated reference algorithms (oracle/, C++ + OpenMP, all host threads): full 2^20 Lasso proof."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O

    ss = O.rand_fr(SRS_SEED, MU)
    t0 = time.perf_counter()
    kz = O.Kzg.from_eqs(ss, srs) if srs is not None else O.Kzg(ss)
    setup_s = time.perf_counter() - t0
    xs = rand_u64s(X_SEED, 1 << MU)
    times, plen = [], 0
    for it in range(warmup + steps):
        tr = O.Transcript()
        t0 = time.perf_counter()
        ok = O.lasso_prove(kz, tr, KIND_RANGE, CHUNKS, MU, xs, None)
        dt = time.perf_counter() - t0
        assert ok
        plen = len(tr.proof())
        if it >= warmup:
            times.append(dt)
    return 1e3 * sum(times) / len(times), O.num_threads(), setup_s, plen


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return
        steps, warmup = max(1, min(args.steps, 2)), 0
        ms, cores, setup_s, plen = cpu_lasso(steps, warmup)
        sample = f"whole 2^{MU}-lookup proof, {steps} timed repetition(s); SRS setup {setup_s:.1f} s untimed"
        print(json.dumps({
            "impl": "reference", "metric": METRIC, "value": ms, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "u256 (BN254 Fr/Fq, Montgomery)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "impl_note": "restated reference algorithms (C++/OpenMP oracle: the Rust "
                       "rayon prover cannot be built here, no cargo; Lasso itself is absent from the snapshot)",
                       "proof_bytes": plen},
            "cpu_baseline": {"value": ms, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ms, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import ctypes as C

    import numpy as np
    import torch

    import halo2_lasso_b200 as hl

    assert torch.cuda.is_available(), "bench.py needs a GPU: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = hl.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream, device=local_rank)
    dev = f"cuda:{local_rank}"
    m = 1 << MU

    def to_mont(raw):
        p = hl.MultilinearPolynomial.new(ctx, raw)
        hl._chk(hl.lib().b200_fr_convert(ctx.h, p.dev, p.dev, C.c_uint64(raw.shape[0]), C.c_int(1)), "fr_convert")
        return p

    # SRS: MultilinearKzg::setup on the device from seeded trapdoor scalars (one-off, untimed)
    ss = to_mont(np.concatenate([rand_canonical(SRS_SEED, MU), np.zeros((32 - MU, 4), dtype=np.uint64)])).evals()[:MU]
    t0 = time.perf_counter()
    kzg = hl.MultilinearKzg.setup(ctx, ss)
    ctx.sync()
    setup_ms = 1e3 * (time.perf_counter() - t0)
    prover = hl.LassoProver(ctx, kzg, KIND_RANGE, CHUNKS)

    xs_host = torch.from_numpy(rand_u64s(X_SEED + 100 * rank, m).view(np.int64)).pin_memory()
    xs_dev = xs_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    proof_len = [0]
    last_proof = [b""]

    def lasso_device():
        hl.Keccak256Transcript(ctx)
        prover.prove_dev(MU, xs_dev.data_ptr())

    def lasso_e2e():
        tr = hl.Keccak256Transcript(ctx)
        prover.prove(xs_host.numpy().view(np.uint64))
        last_proof[0] = tr.into_proof()
        proof_len[0] = len(last_proof[0])

    # cfg2 sum-check on resident tables
    n, N = SC_VARS, 1 << SC_VARS
    polys = [to_mont(rand_canonical(seed + 10 * rank, N)) for seed in (1, 2)]
    y = to_mont(np.concatenate([rand_canonical(3 + 10 * rank, n), np.zeros((32 - n, 4), dtype=np.uint64)])).evals()[:n]
    one = mont_one()

    def sumcheck_device():
        hl.Keccak256Transcript(ctx)
        # any claim yields a well-formed transcript (p(0) is derived, eval.rs:129); timing is claim-independent
        hl.ClassicSumCheck.prove_evals(ctx, n, polys, one.reshape(1, 4), y, one)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s0, s1 in ev:
            with torch.cuda.stream(stream):
                flush.zero_()  # L2 flush between timed iterations (outside the event pair)
            s0.record(stream)
            fn()
            s1.record(stream)
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        return sum(a.elapsed_time(b) for a, b in ev) / steps

    # N > 1: the SAME cfg2 sum-check sharded over the top log2(N) variables (2^20 entries per rank, i.e.
    # n = 20 + log2 N in total), partial sums exchanged inside the round kernels over NVLink peer memory
    sharded = None
    if world > 1 and world & (world - 1) == 0:
        hl.dist_init(ctx, rank, world)
        n_tot = SC_VARS + world.bit_length() - 1
        y_tot = to_mont(np.concatenate([rand_canonical(3, n_tot), np.zeros((32 - n_tot, 4), dtype=np.uint64)])).evals()[:n_tot]

        def sumcheck_sharded():
            hl.Keccak256Transcript(ctx)
            hl.sumcheck_prove_evals_sharded(ctx, n_tot, polys, one.reshape(1, 4), y_tot, one)

        sharded = (n_tot, sumcheck_sharded)
        # ONE proof on all N GPUs: every rank runs the same prover on the same lookups, the commitment MSMs (the
        # largest single cost) are split by point range and the partial commitments summed over NVLink
        xs_shared = torch.from_numpy(rand_u64s(X_SEED, m).view(np.int64)).to(dev)

        def lasso_cooperative():
            hl.Keccak256Transcript(ctx)
            prover.prove_dev(MU, xs_shared.data_ptr())

    # the reference's own `zero_check` criterion bench shape (plonkish_backend/benches/zero_check.rs): generic
    # EvaluationsProver on vanilla_plonk_expression, n = 20, expression compiled inside the library
    from halo2_lasso_b200.expression import vanilla_plonk_expression

    zc_expr = vanilla_plonk_expression(SC_VARS)
    zc_polys = [to_mont(rand_canonical(100 + i + 20 * rank, N)) for i in range(13)]
    zc_ch = [int(x) for x in rand_canonical(7, 3)[:, 0]]
    zc_zero = np.zeros(4, dtype=np.uint64)

    def zero_check_device():
        hl.Keccak256Transcript(ctx)
        hl.prove_expression_native(ctx, n, zc_expr, zc_polys, zc_ch, [y], zc_zero)

    sampler = ClockSampler(local_rank)
    sampler.start()
    ctx.launch_count(reset=True)
    ms = timed(lasso_device, args.steps, max(3, args.warmup))
    launches = ctx.launch_count(reset=True) // (args.steps + max(3, args.warmup))
    ms_e2e = timed(lasso_e2e, max(3, args.steps // 2), 3)
    ms_sc = timed(sumcheck_device, 20, 5)
    ms_sh = timed(sharded[1], 20, 5) if sharded else 0.0
    ms_zc = timed(zero_check_device, 5, 3)
    ms_co = ms_co_sc = 0.0
    if sharded:
        hl.dist_shard_commits(ctx, True)
        ms_co = timed(lasso_cooperative, args.steps, 3)
        phases_co = {}
        for tag, t in hl.profile(ctx, lasso_cooperative):
            if tag >= 1000:
                nm = hl.PHASE_NAMES.get(tag, str(tag))
                phases_co[nm] = round(phases_co.get(nm, 0.0) + t, 4)
        # opt-in (B200_BENCH_SHARD_SUMCHECKS=<min_vars>): additionally evaluate the prover's large sum-checks on each
        # rank's 1/N slice of the hypercube (b200_dist_shard_sumchecks; unverified on hardware in round 1)
        if os.environ.get("B200_BENCH_SHARD_SUMCHECKS"):
            hl.dist_shard_sumchecks(ctx, int(os.environ["B200_BENCH_SHARD_SUMCHECKS"]))
            ms_co_sc = timed(lasso_cooperative, args.steps, 3)
            hl.dist_shard_sumchecks(ctx, 0)
        hl.dist_shard_commits(ctx, False)
    clocks = sampler.stop()

    prof = hl.profile_rounds(ctx, sumcheck_device, SC_VARS, SC_TABLES)
    phases = {}
    for tag, t in hl.profile(ctx, lasso_device):
        if tag >= 1000:
            nm = hl.PHASE_NAMES.get(tag, str(tag))
            phases[nm] = round(phases.get(nm, 0.0) + t, 4)

    if world > 1:
        t = torch.tensor([ms, ms_e2e, ms_sc, ms_sh, ms_co, ms_co_sc], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e, ms_sc, ms_sh, ms_co, ms_co_sc = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
            "peak_source": peak_src}
    if prof:
        roof.update(prof)
        roof["frac"] = roof["achieved"] / peak
        # the same launch against the bound that actually limits it: Montgomery products on the IMAD (fmaheavy) pipe.
        # Round 1 of cfg2 = 2^(n-2) output pairs x (6 binding + 6 evaluation products). Peak: the multiplier itself
        # measured alone on a B200 (tools/micro/pipe_rates.cu, profiles/r01_pipe_rates.txt): 537.8 cycles per warp
        # product per SM sub-partition with 4 resident warps each -> 148 SMs x 4 x 32 lanes x 1.965 GHz / 537.8.
        prods = 12 * (1 << (SC_VARS - 2))
        imad_peak = 148 * 4 * 32 * 1.965 / 537.8
        roof["imad"] = {"products_per_launch": prods, "achieved_gproducts_s": prods / (roof["launch_ms"] * 1e-3) / 1e9,
                        "peak_gproducts_s": round(imad_peak, 1), "frac": prods / (roof["launch_ms"] * 1e-3) / 1e9 / imad_peak,
                        "note": "IMAD-bound kernel; launch time includes the ~29 us single-warp Fiat-Shamir tail"}
    # per-launch DRAM traffic of that kernel from the committed `ncu --set full` capture (profiles/), if present
    try:
        roof["traffic"] = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass

    # the proof the end-to-end leg produced, checked by the product's own CPU verifier (libb200verify.so, pairing form);
    # outside every timed region
    verified = {"accepted": None, "cpu_verify_ms": None}
    try:
        from halo2_lasso_b200 import verifier as V

        t0 = time.perf_counter()
        vk = V.MultilinearKzgVerifier.setup(ss)
        vt = V.ProofTranscript(last_proof[0])
        verified["accepted"] = bool(vk.lasso_verify(vt, KIND_RANGE, CHUNKS, MU) and vt.done())
        verified["cpu_verify_ms"] = round(1e3 * (time.perf_counter() - t0), 1)
    except Exception as e:  # never lose the bench line over the extra check
        verified["error"] = repr(e)[:200]

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        srs = [kzg.eqs(k) for k in range(MU + 1)]  # reuse the device SRS so the CPU leg skips its slow setup
        cms, cores, _, plen = cpu_lasso(1, 0, srs)
        cpu = {"value": cms, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"whole 2^{MU}-lookup proof once (C++/OpenMP restatement of the reference algorithms)",
               "proof_bytes": plen}

    print(json.dumps({
        "metric": METRIC, "value": ms, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
        "ms_per_step": ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "u256 (BN254 Fr/Fq, 8x32-bit Montgomery limbs)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "lookups": m, "chunks": CHUNKS, "subtable": 1 << 16, "proof_bytes": proof_len[0],
                   "l2": "flushed between timed iterations (256 MiB write)", "replicas_per_gpu": 1,
                   "srs_setup_ms_untimed": round(setup_ms, 1)},
        "lookups_per_s": world * m / (ms * 1e-3),
        "clocks": clocks,
        "e2e": {"value": ms_e2e, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": m * 8,
                "d2h_bytes_per_step": proof_len[0]},
        "gpu_launches": launches, "phases_ms": phases,
        "sumcheck": {"workload": "cfg2: ClassicSumCheck deg-3 eq*a*b, n=20, 2560 proof bytes", "ms_per_proof": ms_sc,
                     "algorithmic_bytes": SC_ALGO_BYTES, "GBps": world * SC_ALGO_BYTES / (ms_sc * 1e-3) / 1e9,
                     "frac_of_hbm_peak": SC_ALGO_BYTES / (ms_sc * 1e-3) / 1e9 / peak},
        "zero_check": {"workload": "reference zero_check bench shape: vanilla_plonk_expression (17 tables, degree 5), n=20, "
                       "generic bytecode kernels, through b200_sumcheck_prove_expression (host call, includes its "
                       "small H2D/D2H)", "ms_per_proof": ms_zc},
        "sumcheck_sharded": None if not sharded else {
            "workload": f"cfg2 shape sharded on the top {world.bit_length() - 1} variable(s): n={sharded[0]}, 2^20 entries per GPU, "
                        "per-round partials exchanged inside the kernel over NVLink peer memory",
            "ms_per_proof": ms_sh, "algorithmic_bytes": 32 * SC_TABLES * (4 * (1 << sharded[0]) - 3),
            "GBps": 32 * SC_TABLES * (4 * (1 << sharded[0]) - 3) / (ms_sh * 1e-3) / 1e9},
        "lasso_commit_sharded": None if not sharded else {
            "workload": f"ONE {WORKLOAD} proof on {world} GPUs: commitment MSMs point-sharded, partial commitments summed "
                        "over NVLink peer memory, sum-checks replicated (strong scaling of the proof latency)",
            "ms_per_proof": ms_co, "phases_ms": phases_co,
            "ms_per_proof_with_sharded_sumchecks": ms_co_sc if ms_co_sc else None},
        "proof_verified": verified, "roofline": roof, "cpu_baseline": cpu}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
