"""Generates tests/golden/gkr_golden.json with the pure-Python model of the fractional sum-check (pymodel_gkr.py).
Inputs: the documented splitmix64 stream (pymodel.rand_fr), seeds as listed."""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import pymodel as M
import pymodel_gkr as G

cases = []
for B, n, seed, claimed in ((1, 1, 900, False), (3, 4, 910, False), (2, 6, 920, True), (3, 7, 930, False)):
    ps = [M.rand_fr(seed + b, 1 << n) for b in range(B)]
    qs = [M.rand_fr(seed + 50 + b, 1 << n) for b in range(B)]
    tr = M.Transcript()
    cl = [0] * B if claimed else None  # Some(_): absorbed instead of written
    p_xs, q_xs, x, p0, q0 = G.prove(tr, ps, qs, cl, cl)
    # what the reference's own test asserts (:355-366): the claims are the inputs' evaluations at x
    for b in range(B):
        assert M.evaluate(ps[b], x) == p_xs[b] and M.evaluate(qs[b], x) == q_xs[b]
    # and the statement itself: Σ_i p_i / q_i = p_0 / q_0
    for b in range(B):
        s = sum(p * pow(q, -1, M.R) for p, q in zip(ps[b], qs[b])) % M.R
        assert s * q0[b] % M.R == p0[b]
    cases.append({"batch": B, "num_vars": n, "seed": seed, "claimed": claimed, "proof": tr.stream.hex(),
                  "p_xs": [str(v) for v in p_xs], "q_xs": [str(v) for v in q_xs], "x": [str(v) for v in x],
                  "p_0s": [str(v) for v in p0], "q_0s": [str(v) for v in q0]})
    print(f"batch {B} n {n}: {len(tr.stream)} proof bytes", flush=True)
json.dump({"cases": cases}, open(os.path.join(HERE, "gkr_golden.json"), "w"))
