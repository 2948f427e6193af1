"""Generates tests/golden/lasso_golden.json with the pure-Python Lasso model (pymodel_lasso.py): a few minutes of
CPU. Inputs come from the documented splitmix64 stream (pymodel.sm64 / rand_fr)."""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import pymodel as M
import pymodel_lasso as L

cases = []
for kind, c, mu, seed in ((L.XOR, 2, 2, 41), (L.RANGE, 2, 3, 42), (L.AND, 3, 2, 43)):
    m = 1 << mu
    bits = (16 if kind == L.RANGE else 8) * c
    xs = [M.sm64(seed, i) & ((1 << bits) - 1) for i in range(m)]
    ys = [M.sm64(seed + 1, i) & ((1 << bits) - 1) for i in range(m)] if kind != L.RANGE else None
    xs[m // 2:] = xs[: m // 2]  # repeated addresses: read_ts != 0
    if ys:
        ys[m // 2:] = ys[: m // 2]
    ss = M.rand_fr(7, max(mu, L.SUB_VARS))
    t0 = time.time()
    proof = L.prove(ss, kind, c, mu, xs, ys)
    print(f"kind {kind} c {c} mu {mu}: {len(proof)} proof bytes in {time.time() - t0:.0f} s", flush=True)
    cases.append({"kind": kind, "chunks": c, "mu": mu, "srs_seed": 7, "xs": [str(x) for x in xs],
                  "ys": [str(y) for y in ys] if ys else None, "proof": proof.hex()})
# a table given as data: 8|8-bit OR with a 9-bit output stride (overlapping outputs: g is a SUM, not a concatenation)
vals = [(x >> 8) | (x & 0xFF) for x in range(1 << 16)]
tab = L.CustomTable(2, 8, 9, vals)
mu, c, seed = 2, 2, 44
xs = [M.sm64(seed, i) & 0xFFFF for i in range(1 << mu)]
ys = [M.sm64(seed + 1, i) & 0xFFFF for i in range(1 << mu)]
xs[2:], ys[2:] = xs[:2], ys[:2]
t0 = time.time()
proof = L.prove(M.rand_fr(7, L.SUB_VARS), L.CUSTOM, c, mu, xs, ys, tab)
print(f"custom OR c {c} mu {mu}: {len(proof)} proof bytes in {time.time() - t0:.0f} s", flush=True)
cases.append({"kind": L.CUSTOM, "chunks": c, "mu": mu, "srs_seed": 7, "xs": [str(x) for x in xs], "ys": [str(y) for y in ys],
              "table": {"num_operands": 2, "operand_bits": 8, "out_bits": 9, "rule": "or8"}, "proof": proof.hex()})
json.dump({"cases": cases}, open(os.path.join(HERE, "lasso_golden.json"), "w"))
