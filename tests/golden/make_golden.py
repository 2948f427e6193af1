"""Regenerates tests/golden/golden.json from the independent pure-Python model (pymodel.py).
Run:  python tests/golden/make_golden.py      (takes ~1 minute; output is committed)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pymodel as M

out = {}
out["keccak"] = {h: M.keccak256(bytes.fromhex(h)).hex() for h in ["", "616263", "00" * 135, "ab" * 136, "cd" * 300]}

# transcript: write 5 elements, squeeze 2, common 1, write generator*5, squeeze
tr = M.Transcript()
fes = M.rand_fr(11, 6)
for f in fes[:5]:
    tr.write_fe(f)
c = [tr.squeeze(), tr.squeeze()]
tr.common_fe(fes[5])
tr.write_comm(M.g1_mul(M.G, 5))
c.append(tr.squeeze())
out["transcript"] = {"seed": 11, "challenges": [hex(x) for x in c], "stream": tr.stream.hex()}

out["eq_xy"] = {"seed": 12, "n": 4, "evals": [hex(x) for x in M.eq_xy(M.rand_fr(12, 4))]}

sc = []
for name, n, T, NP, seed in [("eq_a_b", 5, 1, 2, 100), ("batched_products", 4, 3, 2, 200), ("linear_g", 4, 4, 1, 300)]:
    polys = [M.rand_fr(seed + i, 1 << n) for i in range(T * NP)]
    w = M.rand_fr(seed + 50, T) if T > 1 else [1]
    y = M.rand_fr(seed + 51, n)
    claim = M.rand_fr(seed + 52, 1)[0]
    if name == "eq_a_b":  # the true sum, so the proof also verifies
        eq = M.eq_xy(y)
        claim = sum(e * a * b for e, a, b in zip(eq, polys[0], polys[1])) % M.R
    tr = M.Transcript()
    terms = [(w[t], list(range(t * NP, (t + 1) * NP))) for t in range(T)]
    ch, ev = M.sumcheck_prove_evals(tr, n, polys, y, terms, claim)
    sc.append({"name": name, "n": n, "T": T, "NP": NP, "seed": seed, "claim": hex(claim), "proof": tr.stream.hex(),
               "challenges": [hex(x) for x in ch], "evals": [hex(x) for x in ev]})
out["sumcheck_evals"] = sc

n, K, seed = 4, 3, 400
polys = [M.rand_fr(seed + i, 1 << n) for i in range(K)]
scal = M.rand_fr(seed + 50, K)
ys = [M.rand_fr(seed + 60 + k, n) for k in range(K)]
claim = M.rand_fr(seed + 70, 1)[0]
tr = M.Transcript()
ch, ev = M.sumcheck_prove_coeffs(tr, n, polys, [(scal[k], ys[k], k) for k in range(K)], claim)
out["sumcheck_coeffs"] = {"n": n, "K": K, "seed": seed, "proof": tr.stream.hex(), "challenges": [hex(x) for x in ch],
                          "evals": [hex(x) for x in ev]}

# KZG, 4 variables
nv = 4
ss = M.rand_fr(7, nv)
srs = M.kzg_setup(ss)
out["kzg_srs"] = {"seed": 7, "nv": nv, "eqs_level2": [[hex(p[0]), hex(p[1])] for p in srs[2]],
                  "eqs_top_last": [hex(srs[nv][-1][0]), hex(srs[nv][-1][1])]}
poly = M.rand_fr(600, 1 << nv)
cm = M.kzg_commit(srs, poly)
pt = M.rand_fr(601, nv)
tr = M.Transcript()
ev = M.kzg_open(srs, tr, poly, pt)
out["kzg_open"] = {"poly_seed": 600, "point_seed": 601, "commit": [hex(cm[0]), hex(cm[1])], "eval": hex(ev),
                   "proof": tr.stream.hex()}
polys = [M.rand_fr(900 + i, 1 << nv) for i in range(4)]
points = [M.rand_fr(950 + i, nv) for i in range(2)]
pairs = [(0, 0), (1, 1), (2, 1), (3, 0), (0, 1)]
evals = [(p, q, M.evaluate(polys[p], points[q])) for p, q in pairs]
tr = M.Transcript()
M.kzg_batch_open(srs, tr, nv, polys, points, evals)
out["kzg_batch_open"] = {"poly_seed": 900, "point_seed": 950, "pairs": pairs, "proof": tr.stream.hex()}
sc = M.rand_fr(33, 16)
res = M.msm(sc, srs[nv])
out["msm"] = {"seed": 33, "result": [hex(res[0]), hex(res[1])]}
out["g1"] = {"2G": [hex(v) for v in M.g1_mul(M.G, 2)], "rm1_G": [hex(v) for v in M.g1_mul(M.G, M.R - 1)]}

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json"), "w") as f:
    json.dump(out, f, indent=1)
print("written", {k: (len(v) if hasattr(v, "__len__") else 1) for k, v in out.items()})
