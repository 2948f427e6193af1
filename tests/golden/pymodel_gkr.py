"""Independent pure-Python model of the reference's GKR for fractional sum-checks
(plonkish_backend/src/piop/gkr/fractional_sum_check.rs:87-190), built on pymodel's transcript and sum-check.
Written from the reference source; shares no code with oracle/ or the CUDA library. Values are canonical integers."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pymodel as M
from pymodel import R


def layer_up(p, q):
    """Layer::up (:62-85): halves (l, r) of the tables -> p_l q_r + p_r q_l, q_l q_r"""
    half = len(p) // 2
    return ([(p[i] * q[i + half] + p[i + half] * q[i]) % R for i in range(half)],
            [q[i] * q[i + half] % R for i in range(half)])


def prove(tr, ps, qs, claimed_p=None, claimed_q=None):
    """-> (p_xs, q_xs, x, p_0s, q_0s). claimed_*: per element None (write the layer-0 value) or not None (absorb it)."""
    B = len(ps)
    n = len(ps[0]).bit_length() - 1
    claimed_p = claimed_p or [None] * B
    claimed_q = claimed_q or [None] * B
    levels = [[(list(ps[b]), list(qs[b])) for b in range(B)]]  # levels[0] = inputs (2^n), then 2^(n-1), ...
    for _ in range(n):
        levels.append([layer_up(p, q) for p, q in levels[-1]])
    p0 = [levels[n][b][0][0] for b in range(B)]
    q0 = [levels[n][b][1][0] for b in range(B)]
    for cl, vals in ((claimed_p, p0), (claimed_q, q0)):  # :121-146
        for c, v in zip(cl, vals):
            (tr.common_fe if c is not None else tr.write_fe)(v)
    cp, cq, y = list(p0), list(q0), []
    for v in range(n):  # the layer with num_vars = v: halves of the tables with 2^(v+1) entries
        half = 1 << v
        tabs = []
        for p, q in levels[n - 1 - v]:
            tabs += [p[:half], p[half:], q[:half], q[half:]]
        if v == 0:
            x, evals = [], [t[0] for t in tabs]
        else:
            gamma = tr.squeeze()
            terms, claim, pw = [], 0, 1
            for b in range(B):
                terms += [(pw, [4 * b, 4 * b + 3]), (pw, [4 * b + 1, 4 * b + 2])]
                claim = (claim + pw * cp[b]) % R
                pw = pw * gamma % R
                terms.append((pw, [4 * b + 2, 4 * b + 3]))
                claim = (claim + pw * cq[b]) % R
                pw = pw * gamma % R
            x, evals = M.sumcheck_prove_evals(tr, v, tabs, y, terms, claim)
        for e in evals:
            tr.write_fe(e)
        mu = tr.squeeze()
        cp = [(evals[4 * b] + mu * (evals[4 * b + 1] - evals[4 * b])) % R for b in range(B)]
        cq = [(evals[4 * b + 2] + mu * (evals[4 * b + 3] - evals[4 * b + 2])) % R for b in range(B)]
        y = list(x) + [mu]
    return cp, cq, y, p0, q0
