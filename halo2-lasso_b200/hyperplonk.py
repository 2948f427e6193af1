"""Host orchestration of `HyperPlonk::{preprocess, prove}` (pb/backend/hyperplonk.rs:97-291) over the GPU
primitives, for circuits without lookups (the snapshot's LogUp branch is empty when `lookups` is empty,
prover.rs:56-58): instance polys, batch commits, permutation grand product, zero check with the generic
expression kernel, rotated evaluations, additive batch opening. Mirrors

    PlonkishCircuitInfo            pb/backend.rs:46-73               -> VanillaPlonkCircuitInfo
    rand_vanilla_plonk_circuit     pb/backend/hyperplonk/util.rs:100-190 (own satisfiable fixture, same shape)
    permutation_polys              pb/backend/hyperplonk/preprocessor.rs:172-203
    instance_polys / row_mapping   prover.rs:32-48, hyperplonk.rs:365-369
    prove_sum_check, pcs_query,    prover.rs:368-409, verifier.rs:147-182
    points, point_offset
    rotation_eval_points           pb/poly/multilinear.rs:475-541

Values on the host side are canonical Python ints; tables are converted to Montgomery form on the device.
"""
import ctypes as C
import random

import numpy as np

from . import (Keccak256Transcript, MultilinearPolynomial, _chk, _fr, _p, evaluate_many, expression_rows, lib,
               lookup_h_poly, lookup_m_poly, prove_expression)
from .expression import BooleanHypercube, Expression, R_MOD, compose

R_INV = pow(1 << 256, -1, R_MOD)


def mont_to_int(limbs):
    v = sum(int(limbs[k]) << (64 * k) for k in range(4))
    return v * R_INV % R_MOD


def ints_to_raw(vals):
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i, 0] = v & 0xFFFFFFFFFFFFFFFF
        out[i, 1] = (v >> 64) & 0xFFFFFFFFFFFFFFFF
        out[i, 2] = (v >> 128) & 0xFFFFFFFFFFFFFFFF
        out[i, 3] = v >> 192
    return out


def upload_ints(ctx, vals):
    """canonical ints -> device polynomial in Montgomery form"""
    p = MultilinearPolynomial.new(ctx, ints_to_raw(vals))
    _chk(lib().b200_fr_convert(ctx.h, p.dev, p.dev, C.c_uint64(len(vals)), C.c_int(1)), "fr_convert")
    return p


def ints_to_mont(ctx, vals):
    n = max(1, len(vals))
    k = 1 << (n - 1).bit_length()
    p = upload_ints(ctx, list(vals) + [0] * (k - len(vals)))
    return p.evals()[: len(vals)]


class VanillaPlonkCircuitInfo:
    """polys: 0 pi | 1-5 q_l q_r q_m q_o q_c | 6-8 w_l w_r w_o; permutation over the three wire columns."""

    def __init__(self, k, num_instances, preprocess_polys, permutations):
        self.k, self.num_instances = k, num_instances
        self.preprocess_polys = preprocess_polys  # 5 lists of ints
        self.permutations = permutations          # cycles of (poly, row)
        self.permutation_polys = [6, 7, 8]
        self.num_witness_polys = 3
        pi, q_l, q_r, q_m, q_o, q_c, w_l, w_r, w_o = (Expression.polynomial(i) for i in range(9))
        self.constraints = [q_l * w_l + q_r * w_r + q_m * w_l * w_r + q_o * w_o + q_c + pi]
        self.num_poly = 9
        self.lookups = []


class VanillaPlonkWithLookupCircuitInfo:
    """util.rs:63-86 — polys: 0 pi | 1-9 q_l q_r q_m q_o q_c q_lookup t_l t_r t_o | 10-12 w_l w_r w_o; one lookup of
    width 3: (q_lookup*w_l, t_l), (q_lookup*w_r, t_r), (q_lookup*w_o, t_o)."""

    def __init__(self, k, num_instances, preprocess_polys, permutations):
        self.k, self.num_instances = k, num_instances
        self.preprocess_polys = preprocess_polys  # 9 lists of ints
        self.permutations = permutations
        self.permutation_polys = [10, 11, 12]
        self.num_witness_polys = 3
        pi, q_l, q_r, q_m, q_o, q_c, q_lookup, t_l, t_r, t_o, w_l, w_r, w_o = (Expression.polynomial(i) for i in range(13))
        self.constraints = [q_l * w_l + q_r * w_r + q_m * w_l * w_r + q_o * w_o + q_c + pi]
        self.lookups = [[(q_lookup * w_l, t_l), (q_lookup * w_r, t_r), (q_lookup * w_o, t_o)]]
        self.num_poly = 13


def rand_vanilla_plonk_with_lookup_circuit(k, seed, num_instances=None, lookup_fraction=0.4):
    """A random SATISFIABLE vanilla-plonk circuit with copy constraints AND lookup rows (same shape as
    util.rs:216-316: table rows 0 and 1 hold the zero tuple so that gated-off rows look up (0,0,0)).
    Returns (circuit_info, instances, witness polys [w_l, w_r, w_o]) as canonical ints."""
    rng = random.Random(seed)
    N = 1 << k
    bh = BooleanHypercube(k)
    num_instances = min(k, N - 2) if num_instances is None else num_instances
    q = [[0] * N for _ in range(9)]  # q_l q_r q_m q_o q_c q_lookup t_l t_r t_o
    w = [[0] * N for _ in range(3)]
    for t in (6, 7, 8):
        for b in range(2, N):
            q[t][b] = rng.randrange(R_MOD)
    instances = [rng.randrange(R_MOD) for _ in range(num_instances)]
    used = {0}
    for i, v in enumerate(instances):  # instance rows: -w_l + pi = 0
        b = bh.nth(i + 1)
        q[0][b], w[0][b] = R_MOD - 1, v
        used.add(b)
    cycles, outputs = [], []
    for b in range(1, N):
        if b in used:
            continue
        if rng.random() < lookup_fraction or not any(q[5]):  # lookup row: the wire tuple is a table row (repeats give m > 1)
            idx = rng.randrange(1, N) if rng.random() < 0.7 or b < 3 else rng.randrange(1, min(N, 4))
            q[5][b] = 1
            w[0][b], w[1][b], w[2][b] = q[6][idx], q[7][idx], q[8][idx]
            continue
        if outputs and rng.random() < 0.5:  # copy an earlier output into w_l
            src = outputs.pop(rng.randrange(len(outputs)))
            w[0][b] = w[2][src]
            cycles.append([(12, src), (10, b)])
        else:
            w[0][b] = rng.randrange(R_MOD)
        w[1][b] = rng.randrange(R_MOD)
        if rng.randrange(2) == 0:
            q[0][b] = q[1][b] = 1
            w[2][b] = (w[0][b] + w[1][b]) % R_MOD
        else:
            q[2][b] = 1
            w[2][b] = w[0][b] * w[1][b] % R_MOD
        q[3][b] = R_MOD - 1
        outputs.append(b)
    return VanillaPlonkWithLookupCircuitInfo(k, num_instances, q, cycles), instances, w


def rand_vanilla_plonk_circuit(k, seed, num_instances=None):
    """A random SATISFIABLE vanilla-plonk circuit with copy constraints (same shape as util.rs:100-190).
    Returns (circuit_info, instances, witness polys [w_l, w_r, w_o]) as canonical ints."""
    rng = random.Random(seed)
    N = 1 << k
    order = BooleanHypercube(k).iter()
    num_instances = k if num_instances is None else num_instances
    q = [[0] * N for _ in range(5)]  # q_l q_r q_m q_o q_c
    w = [[0] * N for _ in range(3)]
    instances = [rng.randrange(R_MOD) for _ in range(num_instances)]
    used = {0}
    for i, v in enumerate(instances):  # instance rows: -w_l + pi = 0
        b = order[i + 1]
        q[0][b], w[0][b] = R_MOD - 1, v
        used.add(b)
    cycles, outputs = [], []
    for b in range(1, N):
        if b in used:
            continue
        if outputs and rng.random() < 0.5:  # copy an earlier output into w_l
            src = outputs.pop(rng.randrange(len(outputs)))
            w[0][b] = w[2][src]
            cycles.append([(8, src), (6, b)])
        else:
            w[0][b] = rng.randrange(R_MOD)
        w[1][b] = rng.randrange(R_MOD)
        kind = rng.randrange(3)
        if kind == 0:  # addition
            q[0][b] = q[1][b] = 1
            w[2][b] = (w[0][b] + w[1][b]) % R_MOD
        elif kind == 1:  # multiplication
            q[2][b] = 1
            w[2][b] = w[0][b] * w[1][b] % R_MOD
        else:  # affine with a constant
            c = rng.randrange(R_MOD)
            q[0][b], q[4][b] = 3, c
            w[2][b] = (3 * w[0][b] + c) % R_MOD
        q[3][b] = R_MOD - 1
        outputs.append(b)
    return VanillaPlonkCircuitInfo(k, num_instances, q, cycles), instances, w


def permutation_polys(k, permutation_polys_idx, cycles):
    """preprocessor.rs:172-203"""
    N = 1 << k
    index = {p: i for i, p in enumerate(permutation_polys_idx)}
    perms = [[(i << k) + j for j in range(N)] for i in range(len(permutation_polys_idx))]
    for cyc in cycles:
        i0, j0 = cyc[0]
        last = perms[index[i0]][j0]
        for t in range(1, len(cyc) + 1):
            i, j = cyc[t % len(cyc)]
            assert j != 0
            perms[index[i]][j], last = last, perms[index[i]][j]
    return perms


def rotation_eval_point_pattern(next_, num_vars, distance):
    bh = BooleanHypercube(num_vars)
    rem = bh.primitive if next_ else bh.x_inv
    pat = [0] * (1 << distance)
    for depth in range(distance):
        step = 1 << (distance - depth)
        for e in range(0, len(pat), step):
            o = e + (step >> 1)
            rotated = pat[e] << 1 if next_ else pat[e] >> 1
            pat[o], pat[e] = rotated ^ rem, rotated
    return pat


def rotation_eval_points(x, rotation):
    """multilinear.rs:475-517 on canonical ints"""
    if rotation == 0:
        return [list(x)]
    n, d = len(x), abs(rotation)
    num_x = n - d
    out = []
    if rotation < 0:
        for pat in rotation_eval_point_pattern(False, n, d):
            p = [(1 - x[d + i]) % R_MOD if (pat >> i) & 1 else x[d + i] for i in range(num_x)]
            p += [(pat >> (i + num_x)) & 1 for i in range(d)]
            out.append(p)
    else:
        for pat in rotation_eval_point_pattern(True, n, d):
            p = [(pat >> i) & 1 for i in range(d)]
            p += [(1 - x[i]) % R_MOD if (pat >> (i + d)) & 1 else x[i] for i in range(num_x)]
            out.append(p)
    return out


class HyperPlonk:
    """`HyperPlonk<MultilinearKzg<Bn256>>` prover (vanilla plonk, with or without the LogUp lookup argument)."""

    def __init__(self, ctx, kzg, info):
        """preprocess (hyperplonk.rs:97-162): commit the preprocess and permutation polynomials, compose the
        zero-check expression."""
        self.ctx, self.kzg, self.info = ctx, kzg, info
        k = info.k
        self.preprocess = [upload_ints(ctx, p) for p in info.preprocess_polys]
        self.preprocess_comms = kzg.batch_commit(self.preprocess)
        self.perm_ints = permutation_polys(k, info.permutation_polys, info.permutations)
        self.perm = [upload_ints(ctx, p) for p in self.perm_ints]
        self.permutation_comms = kzg.batch_commit(self.perm)
        self.num_z, self.expression = compose(k, info.constraints, info.num_poly, info.permutation_polys,
                                              lookups=info.lookups)
        assert self.num_z == 1

    def prove(self, instances, witness_ints=None, witness_polys=None):
        """hyperplonk.rs:164-291; appends to the context transcript (create a Keccak256Transcript first)."""
        ctx, kzg, info, k = self.ctx, self.kzg, self.info, self.info.k
        tr = Keccak256Transcript.__new__(Keccak256Transcript)
        tr.ctx = ctx
        inst_mont = ints_to_mont(ctx, instances)
        tr.common_field_elements(inst_mont)
        # instance_polys (prover.rs:32-48): instance i sits on row bh[i+1]
        order = BooleanHypercube(k)
        raw = np.zeros((1 << k, 4), dtype=np.uint64)
        b = 1
        for row in ints_to_raw(instances):
            raw[b] = row
            b = order.next(b)
        inst_poly = MultilinearPolynomial.new(ctx, raw)
        _chk(lib().b200_fr_convert(ctx.h, inst_poly.dev, inst_poly.dev, C.c_uint64(1 << k), C.c_int(1)), "fr_convert")
        wit = witness_polys if witness_polys is not None else [upload_ints(ctx, w) for w in witness_ints]
        kzg.batch_commit_and_write(wit)
        polys = [inst_poly] + self.preprocess + wit
        # LogUp (prover.rs:50-250): compressed input/table polys, multiplicities m, then h once gamma is known
        beta = tr.squeeze_challenge()
        compressed, ms = [], []
        for lookup in info.lookups:
            b_ = Expression.challenge(0)
            ci = expression_rows(ctx, k, Expression.distribute_powers([a for a, _ in lookup], b_), polys, [mont_to_int(beta)])
            ct = expression_rows(ctx, k, Expression.distribute_powers([t for _, t in lookup], b_), polys, [mont_to_int(beta)])
            compressed.append((ci, ct))
            ms.append(lookup_m_poly(ctx, k, ci, ct))
        if ms:
            kzg.batch_commit_and_write(ms)
        gamma = tr.squeeze_challenge()
        hs = [lookup_h_poly(ctx, k, ci, ct, m, gamma) for (ci, ct), m in zip(compressed, ms)]
        # permutation_z_polys (prover.rs:252-345)
        z = MultilinearPolynomial.alloc(ctx, k)
        nper = len(info.permutation_polys)
        wires = (C.c_void_p * nper)(*[wit[p - (info.num_poly - info.num_witness_polys)].dev for p in info.permutation_polys])
        sig = (C.c_void_p * nper)(*[p.dev for p in self.perm])
        offs = (C.c_uint64 * nper)(*[i << k for i in range(nper)])
        bg = np.ascontiguousarray(np.stack([beta, gamma]))
        _chk(lib().b200_permutation_z(ctx.h, C.c_int(k), C.c_int(nper), wires, sig, offs, _p(bg), z.dev), "permutation_z")
        kzg.batch_commit_and_write(hs + [z])
        alpha = tr.squeeze_challenge()
        y = tr.squeeze_challenges(k)
        polys = polys + self.perm + ms + hs + [z]
        challenges = [mont_to_int(c) for c in (beta, gamma, alpha)]
        zero = np.zeros(4, dtype=np.uint64)
        x, evals = prove_expression(ctx, k, self.expression, polys, challenges, [y], zero)
        # prove_sum_check tail (prover.rs:388-409): evaluations per pcs_query, rotated ones at rotation_eval_points
        queries = sorted({(l[1], l[2]) for l in self.expression.leaves() if l[0] == "poly" and l[1] >= 1})
        rotations = sorted({r for _, r in queries})
        x_int = [mont_to_int(v) for v in x]
        points_int, offset = [], {}
        for r in rotations:
            offset[r] = len(points_int)
            points_int += rotation_eval_points(x_int, r)
        points = [x if r == 0 and i == 0 else None for r in rotations for i in range(1 << abs(r))]
        pts_mont = ints_to_mont(ctx, [v for p in points_int for v in p]).reshape(len(points_int), k, 4)
        points = [pts_mont[i] for i in range(len(points_int))]
        ev_list = []
        for (p, r) in queries:
            if r == 0:
                ev_list.append((p, offset[0], evals[p]))
            else:
                for j in range(1 << abs(r)):
                    ev_list.append((p, offset[r] + j, evaluate_many(ctx, [polys[p]], points[offset[r] + j])[0]))
        tr.write_field_elements(np.stack([e[2] for e in ev_list]))
        kzg.batch_open(polys, points, ev_list)
