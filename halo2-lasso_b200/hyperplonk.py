"""Host mirror of `HyperPlonk::{preprocess, prove}` (pb/backend/hyperplonk.rs:97-291): circuit descriptions, the
reference's test-circuit fixtures and the binding of the library's HyperPlonk prover (csrc/hyperplonk.cu, which holds
the orchestration: instance polys, batch commits, LogUp m/h polys, permutation grand product, zero check with the
generic expression kernel, rotated evaluations, additive batch opening). Mirrors

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

from . import MultilinearPolynomial, _chk, _mont_consts, _p, lib
from .expression import BooleanHypercube, Expression, R_MOD, serialize_expression

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
    """`HyperPlonk<MultilinearKzg<Bn256>>` prover (vanilla plonk, with or without the LogUp lookup argument): a thin
    binding of `b200_hyperplonk_preprocess` / `b200_hyperplonk_prove` — the orchestration (hyperplonk.rs:97-291) runs
    inside the library (csrc/hyperplonk.cu), the circuit crosses the boundary as prefix-token expressions."""

    def __init__(self, ctx, kzg, info):
        """preprocess (hyperplonk.rs:97-162): commit the preprocess and permutation polynomials, compose the
        zero-check expression."""
        self.ctx, self.kzg, self.info = ctx, kzg, info
        k = info.k
        self.preprocess = [upload_ints(ctx, p) for p in info.preprocess_polys]
        ctok, ltok, consts = [], [], []
        for c in info.constraints:
            serialize_expression(c, ctok, consts)
        for lookup in info.lookups:
            ltok.append(len(lookup))
            for a, t in lookup:
                serialize_expression(a, ltok, consts)
                serialize_expression(t, ltok, consts)
        cm = _mont_consts(ctx, consts)
        flat = []
        for cyc in info.permutations:
            flat.append(len(cyc))
            for (p, r) in cyc:
                flat += [p, r]
        ctok = np.asarray(ctok, dtype=np.int32)
        ltok = np.asarray(ltok if ltok else [0], dtype=np.int32)
        flat = np.asarray(flat if flat else [0], dtype=np.int32)
        pidx = np.asarray(info.permutation_polys, dtype=np.int32)
        pre = (C.c_void_p * max(1, len(self.preprocess)))(*[p.dev for p in self.preprocess])
        self.h = C.c_void_p()
        _chk(lib().b200_hyperplonk_preprocess(
            ctx.h, C.c_int(k), C.c_int(info.num_instances), C.c_int(info.num_witness_polys), C.c_int(len(self.preprocess)),
            pre, C.c_int(len(info.constraints)), _p(ctok), C.c_int(len(ctok)), C.c_int(len(info.lookups)), _p(ltok),
            C.c_int(len(ltok) if info.lookups else 0), _p(cm), C.c_int(len(consts)), C.c_int(len(pidx)), _p(pidx),
            C.c_int(len(info.permutations)), _p(flat), C.c_int(getattr(info, "max_degree", 4)), C.byref(self.h)),
            "hyperplonk_preprocess")
        nz, deg, npolys = C.c_int(), C.c_int(), C.c_int()
        _chk(lib().b200_hyperplonk_info(self.h, C.byref(nz), C.byref(deg), C.byref(npolys)), "hyperplonk_info")
        self.num_z, self.degree, self.num_polys = nz.value, deg.value, npolys.value

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.ctx.h:
                lib().b200_hyperplonk_free(self.h)
                self.h = None
        except Exception:
            pass

    def commitments(self):
        """(preprocess_comms, permutation_comms): the verifier parameters (hyperplonk.rs:127-150)"""
        a = np.zeros((max(1, len(self.preprocess)), 8), dtype=np.uint64)
        b = np.zeros((max(1, len(self.info.permutation_polys)), 8), dtype=np.uint64)
        _chk(lib().b200_hyperplonk_commitments(self.h, _p(a), _p(b)), "hyperplonk_commitments")
        return a[: len(self.preprocess)], b[: len(self.info.permutation_polys)]

    def permutation_poly(self, i):
        out = np.zeros((1 << self.info.k, 4), dtype=np.uint64)
        _chk(lib().b200_hyperplonk_permutation_poly(self.h, C.c_int(i), _p(out)), "hyperplonk_permutation_poly")
        return out

    def prove(self, instances, witness_ints=None, witness_polys=None):
        """hyperplonk.rs:164-291; appends to the context transcript (create a Keccak256Transcript first)."""
        ctx = self.ctx
        inst = ints_to_mont(ctx, instances) if len(instances) else np.zeros((1, 4), dtype=np.uint64)
        wit = witness_polys if witness_polys is not None else [upload_ints(ctx, w) for w in witness_ints]
        ptrs = (C.c_void_p * len(wit))(*[p.dev for p in wit])
        _chk(lib().b200_hyperplonk_prove(self.h, _p(np.ascontiguousarray(inst)), C.c_int(len(instances)), ptrs),
             "hyperplonk_prove")
