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
# b200_synthesize_fn (include/b200_lasso.h)
SYNTHESIZE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p))


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


class TwoPhaseCircuitInfo:
    """A circuit with TWO instance columns and TWO witness phases (the shape `Halo2Circuit` yields for circuits with
    challenges, pb/frontend/halo2.rs:150-160; pb/backend.rs:50-60). polys:
        0 pi_a | 1 pi_b | 2 q_mix 3 q_io 4 q_io2 5 q_mul 6 q_lk 7 t | 8 a 9 b (phase 0) | 10 c 11 d (phase 1)
    challenges 0, 1 = r0, r1 squeezed after the phase-0 commitments; c = a + r0*b and d = r1*a*b can only be
    synthesized once they are known. pi_b is queried at Rotation::next (instance_evals, verifier.rs:92-145); the
    lookup input q_lk*(c - r0*b) uses a circuit challenge inside a lookup expression (prover.rs:66-76)."""

    def __init__(self, k, num_instances, preprocess_polys, permutations, with_lookup=True):
        self.k, self.num_instances = k, list(num_instances)
        self.preprocess_polys = preprocess_polys  # 6 lists of ints
        self.permutations = permutations
        self.permutation_polys = [8, 9, 10]
        self.num_witness_polys = [2, 2]
        self.num_challenges = [2, 0]
        P, Ch = Expression.polynomial, Expression.challenge
        pi_a, pi_b_next = P(0), P(1, 1)
        q_mix, q_io, q_io2, q_mul, q_lk, t, a, b, c, d = (P(i) for i in range(2, 12))
        self.constraints = [q_mix * (a + Ch(0) * b - c), q_io * (a - pi_a), q_io2 * (b - pi_b_next),
                            q_mul * (Ch(1) * a * b - d)]
        self.lookups = [[(q_lk * (c - Ch(0) * b), t)]] if with_lookup else []
        self.num_poly = 12


def rand_two_phase_circuit(k, seed, with_lookup=True):
    """A random SATISFIABLE instance of TwoPhaseCircuitInfo. Returns (info, instance columns [pi_a, pi_b],
    synthesize) where synthesize(round, challenges) -> witness columns of that phase, all canonical ints."""
    assert k >= 4
    rng = random.Random(seed)
    N = 1 << k
    order = BooleanHypercube(k).iter()
    q = [[0] * N for _ in range(6)]  # q_mix q_io q_io2 q_mul q_lk t
    a = [rng.randrange(R_MOD) for _ in range(N)]
    b = [rng.randrange(R_MOD) for _ in range(N)]
    for r in range(1, N):
        q[5][r] = rng.randrange(R_MOD)  # table; row 0 holds 0 for the gated-off rows
    inst_a = [rng.randrange(R_MOD) for _ in range(2)]
    inst_b = [rng.randrange(R_MOD) for _ in range(3)]
    used = {0}
    for j, v in enumerate(inst_a):  # pi_a at Rotation::cur: instance j sits on row order[j + 1]
        row = order[j + 1]
        q[1][row], a[row] = 1, v
        used.add(row)
    for j, v in enumerate(inst_b):  # pi_b at Rotation::next: read from the row BEFORE order[j + 1] in the LFSR cycle
        row = order[j] if j else order[N - 1]
        q[2][row], b[row] = 1, v
        used.add(row)
    free = [r for r in range(1, N) if r not in used]
    rng.shuffle(free)
    cycles = []
    for _ in range(max(1, len(free) // 8)):  # a[t] = b[s]
        s_, t_ = free.pop(), free.pop()
        a[t_] = b[s_]
        cycles.append([(9, s_), (8, t_)])
    twins = []
    for _ in range(max(1, len(free) // 8)):  # rows with the same (a, b): their c values are copies of each other
        s_, t_ = free.pop(), free.pop()
        a[t_], b[t_] = a[s_], b[s_]
        cycles.append([(10, s_), (10, t_)])
        twins.append((s_, t_))
    twin_rows = {r for tw in twins for r in tw}
    for r in range(N):
        q[0][r] = 1 if r in twin_rows else rng.randrange(2)
        q[3][r] = rng.randrange(2)
    for r in free[: len(free) // 2]:  # lookup rows: a is a table value
        q[4][r] = 1
        a[r] = q[5][rng.randrange(N)]
        q[0][r] = 1  # the lookup input c - r0*b equals a only where the mix gate holds
    info = TwoPhaseCircuitInfo(k, [len(inst_a), len(inst_b)], q, cycles, with_lookup)

    fill = [[rng.randrange(R_MOD) for _ in range(N)] for _ in range(2)]  # c, d where their gates are off

    def synthesize(rnd, challenges):
        """PlonkishCircuit::synthesize (pb/backend.rs:100-110); deterministic: every prover sees the same witness"""
        if rnd == 0:
            assert len(challenges) == 0
            return [list(a), list(b)]
        r0, r1 = challenges
        c = [(x + r0 * y) % R_MOD if q[0][r] else fill[0][r] for r, (x, y) in enumerate(zip(a, b))]
        d = [r1 * x * y % R_MOD if q[3][r] else fill[1][r] for r, (x, y) in enumerate(zip(a, b))]
        return [c, d]

    return info, [inst_a, inst_b], synthesize


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


def range_checked_plonk_circuit(k, seed, bits=32, num_instances=2):
    """A satisfiable vanilla-plonk circuit whose output wire w_o holds a `bits`-bit value in EVERY row (additions of two
    (bits - 1)-bit operands, copies between rows): the circuit shape of BASELINE cfg1, where a Lasso range check over the
    whole column w_o is the circuit's lookup argument (`HyperPlonkLasso`). Returns (circuit_info, instances, [w_l, w_r, w_o])."""
    rng = random.Random(seed)
    N = 1 << k
    order = BooleanHypercube(k).iter()
    q = [[0] * N for _ in range(5)]  # q_l q_r q_m q_o q_c
    w = [[0] * N for _ in range(3)]
    instances = [rng.randrange(1 << (bits - 1)) for _ in range(num_instances)]
    used = {0}
    for i, v in enumerate(instances):  # instance rows: -w_l + pi = 0 (w_o = 0 is in range)
        b = order[i + 1]
        q[0][b], w[0][b] = R_MOD - 1, v
        used.add(b)
    cycles, outputs = [], []
    for b in range(1, N):
        if b in used:
            continue
        if outputs and rng.random() < 0.3:  # copy an earlier output (halved so that the sum stays in range) ... as w_r
            src = outputs[rng.randrange(len(outputs))]
            w[0][b] = w[2][src] >> 1
        else:
            w[0][b] = rng.randrange(1 << (bits - 1))
        w[1][b] = rng.randrange(1 << (bits - 1))
        q[0][b] = q[1][b] = 1
        q[3][b] = R_MOD - 1
        w[2][b] = w[0][b] + w[1][b]
        if outputs and rng.random() < 0.25:  # an exact copy constraint between two output cells with equal values
            src = outputs[rng.randrange(len(outputs))]
            w[0][b], w[1][b] = w[0][src], w[1][src]
            w[2][b] = w[2][src]
            cycles.append([(8, src), (8, b)])
        outputs.append(b)
    assert all(v < (1 << bits) for v in w[2])
    return VanillaPlonkCircuitInfo(k, num_instances, q, cycles), instances, w


class HyperPlonkLasso:
    """Lasso as the lookup argument of a HyperPlonk-proved circuit (BASELINE cfg1): the HyperPlonk section proves the gate
    and copy constraints, the Lasso section — on the same transcript — proves that EVERY entry of one witness column lies
    in a decomposable table (range) or is the table's output for two operand columns; the two sections are linked by
    commitment equality: the commitment to `a` the Lasso section writes must be the commitment to that witness column the
    HyperPlonk section wrote (same SRS level: one lookup per row, mu = k). `HyperPlonkLassoVerifier` (verifier.py) checks
    both sections and the link."""

    def __init__(self, ctx, kzg, info, kind, chunks, lookup_witness):
        from . import LassoProver

        self.hp = HyperPlonk(ctx, kzg, info)
        self.lasso = LassoProver(ctx, kzg, kind, chunks)
        self.info, self.lookup_witness = info, lookup_witness

    def prove(self, instances, witness_ints, ys_ints=None):
        import numpy as np

        self.hp.prove(instances, witness_ints=witness_ints)
        xs = np.asarray(witness_ints[self.lookup_witness], dtype=np.uint64)
        self.lasso.prove(xs, None if ys_ints is None else np.asarray(ys_ints, dtype=np.uint64))


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
        # one instance column and one witness phase (the reference's own test circuits), or the general shape
        self.instance_cols = list(info.num_instances) if isinstance(info.num_instances, (list, tuple)) else [info.num_instances]
        self.phase_witness = list(info.num_witness_polys) if isinstance(info.num_witness_polys, (list, tuple)) else [info.num_witness_polys]
        self.phase_challenges = list(getattr(info, "num_challenges", [0] * len(self.phase_witness)))
        tail = (C.c_int(len(self.preprocess)),
                pre, C.c_int(len(info.constraints)), _p(ctok), C.c_int(len(ctok)), C.c_int(len(info.lookups)), _p(ltok),
                C.c_int(len(ltok) if info.lookups else 0), _p(cm), C.c_int(len(consts)), C.c_int(len(pidx)), _p(pidx),
                C.c_int(len(info.permutations)), _p(flat), C.c_int(getattr(info, "max_degree", 4)), C.byref(self.h))
        if isinstance(info.num_instances, int) and isinstance(info.num_witness_polys, int):
            _chk(lib().b200_hyperplonk_preprocess(ctx.h, C.c_int(k), C.c_int(info.num_instances), C.c_int(info.num_witness_polys),
                                                  *tail), "hyperplonk_preprocess")
        else:
            ni, nw, nc = (np.asarray(v if v else [0], dtype=np.int32) for v in
                          (self.instance_cols, self.phase_witness, self.phase_challenges))
            _chk(lib().b200_hyperplonk_preprocess_phased(ctx.h, C.c_int(k), C.c_int(len(self.instance_cols)), _p(ni),
                                                         C.c_int(len(self.phase_witness)), _p(nw), _p(nc), *tail),
                 "hyperplonk_preprocess_phased")
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

    def verifier(self, kzg_verifier):
        """The `vp` half of `HyperPlonk::preprocess -> (pp, vp)` (hyperplonk.rs:138-161) for the CPU verifier
        (verifier.py / include/b200_verify.h): same circuit shape, the commitments of this prover parameter object, the
        zero-check expression composed by the library (b200_expression_compose)."""
        from . import compose_native
        from .verifier import HyperPlonkVerifier

        info = self.info
        nz, tok, cm = compose_native(info.k, info.constraints, info.num_poly, info.permutation_polys,
                                     sum(self.phase_challenges), getattr(info, "max_degree", 4), info.lookups)
        assert nz == self.num_z
        pre, perm = self.commitments()
        return HyperPlonkVerifier(kzg_verifier, info.k, self.instance_cols, self.phase_witness, self.phase_challenges,
                                  len(info.lookups), nz, (tok, cm), pre, perm)

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

    def prove_phased(self, instance_cols, synthesize):
        """The phase loop of hyperplonk.rs:183-204 through `b200_hyperplonk_prove_phased`: `synthesize(round, challenges)`
        plays PlonkishCircuit::synthesize — canonical-int challenges in, that phase's witness columns (canonical ints, or
        device MultilinearPolynomials) out. `instance_cols`: one list per instance column."""
        ctx = self.ctx
        flat = [v for col in instance_cols for v in col]
        inst = ints_to_mont(ctx, flat) if flat else np.zeros((1, 4), dtype=np.uint64)
        keep, err = [], []

        @SYNTHESIZE_FN
        def cb(_user, rnd, chal_ptr, nchal, out):
            try:
                ch = np.zeros((max(1, nchal), 4), dtype=np.uint64)
                if nchal:
                    C.memmove(ch.ctypes.data, chal_ptr, nchal * 32)
                cols = synthesize(rnd, [mont_to_int(ch[i]) for i in range(nchal)])
                if len(cols) != self.phase_witness[rnd]:
                    return 1
                for i, col in enumerate(cols):
                    poly = col if isinstance(col, MultilinearPolynomial) else upload_ints(ctx, col)
                    keep.append(poly)  # alive until prove returns
                    out[i] = poly.dev.value if hasattr(poly.dev, "value") else poly.dev
                return 0
            except Exception as e:  # never unwind through the C frames
                err.append(e)
                return 1

        rc = lib().b200_hyperplonk_prove_phased(self.h, _p(np.ascontiguousarray(inst)), C.c_int(len(flat)), cb, None)
        if err:
            raise err[0]
        _chk(rc, "hyperplonk_prove_phased")
