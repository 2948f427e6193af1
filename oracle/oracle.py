"""ORACLE (test infrastructure only): ctypes driver for oracle/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. Field elements are numpy uint64 arrays of shape (..., 4) holding Montgomery
limbs; G1 affine points are (..., 8) uint64 (x limbs then y limbs).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_DIR, "liboracle.so")

R_MOD = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
Q_MOD = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47


def build(force=False):
    srcs = [os.path.join(_DIR, f) for f in os.listdir(_DIR) if f.endswith((".hpp", ".cpp"))]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _DIR, "-s"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_tr_new.restype = C.c_void_p
        _lib.orc_tr_from_proof.restype = C.c_void_p
        _lib.orc_kzg_setup.restype = C.c_void_p
        _lib.orc_kzg_import.restype = C.c_void_p
        _lib.orc_hp_preprocess.restype = C.c_void_p
        _lib.orc_tr_proof_len.restype = C.c_uint64
        _lib.orc_rand_u64.restype = C.c_uint64
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _fr(n=None):
    return np.zeros((4,) if n is None else (n, 4), dtype=np.uint64)


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(C.c_int(n))


# ---- conversions ----------------------------------------------------------------------------
def ints_to_raw(vals):
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for k in range(4):
            out[i, k] = (v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF
    return out


def raw_to_ints(raw):
    raw = np.asarray(raw, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(raw[i, k]) << (64 * k) for k in range(4)) for i in range(raw.shape[0])]


def fr_from_ints(vals, field="fr"):
    raw = ints_to_raw([v % (R_MOD if field == "fr" else Q_MOD) for v in vals])
    out = np.zeros_like(raw)
    getattr(lib(), f"orc_{field}_from_raw")(_p(raw), _p(out), C.c_uint64(len(vals)))
    return out


def fr_to_ints(a, field="fr"):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    raw = np.zeros_like(a)
    getattr(lib(), f"orc_{field}_to_raw")(_p(a), _p(raw), C.c_uint64(a.shape[0]))
    return raw_to_ints(raw)


def field_op(op, a, b=None, field="fr"):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros_like(a)
    f = getattr(lib(), f"orc_{field}_{op}")
    if b is None:
        f(_p(a), _p(out), C.c_uint64(a.shape[0]))
    else:
        b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        f(_p(a), _p(b), _p(out), C.c_uint64(a.shape[0]))
    return out


def rand_fr(seed, n):
    out = _fr(n)
    lib().orc_rand_fr(C.c_uint64(seed), C.c_uint64(n), _p(out))
    return out


def rand_u64s(seed, n):
    out = np.zeros(n, dtype=np.uint64)
    lib().orc_rand_u64s(C.c_uint64(seed), C.c_uint64(n), _p(out))
    return out


def keccak256(data: bytes, pad=0x01) -> bytes:
    out = (C.c_uint8 * 32)()
    lib().orc_keccak256(data, C.c_uint64(len(data)), C.c_uint8(pad), out)
    return bytes(out)


# ---- transcript -----------------------------------------------------------------------------
class Transcript:
    def __init__(self, proof: bytes = None):
        L = lib()
        self.h = C.c_void_p(L.orc_tr_new() if proof is None else L.orc_tr_from_proof(proof, C.c_uint64(len(proof))))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tr_free(self.h)
            self.h = None

    def proof(self) -> bytes:
        n = lib().orc_tr_proof_len(self.h)
        buf = (C.c_uint8 * n)()
        lib().orc_tr_proof(self.h, buf)
        return bytes(buf)

    def common_fe(self, fe):
        fe = np.ascontiguousarray(fe, dtype=np.uint64)
        lib().orc_tr_common_fe(self.h, _p(fe))

    def write_fe(self, fe):
        fe = np.ascontiguousarray(fe, dtype=np.uint64)
        lib().orc_tr_write_fe(self.h, _p(fe))

    def read_fe(self):
        out = _fr()
        if lib().orc_tr_read_fe(self.h, _p(out)):
            raise ValueError("Invalid field element encoding in proof")
        return out

    def squeeze(self):
        out = _fr()
        lib().orc_tr_squeeze(self.h, _p(out))
        return out

    def squeeze_n(self, n):
        return np.stack([self.squeeze() for _ in range(n)]) if n else _fr(0)

    def write_comm(self, pt):
        pt = np.ascontiguousarray(pt, dtype=np.uint64)
        if lib().orc_tr_write_comm(self.h, _p(pt)):
            raise ValueError("Invalid elliptic curve point encoding")

    def read_comm(self):
        out = np.zeros(8, dtype=np.uint64)
        if lib().orc_tr_read_comm(self.h, _p(out)):
            raise ValueError("Invalid elliptic curve point encoding in proof")
        return out


# ---- G1 -------------------------------------------------------------------------------------
def g1_generator():
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_g1_generator(_p(out))
    return out


def g1_mul(p, k):
    out = np.zeros(8, dtype=np.uint64)
    p = np.ascontiguousarray(p, dtype=np.uint64)
    k = np.ascontiguousarray(k, dtype=np.uint64)
    lib().orc_g1_mul(_p(p), _p(k), _p(out))
    return out


def g1_add(a, b):
    out = np.zeros(8, dtype=np.uint64)
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    lib().orc_g1_add(_p(a), _p(b), _p(out))
    return out


def g1_on_curve(p):
    p = np.ascontiguousarray(p, dtype=np.uint64)
    return bool(lib().orc_g1_on_curve(_p(p)))


def g1_to_ints(p):
    return tuple(fr_to_ints(np.asarray(p).reshape(2, 4), field="fq"))


# ---- MLE ------------------------------------------------------------------------------------
def eq_xy(y):
    y = np.ascontiguousarray(y, dtype=np.uint64).reshape(-1, 4)
    out = _fr(1 << y.shape[0])
    lib().orc_eq_xy(_p(y), C.c_int(y.shape[0]), _p(out))
    return out


def evaluate(p, x):
    p = np.ascontiguousarray(p, dtype=np.uint64)
    x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4)
    out = _fr()
    lib().orc_evaluate(_p(p), C.c_int(x.shape[0]), _p(x), _p(out))
    return out


def fix_var(p, r):
    p = np.ascontiguousarray(p, dtype=np.uint64)
    nv = int(p.shape[0]).bit_length() - 1
    out = _fr(p.shape[0] // 2)
    r = np.ascontiguousarray(r, dtype=np.uint64)
    lib().orc_fix_var(_p(p), C.c_int(nv), _p(r), _p(out))
    return out


def eq_xy_eval(x, y):
    x = np.ascontiguousarray(x, dtype=np.uint64).reshape(-1, 4)
    y = np.ascontiguousarray(y, dtype=np.uint64).reshape(-1, 4)
    out = _fr()
    lib().orc_eq_xy_eval(_p(x), _p(y), C.c_int(x.shape[0]), _p(out))
    return out


# ---- sum-check ------------------------------------------------------------------------------
def _ptr_array(arrs):
    arrs = [np.ascontiguousarray(a, dtype=np.uint64) for a in arrs]
    ptrs = (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])
    return arrs, ptrs


def sumcheck_prove_evals(tr, num_vars, polys, y, terms, claimed_sum):
    """terms: list of (coeff Fr, [poly indices]); y: eq point or None."""
    arrs, ptrs = _ptr_array(polys)
    coeffs = np.ascontiguousarray(np.stack([t[0] for t in terms]), dtype=np.uint64)
    off = np.zeros(len(terms) + 1, dtype=np.int32)
    idx = []
    for k, t in enumerate(terms):
        idx += list(t[1])
        off[k + 1] = len(idx)
    idx = np.asarray(idx, dtype=np.int32)
    ch, ev = _fr(num_vars), _fr(len(polys))
    yy = np.ascontiguousarray(y, dtype=np.uint64) if y is not None else _fr(num_vars)
    s = np.ascontiguousarray(claimed_sum, dtype=np.uint64)
    lib().orc_sumcheck_prove_evals(tr.h, C.c_int(num_vars), C.c_int(len(polys)), ptrs, C.c_int(y is not None),
                                   _p(yy), C.c_int(len(terms)), _p(coeffs), _p(off), _p(idx), _p(s), _p(ch), _p(ev))
    return ch, ev


def sumcheck_prove_coeffs(tr, num_vars, polys, prods, claimed_sum):
    """prods: list of (scalar Fr, y point (num_vars,4), poly index)."""
    arrs, ptrs = _ptr_array(polys)
    scalars = np.ascontiguousarray(np.stack([p[0] for p in prods]), dtype=np.uint64)
    ys = np.ascontiguousarray(np.stack([p[1] for p in prods]), dtype=np.uint64)
    pidx = np.asarray([p[2] for p in prods], dtype=np.int32)
    ch, ev = _fr(num_vars), _fr(len(polys))
    s = np.ascontiguousarray(claimed_sum, dtype=np.uint64)
    lib().orc_sumcheck_prove_coeffs(tr.h, C.c_int(num_vars), C.c_int(len(polys)), ptrs, C.c_int(len(prods)),
                                    _p(scalars), _p(ys), _p(pidx), _p(s), _p(ch), _p(ev))
    return ch, ev


def sumcheck_verify(tr, num_vars, degree, claimed_sum, coeffs=False):
    fin, ch = _fr(), _fr(num_vars)
    s = np.ascontiguousarray(claimed_sum, dtype=np.uint64)
    rc = lib().orc_sumcheck_verify(tr.h, C.c_int(num_vars), C.c_int(degree), _p(s), C.c_int(int(coeffs)), _p(fin), _p(ch))
    if rc:
        raise ValueError("InvalidSumcheck")
    return fin, ch


def serialize_expression(expr, tokens=None, consts=None, finish=True):
    """Expression tree (halo2_lasso_b200.expression.Expression) -> (prefix int tokens, constants as Fr).
    With finish=False the tokens/constants are appended to the given lists (several expressions, one stream)."""
    tokens = [] if tokens is None else tokens
    consts = [] if consts is None else consts

    def const_idx(v):
        consts.append(v % R_MOD)
        return len(consts) - 1

    def walk(e):
        k = e[0]
        if k == "const":
            tokens.extend([0, const_idx(e[1])])
        elif k == "identity":
            tokens.append(1)
        elif k == "lagrange":
            tokens.extend([2, e[1]])
        elif k == "eq":
            tokens.extend([3, e[1]])
        elif k == "poly":
            tokens.extend([4, e[1], e[2]])
        elif k == "chal":
            tokens.extend([5, e[1]])
        elif k == "neg":
            tokens.append(6)
            walk(e[1])
        elif k in ("sum", "prod"):
            tokens.append(7 if k == "sum" else 8)
            walk(e[1])
            walk(e[2])
        elif k == "scaled":
            tokens.extend([9, const_idx(e[2])])
            walk(e[1])
        elif k == "dpow":
            tokens.extend([10, len(e[1])])
            for c in e[1]:
                walk(c)
            walk(e[2])
        else:
            raise ValueError(k)

    walk(expr.node if hasattr(expr, "node") else expr)
    if not finish:
        return tokens, consts
    return np.asarray(tokens, dtype=np.int32), fr_from_ints(consts if consts else [0])


def serialize_lookups(lookups):
    """[[(input, table), ...], ...] -> one token stream: per lookup [width, input_0, table_0, ...]"""
    tokens, consts = [], []
    for lookup in lookups:
        tokens.append(len(lookup))
        for inp, tab in lookup:
            serialize_expression(inp, tokens, consts, finish=False)
            serialize_expression(tab, tokens, consts, finish=False)
    return np.asarray(tokens if tokens else [0], dtype=np.int32), fr_from_ints(consts if consts else [0])


def expression_rows(num_vars, expr, polys, challenges=()):
    """Expression::evaluate on every hypercube row (prover.rs:96-117)."""
    tokens, consts = serialize_expression(expr)
    arrs, ptrs = _ptr_array(polys)
    ch = np.ascontiguousarray(challenges, dtype=np.uint64).reshape(-1, 4) if len(challenges) else _fr(1)
    out = _fr(1 << num_vars)
    lib().orc_expression_rows(C.c_int(num_vars), _p(tokens), _p(consts), C.c_int(len(polys)), ptrs, _p(ch),
                              C.c_int(len(challenges)), _p(out))
    return out


def lookup_m(input_poly, table_poly):
    """lookup_m_poly (prover.rs:143-192); None = Invalid lookup input"""
    a, b = np.ascontiguousarray(input_poly, dtype=np.uint64), np.ascontiguousarray(table_poly, dtype=np.uint64)
    out = _fr(a.shape[0])
    rc = lib().orc_lookup_m(C.c_int(a.shape[0].bit_length() - 1), _p(a), _p(b), _p(out))
    return out if rc == 0 else None


def lookup_h(input_poly, table_poly, m, gamma):
    a, b = np.ascontiguousarray(input_poly, dtype=np.uint64), np.ascontiguousarray(table_poly, dtype=np.uint64)
    m = np.ascontiguousarray(m, dtype=np.uint64)
    out = _fr(a.shape[0])
    lib().orc_lookup_h(C.c_int(a.shape[0].bit_length() - 1), _p(a), _p(b), _p(m),
                       _p(np.ascontiguousarray(gamma, dtype=np.uint64)), _p(out))
    return out


def sumcheck_prove_generic(tr, num_vars, expr, polys, challenges, ys, claimed_sum):
    tokens, consts = serialize_expression(expr)
    arrs, ptrs = _ptr_array(polys)
    ch_in = np.ascontiguousarray(challenges, dtype=np.uint64).reshape(-1, 4)
    ys = np.ascontiguousarray(np.stack(ys), dtype=np.uint64)
    ch, ev = _fr(num_vars), _fr(len(polys))
    deg = C.c_int()
    s = np.ascontiguousarray(claimed_sum, dtype=np.uint64)
    lib().orc_sumcheck_prove_generic(tr.h, C.c_int(num_vars), _p(tokens), _p(consts), C.c_int(len(polys)), ptrs,
                                     _p(ch_in) if ch_in.shape[0] else None, C.c_int(ch_in.shape[0]), _p(ys),
                                     C.c_int(ys.shape[0]), _p(s), _p(ch), _p(ev), C.byref(deg))
    return ch, ev, deg.value


def bh_rotate(num_vars, b, rotation):
    lib().orc_bh_rotate.restype = C.c_uint64
    return int(lib().orc_bh_rotate(C.c_int(num_vars), C.c_uint64(b), C.c_int(rotation)))


def bh_iter(num_vars):
    out = np.zeros(1 << num_vars, dtype=np.uint64)
    lib().orc_bh_iter(C.c_int(num_vars), _p(out))
    return out


def sum_eq_ab(y, a, b):
    y = np.ascontiguousarray(y, dtype=np.uint64)
    out = _fr()
    lib().orc_sum_eq_ab(C.c_int(y.shape[0]), _p(y), _p(np.ascontiguousarray(a)), _p(np.ascontiguousarray(b)), _p(out))
    return out


# ---- MSM / KZG ------------------------------------------------------------------------------
def msm(scalars, bases):
    scalars = np.ascontiguousarray(scalars, dtype=np.uint64)
    bases = np.ascontiguousarray(bases, dtype=np.uint64)
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_msm(_p(scalars), _p(bases), C.c_uint64(scalars.shape[0]), _p(out))
    return out


class Kzg:
    def __init__(self, ss):
        ss = np.ascontiguousarray(ss, dtype=np.uint64).reshape(-1, 4)
        self.num_vars = ss.shape[0]
        self.ss = ss
        self.h = C.c_void_p(lib().orc_kzg_setup(_p(ss), C.c_int(self.num_vars)))

    @classmethod
    def from_eqs(cls, ss, levels):
        """Wrap an SRS computed elsewhere (levels[k]: (2^k, 8) affine points)."""
        ss = np.ascontiguousarray(ss, dtype=np.uint64).reshape(-1, 4)
        flat = np.ascontiguousarray(np.concatenate([np.asarray(l, dtype=np.uint64).reshape(-1, 8) for l in levels]))
        self = cls.__new__(cls)
        self.num_vars, self.ss = ss.shape[0], ss
        self.h = C.c_void_p(lib().orc_kzg_import(_p(ss), C.c_int(ss.shape[0]), _p(flat)))
        return self

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_kzg_free(self.h)
            self.h = None

    def eqs(self, level):
        out = np.zeros((1 << level, 8), dtype=np.uint64)
        lib().orc_kzg_eqs(self.h, C.c_int(level), _p(out))
        return out

    def set_pairing_check(self, on=True):
        """verify openings with the pairing product (kzg.rs:330-361) instead of the trapdoor identity"""
        lib().orc_kzg_set_pairing_check(self.h, C.c_int(int(on)))

    def commit(self, poly):
        poly = np.ascontiguousarray(poly, dtype=np.uint64)
        nv = int(poly.shape[0]).bit_length() - 1
        out = np.zeros(8, dtype=np.uint64)
        lib().orc_kzg_commit(self.h, _p(poly), C.c_int(nv), _p(out))
        return out

    def open(self, tr, poly, point):
        poly = np.ascontiguousarray(poly, dtype=np.uint64)
        point = np.ascontiguousarray(point, dtype=np.uint64).reshape(-1, 4)
        ev = _fr()
        if lib().orc_kzg_open(self.h, tr.h, _p(poly), C.c_int(point.shape[0]), _p(point), _p(ev)):
            raise ValueError("identity quotient commitment")
        return ev

    def verify(self, tr, comm, point, ev):
        comm = np.ascontiguousarray(comm, dtype=np.uint64)
        point = np.ascontiguousarray(point, dtype=np.uint64).reshape(-1, 4)
        ev = np.ascontiguousarray(ev, dtype=np.uint64)
        return lib().orc_kzg_verify(self.h, tr.h, _p(comm), C.c_int(point.shape[0]), _p(point), _p(ev)) == 0

    @staticmethod
    def _batch_args(points, evals):
        points = np.ascontiguousarray(np.stack(points), dtype=np.uint64)
        ev_poly = np.asarray([e[0] for e in evals], dtype=np.int32)
        ev_point = np.asarray([e[1] for e in evals], dtype=np.int32)
        ev_val = np.ascontiguousarray(np.stack([e[2] for e in evals]), dtype=np.uint64)
        return points, ev_poly, ev_point, ev_val

    def batch_open(self, tr, polys, points, evals):
        """evals: list of (poly idx, point idx, value Fr)."""
        nv = np.asarray(points[0]).reshape(-1, 4).shape[0]
        arrs, ptrs = _ptr_array(polys)
        pts, ep, ept, ev = self._batch_args(points, evals)
        rc = lib().orc_kzg_batch_open(self.h, tr.h, C.c_int(nv), C.c_int(len(polys)), ptrs, C.c_int(len(points)),
                                      _p(pts), C.c_int(len(evals)), _p(ep), _p(ept), _p(ev))
        if rc:
            raise ValueError("batch_open failed")

    def batch_verify(self, tr, comms, points, evals):
        nv = np.asarray(points[0]).reshape(-1, 4).shape[0]
        comms = np.ascontiguousarray(np.stack(comms), dtype=np.uint64)
        pts, ep, ept, ev = self._batch_args(points, evals)
        return lib().orc_kzg_batch_verify(self.h, tr.h, C.c_int(nv), C.c_int(comms.shape[0]), _p(comms),
                                          C.c_int(len(points)), _p(pts), C.c_int(len(evals)), _p(ep), _p(ept), _p(ev)) == 0


# ---- HyperPlonk (with or without LogUp lookups) ----------------------------------------------------------------
class HyperPlonk:
    """Oracle `HyperPlonk<MultilinearKzg>`: preprocess at construction, then prove / verify."""

    def __init__(self, kzg, num_vars, expression, num_instances, num_witness, preprocess_polys, perm_idx, cycles, num_z=1,
                 lookups=(), num_challenges=None):
        """num_instances: int (one instance column) or a list (one entry per column); num_witness: int (one phase) or
        a list per phase, then num_challenges is the list of challenges squeezed after each phase (backend.rs:50-60)."""
        tokens, consts = serialize_expression(expression)
        ltok, lconsts = serialize_lookups(lookups)
        arrs, ptrs = _ptr_array(preprocess_polys)
        flat = []
        for cyc in cycles:
            flat.append(len(cyc))
            for (p, r) in cyc:
                flat += [p, r]
        flat = np.asarray(flat if flat else [0], dtype=np.int32)
        pidx = np.asarray(perm_idx, dtype=np.int32)
        self.kzg, self.num_vars, self.nperm = kzg, num_vars, len(perm_idx)
        cols = [num_instances] if isinstance(num_instances, int) else list(num_instances)
        phases = [num_witness] if isinstance(num_witness, int) else list(num_witness)
        chals = [0] * len(phases) if num_challenges is None else list(num_challenges)
        assert len(chals) == len(phases)
        self.phase_witness = phases
        self.h = C.c_void_p(lib().orc_hp_preprocess(kzg.h, C.c_int(num_vars), _p(tokens), _p(consts), C.c_int(cols[0] if cols else 0),
                                                    C.c_int(sum(phases)), C.c_int(len(preprocess_polys)), ptrs,
                                                    C.c_int(len(perm_idx)), _p(pidx), _p(flat), C.c_int(len(cycles)),
                                                    C.c_int(num_z), C.c_int(len(lookups)), _p(ltok), _p(lconsts)))
        if len(cols) != 1 or len(phases) != 1 or chals != [0]:
            a, b, c = (np.asarray(v if v else [0], dtype=np.int32) for v in (cols, phases, chals))
            assert lib().orc_hp_set_phases(self.h, C.c_int(len(cols)), _p(a), C.c_int(len(phases)), _p(b), _p(c)) == 0

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_hp_free(self.h)
            self.h = None

    def permutation_poly(self, i):
        out = _fr(1 << self.num_vars)
        lib().orc_hp_permutation_poly(self.h, C.c_int(i), _p(out))
        return out

    def prove(self, tr, instances, witness_polys):
        inst = np.ascontiguousarray(instances, dtype=np.uint64).reshape(-1, 4)
        arrs, ptrs = _ptr_array(witness_polys)
        return lib().orc_hp_prove(self.h, tr.h, _p(inst), C.c_int(inst.shape[0]), ptrs, C.c_int(len(witness_polys))) == 0

    def prove_phased(self, tr, instances, synthesize):
        """`synthesize(round, challenges) -> list of (2^k, 4) uint64 witness tables` plays PlonkishCircuit::synthesize
        (hyperplonk.rs:192-199); `instances`: all instance columns back to back."""
        inst = np.ascontiguousarray(instances, dtype=np.uint64).reshape(-1, 4)
        N = 1 << self.num_vars
        err = []

        @C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p))
        def cb(_user, rnd, chal_ptr, nchal, out):
            try:
                ch = np.zeros((nchal, 4), dtype=np.uint64)
                if nchal:
                    C.memmove(ch.ctypes.data, chal_ptr, nchal * 32)
                polys = synthesize(rnd, ch)
                if len(polys) != self.phase_witness[rnd]:
                    return 1
                for i, p in enumerate(polys):
                    a = np.ascontiguousarray(p, dtype=np.uint64).reshape(N, 4)
                    C.memmove(out[i], a.ctypes.data, N * 32)
                return 0
            except Exception as e:  # never unwind through the C frames
                err.append(e)
                return 1

        rc = lib().orc_hp_prove_phased(self.h, tr.h, _p(inst), C.c_int(inst.shape[0]), cb, None)
        if err:
            raise err[0]
        return rc == 0

    def verify(self, tr, instances):
        inst = np.ascontiguousarray(instances, dtype=np.uint64).reshape(-1, 4)
        return lib().orc_hp_verify(self.h, tr.h, _p(inst), C.c_int(inst.shape[0])) == 0


# ---- pairing -----------------------------------------------------------------------------------
def pairing_gen_multiples(a, b):
    """e(a * G1, b * G2) as 12 canonical Fq coefficients (Python ints)"""
    out = np.zeros(48, dtype=np.uint64)
    lib().orc_pairing_gen_multiples(_p(np.ascontiguousarray(a, dtype=np.uint64)), _p(np.ascontiguousarray(b, dtype=np.uint64)), _p(out))
    return [sum(int(out[4 * i + j]) << (64 * j) for j in range(4)) for i in range(12)]


def pairing_gen_pow(e):
    """e(G1, G2)^e"""
    out = np.zeros(48, dtype=np.uint64)
    lib().orc_pairing_gen_pow(_p(np.ascontiguousarray(e, dtype=np.uint64)), _p(out))
    return [sum(int(out[4 * i + j]) << (64 * j) for j in range(4)) for i in range(12)]


def g2_checks(k):
    return lib().orc_g2_checks(_p(np.ascontiguousarray(k, dtype=np.uint64)))


def pairing_product_is_identity(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    return lib().orc_pairing_product_is_identity(_p(a), _p(b), C.c_int(a.shape[0])) == 1


def permutation_z(perm_polys, wires, beta, gamma):
    a1, p1 = _ptr_array(perm_polys)
    a2, p2 = _ptr_array(wires)
    n = a1[0].shape[0]
    out = _fr(n)
    lib().orc_permutation_z(C.c_int(n.bit_length() - 1), C.c_int(len(perm_polys)), p1, p2,
                            _p(np.ascontiguousarray(beta, dtype=np.uint64)), _p(np.ascontiguousarray(gamma, dtype=np.uint64)), _p(out))
    return out


# ---- Lasso ----------------------------------------------------------------------------------
TABLE_RANGE, TABLE_AND, TABLE_XOR = 0, 1, 2


def _u64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.uint64)


def lasso_prove(kzg, tr, kind, chunks, mu, xs, ys=None):
    xs, ys = _u64(xs), _u64(ys)
    return lib().orc_lasso_prove(kzg.h, tr.h, C.c_int(kind), C.c_int(chunks), C.c_int(mu), _p(xs),
                                 _p(ys) if ys is not None else None) == 0


class CustomTable:
    """A decomposable table given as data (TABLE_CUSTOM, oracle/lasso.hpp): 2^16 subtable values, 1 or 2 operands of
    `operand_bits` bits per chunk, g = Σ_t 2^(out_bits t) E_t."""

    def __init__(self, chunks, num_operands, operand_bits, out_bits, values):
        self.chunks, self.num_operands, self.operand_bits, self.out_bits = chunks, num_operands, operand_bits, out_bits
        self.values = np.ascontiguousarray(values, dtype=np.uint32)
        assert self.values.shape == (1 << 16,)


def lasso_prove_custom(kzg, tr, tab, mu, xs, ys=None):
    """0 = proof written, 1 = identity commitment, 2 = invalid descriptor / operand outside the table"""
    xs, ys = _u64(xs), _u64(ys)
    return lib().orc_lasso_prove_custom(kzg.h, tr.h, C.c_int(tab.chunks), C.c_int(tab.num_operands), C.c_int(tab.operand_bits),
                                        C.c_int(tab.out_bits), _p(tab.values), C.c_int(mu), _p(xs),
                                        _p(ys) if ys is not None else None)


def lasso_verify_custom(kzg, tr, tab, mu):
    return lib().orc_lasso_verify_custom(kzg.h, tr.h, C.c_int(tab.chunks), C.c_int(tab.num_operands),
                                         C.c_int(tab.operand_bits), C.c_int(tab.out_bits), _p(tab.values), C.c_int(mu)) == 0


def lasso_verify(kzg, tr, kind, chunks, mu):
    return lib().orc_lasso_verify(kzg.h, tr.h, C.c_int(kind), C.c_int(chunks), C.c_int(mu)) == 0


def lasso_witness(kind, chunks, mu, xs, ys=None):
    xs, ys = _u64(xs), _u64(ys)
    mt = np.zeros((1 + 3 * chunks, 1 << mu, 4), dtype=np.uint64)
    st = np.zeros((chunks, 1 << 16, 4), dtype=np.uint64)
    lib().orc_lasso_witness(C.c_int(kind), C.c_int(chunks), C.c_int(mu), _p(xs), _p(ys) if ys is not None else None,
                            _p(mt), _p(st))
    return mt, st


def grand_product_prove(tr, leaves):
    arrs, ptrs = _ptr_array(leaves)
    T = len(leaves)
    h = int(arrs[0].shape[0]).bit_length() - 1
    claims, point = _fr(T), _fr(h)
    lib().orc_grand_product_prove(tr.h, C.c_int(T), C.c_int(h), ptrs, _p(claims), _p(point))
    return claims, point


def _claimed_mask(claimed_p, claimed_q):
    mask = 0
    for b, v in enumerate(claimed_p):
        mask |= (v is not None) << b
    for b, v in enumerate(claimed_q):
        mask |= (v is not None) << (16 + b)
    return mask


def fractional_sum_check_prove(tr, ps, qs, claimed_p=None, claimed_q=None):
    """prove_fractional_sum_check (pb/piop/gkr/fractional_sum_check.rs:87-190). claimed_*: per batch element None (the
    layer-0 value is written) or anything else (it is absorbed). Returns (p_xs, q_xs, x, p_0s, q_0s)."""
    pa, pptr = _ptr_array(ps)
    qa, qptr = _ptr_array(qs)
    B = len(ps)
    n = int(pa[0].shape[0]).bit_length() - 1
    claimed_p = [None] * B if claimed_p is None else claimed_p
    claimed_q = [None] * B if claimed_q is None else claimed_q
    p_xs, q_xs, x, p0, q0 = _fr(B), _fr(B), _fr(n), _fr(B), _fr(B)
    lib().orc_fractional_prove(tr.h, C.c_int(B), C.c_int(n), pptr, qptr, C.c_uint32(_claimed_mask(claimed_p, claimed_q)),
                               _p(p_xs), _p(q_xs), _p(x), _p(p0), _p(q0))
    return p_xs, q_xs, x, p0, q0


def fractional_sum_check_verify(tr, num_vars, claimed_p, claimed_q):
    """verify_fractional_sum_check (:192-265); claimed_*: list of None / Fr. Returns None on reject, else the tuple of
    fractional_sum_check_prove."""
    B = len(claimed_p)
    cp, cq = _fr(B), _fr(B)
    for b in range(B):
        if claimed_p[b] is not None:
            cp[b] = claimed_p[b]
        if claimed_q[b] is not None:
            cq[b] = claimed_q[b]
    p_xs, q_xs, x, p0, q0 = _fr(B), _fr(B), _fr(num_vars), _fr(B), _fr(B)
    rc = lib().orc_fractional_verify(tr.h, C.c_int(B), C.c_int(num_vars), C.c_uint32(_claimed_mask(claimed_p, claimed_q)),
                                     _p(cp), _p(cq), _p(p_xs), _p(q_xs), _p(x), _p(p0), _p(q0))
    return None if rc else (p_xs, q_xs, x, p0, q0)
