"""halo2-lasso_b200 — host-side mirror of the reference's prover interfaces over libb200lasso.so.

The reference is Rust and no Rust toolchain exists in this image, so the host layer above the C ABI
(include/b200_lasso.h) is this thin ctypes binding whose class / method names follow the reference:

    Keccak256Transcript      pb/util/transcript.rs   (write_field_elements, squeeze_challenges, into_proof)
    MultilinearPolynomial    pb/poly/multilinear.rs  (eq_xy, fix_var, evaluate)
    ClassicSumCheck          pb/piop/sum_check/classic.rs (prove; EvaluationsProver / CoefficientsProver)

Field elements are numpy uint64 arrays (..., 4): Montgomery limbs, the halo2curves memory layout.
There is NO CPU fallback: importing works anywhere (symbol checks), but every compute call needs the
CUDA library and a GPU and raises otherwise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libb200lasso.so")
HEADER_PATH = os.path.join(os.path.dirname(_DIR), "include", "b200_lasso.h")

B200_OK, B200_ERR_CUDA, B200_ERR_ARG, B200_ERR_TRANSCRIPT, B200_ERR_NOMEM, B200_ERR_LOOKUP, B200_ERR_PEER = range(7)


class B200Error(RuntimeError):
    """Maps the C status codes onto the reference's `Error` variants (pb/lib.rs:12-20)."""

    NAMES = {1: "Cuda", 2: "InvalidPcsParam/InvalidSumcheck (bad argument)", 3: "Transcript", 4: "OutOfMemory",
             5: "InvalidSnark (Invalid lookup input)", 6: "peer wait timed out (a rank left the collective)"}

    def __init__(self, code, where):
        super().__init__(f"{where}: {self.NAMES.get(code, code)}")
        self.code = code


def build(force=False):
    """Compile the sm_100a library in-tree (nvcc cross-compiles without a GPU)."""
    if force:
        subprocess.check_call(["make", "-C", _DIR, "clean"])
    subprocess.check_call(["make", "-C", _DIR, "-j8", "-s"])
    return LIB_PATH


_lib = None


def lib():
    """Load libb200lasso.so; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(LIB_PATH)
        _lib.b200_launch_count.restype = C.c_uint64
        _lib.b200_stream.restype = C.c_void_p
    return _lib


def _chk(rc, where):
    if rc != 0:
        raise B200Error(rc, where)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _fr(a):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.shape[-1] == 4
    return a


class Context:
    """b200_ctx: one per process / GPU."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _chk(lib().b200_ctx_create(C.c_int(device), C.byref(self.h)), "ctx_create")

    def close(self):
        if getattr(self, "h", None) and self.h:
            lib().b200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        _chk(lib().b200_sync(self.h), "sync")

    def launch_count(self, reset=False):
        return int(lib().b200_launch_count(self.h, C.c_int(int(reset))))

    @property
    def stream(self):
        return lib().b200_stream(self.h)


def sumcheck_eq_factored(ctx, on=True):
    """Select the eq-factored (default) or the plain EVAL-shape round kernel; both produce the same bytes."""
    _chk(lib().b200_sumcheck_eq_factored(ctx.h, C.c_int(int(on))), "sumcheck_eq_factored")


class MultilinearPolynomial:
    """Device-resident `MultilinearPolynomial<Fr>` (pb/poly/multilinear.rs:20-24)."""

    def __init__(self, ctx, dev, num_vars, owned=True):
        self.ctx, self.dev, self.num_vars, self.owned = ctx, dev, num_vars, owned

    @classmethod
    def new(cls, ctx, evals):
        evals = _fr(evals)
        n = evals.shape[0]
        assert n & (n - 1) == 0 and n > 0
        dev = C.c_void_p()
        _chk(lib().b200_poly_upload(ctx.h, _p(evals), C.c_uint64(n), C.byref(dev)), "poly_upload")
        ctx.sync()  # the host array may be pageable and die before the copy is staged
        return cls(ctx, dev, n.bit_length() - 1)

    @classmethod
    def alloc(cls, ctx, num_vars):
        dev = C.c_void_p()
        _chk(lib().b200_poly_alloc(ctx.h, C.c_uint64(1 << num_vars), C.byref(dev)), "poly_alloc")
        return cls(ctx, dev, num_vars)

    @classmethod
    def eq_xy(cls, ctx, y):
        y = _fr(y)
        out = cls.alloc(ctx, y.shape[0])
        _chk(lib().b200_eq_xy(ctx.h, _p(y), C.c_int(y.shape[0]), out.dev), "eq_xy")
        return out

    def __len__(self):
        return 1 << self.num_vars

    def evals(self):
        out = np.zeros((len(self), 4), dtype=np.uint64)
        _chk(lib().b200_poly_download(self.ctx.h, self.dev, C.c_uint64(len(self)), _p(out)), "poly_download")
        return out

    def fix_var(self, r):
        out = MultilinearPolynomial.alloc(self.ctx, self.num_vars - 1)
        _chk(lib().b200_fix_var(self.ctx.h, self.dev, C.c_int(self.num_vars), _p(_fr(r)), out.dev), "fix_var")
        return out

    def evaluate(self, x):
        return evaluate_many(self.ctx, [self], x)[0]

    def free(self):
        if self.owned and self.dev:
            lib().b200_poly_free(self.ctx.h, self.dev)
            self.dev = None

    def __del__(self):
        try:
            if self.ctx.h:
                self.free()
        except Exception:
            pass


def evaluate_many(ctx, polys, x):
    x = _fr(x)
    ptrs = (C.c_void_p * len(polys))(*[p.dev for p in polys])
    out = np.zeros((len(polys), 4), dtype=np.uint64)
    _chk(lib().b200_evaluate(ctx.h, ptrs, C.c_int(len(polys)), C.c_int(x.shape[0]), _p(x), _p(out)), "evaluate")
    return out


_proof_buf = None


class Keccak256Transcript:
    """FiatShamirTranscript<Keccak256, Cursor<Vec<u8>>> living on the device inside the context."""

    def __init__(self, ctx):
        self.ctx = ctx
        _chk(lib().b200_transcript_reset(ctx.h), "transcript_reset")

    def common_field_elements(self, fes):
        fes = _fr(fes).reshape(-1, 4)
        _chk(lib().b200_transcript_common_field_elements(self.ctx.h, _p(fes), C.c_int(fes.shape[0])), "common_fe")

    def write_field_elements(self, fes):
        fes = _fr(fes).reshape(-1, 4)
        _chk(lib().b200_transcript_write_field_elements(self.ctx.h, _p(fes), C.c_int(fes.shape[0])), "write_fe")

    def squeeze_challenges(self, n):
        out = np.zeros((n, 4), dtype=np.uint64)
        _chk(lib().b200_transcript_squeeze_challenges(self.ctx.h, C.c_int(n), _p(out)), "squeeze")
        return out

    def squeeze_challenge(self):
        return self.squeeze_challenges(1)[0]

    def write_commitments(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, 8)
        _chk(lib().b200_transcript_write_commitments(self.ctx.h, _p(pts), C.c_int(pts.shape[0])), "write_commitments")

    def into_proof(self) -> bytes:
        global _proof_buf
        n = C.c_uint64()
        if _proof_buf is None:
            _proof_buf = np.empty(8 << 20, dtype=np.uint8)  # the device stream's capacity; reused across proofs
        _chk(lib().b200_transcript_proof(self.ctx.h, _p(_proof_buf), C.c_uint64(_proof_buf.size), C.byref(n)), "into_proof")
        return _proof_buf[: n.value].tobytes()


class ClassicSumCheck:
    """`SumCheck::prove` (pb/piop/sum_check.rs:39-58) for the expression shapes of the Lasso hot path."""

    @staticmethod
    def prove_evals(ctx, num_vars, polys, weights, y, claimed_sum, np_per_term=2):
        """EvaluationsProver:  eq(x,y) * Σ_t w_t Π_k P[t*np+k](x). Returns (challenges, evals)."""
        nterms = len(polys) // np_per_term
        ptrs = (C.c_void_p * len(polys))(*[p.dev for p in polys])
        ch = np.zeros((num_vars, 4), dtype=np.uint64)
        ev = np.zeros((len(polys), 4), dtype=np.uint64)
        _chk(lib().b200_sumcheck_prove_evals(ctx.h, C.c_int(num_vars), C.c_int(nterms), C.c_int(np_per_term), ptrs,
                                             _p(_fr(weights)), _p(_fr(y)), _p(_fr(claimed_sum)), _p(ch), _p(ev)),
             "sumcheck_prove_evals")
        return ch, ev

    @staticmethod
    def prove_evals_host(ctx, num_vars, host_tables, weights, y, claimed_sum, np_per_term=2):
        """Same with HOST tables (pinned or pageable); host<->device copies happen inside the call."""
        nterms = len(host_tables) // np_per_term
        ptrs = (C.c_void_p * len(host_tables))(*[int(t) if isinstance(t, int) else t.ctypes.data for t in host_tables])
        ch = np.zeros((num_vars, 4), dtype=np.uint64)
        ev = np.zeros((len(host_tables), 4), dtype=np.uint64)
        _chk(lib().b200_sumcheck_prove_evals_host(ctx.h, C.c_int(num_vars), C.c_int(nterms), C.c_int(np_per_term), ptrs,
                                                  _p(_fr(weights)), _p(_fr(y)), _p(_fr(claimed_sum)), _p(ch), _p(ev)),
             "sumcheck_prove_evals_host")
        return ch, ev

    @staticmethod
    def prove_coeffs(ctx, num_vars, polys, scalars, ys, claimed_sum):
        """CoefficientsProver:  Σ_k s_k eq(x, y_k) P_k(x). Returns (challenges, evals)."""
        K = len(polys)
        ptrs = (C.c_void_p * K)(*[p.dev for p in polys])
        ys = _fr(ys).reshape(K * num_vars, 4)
        ch = np.zeros((num_vars, 4), dtype=np.uint64)
        ev = np.zeros((K, 4), dtype=np.uint64)
        _chk(lib().b200_sumcheck_prove_coeffs(ctx.h, C.c_int(num_vars), C.c_int(K), ptrs, _p(_fr(scalars)), _p(ys),
                                              _p(_fr(claimed_sum)), _p(ch), _p(ev)), "sumcheck_prove_coeffs")
        return ch, ev


PHASE_NAMES = {1000: "witness", 1001: "commit", 1002: "primary_sumcheck", 1004: "grand_product_m", 1006: "grand_product_s",
               1007: "leaf_evals", 1008: "open_m", 1009: "open_s", 1100: "msm_sort", 1101: "msm_accumulate",
               1102: "msm_reduce"}


def profile(ctx, fn):
    """Run fn() with the library's CUDA-event markers on; returns [(tag, ms), ...] in launch order."""
    _chk(lib().b200_profile_enable(ctx.h, C.c_int(1)), "profile_enable")
    fn()
    cap = 1 << 16
    ms = (C.c_float * cap)()
    tags = (C.c_int * cap)()
    n = C.c_int()
    _chk(lib().b200_profile_read(ctx.h, ms, tags, C.c_int(cap), C.byref(n)), "profile_read")
    _chk(lib().b200_profile_enable(ctx.h, C.c_int(0)), "profile_enable")
    return [(tags[i], ms[i]) for i in range(min(n.value, cap))]


def profile_rounds(ctx, fn, num_vars=20, tables=3):
    """Time every sum-check round launch of `fn()` with CUDA events on the library stream and return the
    roofline entry of the dominant launch (round 1: the first fused bind+evaluate over full tables)."""
    _chk(lib().b200_profile_enable(ctx.h, C.c_int(1)), "profile_enable")
    fn()
    cap = 4096
    ms = (C.c_float * cap)()
    tags = (C.c_int * cap)()
    n = C.c_int()
    _chk(lib().b200_profile_read(ctx.h, ms, tags, C.c_int(cap), C.byref(n)), "profile_read")
    _chk(lib().b200_profile_enable(ctx.h, C.c_int(0)), "profile_enable")
    rounds = {}
    for i in range(min(n.value, cap)):
        rounds.setdefault(tags[i], []).append(ms[i])
    per_round = [sum(v) / len(v) for _, v in sorted(rounds.items())]
    if len(per_round) < 2:
        return None
    # round 1 reads `tables` tables of 2^n and writes them bound to 2^(n-1): the algorithmic bytes of that launch
    # (the reference algorithm binds the eq table like any other table). The eq-factored kernel that runs
    # this launch never binds an eq table: it moves tables - 1 tables plus one 2^(n-2)-entry suffix-eq table.
    algo = 32 * tables * ((1 << num_vars) + (1 << (num_vars - 1)))
    moved = 32 * ((tables - 1) * ((1 << num_vars) + (1 << (num_vars - 1))) + (1 << (num_vars - 2)))
    t = per_round[1]
    return {"kernel": "sc_eval_fact_kernel<2,true> (round 1: fused bind + eq-factored evaluate + Fiat-Shamir step)",
            "launch_ms": t, "algorithmic_bytes_per_launch": algo, "achieved": algo / (t * 1e-3) / 1e9,
            "kernel_bytes_per_launch": moved, "round_ms": [round(x, 5) for x in per_round]}


def declared_symbols():
    """Every function name include/b200_lasso.h declares (used by the CPU-side ABI test)."""
    import re

    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", src)))


def variable_base_msm(ctx, scalars, bases):
    """`variable_base_msm(scalars, bases)` (pb/util/arithmetic/msm.rs:84-87) with host inputs."""
    scalars = _fr(scalars).reshape(-1, 4)
    bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    assert scalars.shape[0] == bases.shape[0]
    out = np.zeros(8, dtype=np.uint64)
    _chk(lib().b200_variable_base_msm(ctx.h, _p(scalars), _p(bases), C.c_uint64(scalars.shape[0]), _p(out)), "msm")
    return out


class MultilinearKzg:
    """`MultilinearKzg<Bn256>` prover side (pb/pcs/multilinear/kzg.rs): the ProverParam (eqs levels) is
    uploaded once; commit / open / batch_open run on the device and append to the context transcript."""

    def __init__(self, ctx, eqs_levels):
        """eqs_levels[k]: (2^k, 8) uint64 affine points = MultilinearKzgProverParams::eqs[k]."""
        self.ctx = ctx
        self.num_vars = len(eqs_levels) - 1
        for k, lv in enumerate(eqs_levels):
            lv = np.ascontiguousarray(lv, dtype=np.uint64).reshape(-1, 8)
            assert lv.shape[0] == 1 << k
            _chk(lib().b200_kzg_srs_upload(ctx.h, C.c_int(k), _p(lv)), "srs_upload")

    @classmethod
    def setup(cls, ctx, ss):
        """`MultilinearKzg::setup` + `trim` on the device from the trapdoor scalars ss (kzg.rs:166-250)."""
        ss = _fr(ss).reshape(-1, 4)
        self = cls.__new__(cls)
        self.ctx, self.num_vars = ctx, ss.shape[0]
        _chk(lib().b200_kzg_setup(ctx.h, _p(ss), C.c_int(ss.shape[0])), "kzg_setup")
        return self

    def eqs(self, level):
        out = np.zeros((1 << level, 8), dtype=np.uint64)
        _chk(lib().b200_kzg_srs_download(self.ctx.h, C.c_int(level), _p(out)), "srs_download")
        return out

    def batch_commit(self, polys, write=False):
        ptrs = (C.c_void_p * len(polys))(*[p.dev for p in polys])
        nv = (C.c_int * len(polys))(*[p.num_vars for p in polys])
        out = np.zeros((len(polys), 8), dtype=np.uint64)
        _chk(lib().b200_kzg_batch_commit(self.ctx.h, ptrs, nv, C.c_int(len(polys)), C.c_int(int(write)), _p(out)),
             "batch_commit")
        return out

    def commit(self, poly):
        return self.batch_commit([poly])[0]

    def batch_commit_and_write(self, polys):
        return self.batch_commit(polys, write=True)

    def open(self, poly, point):
        point = _fr(point).reshape(-1, 4)
        _chk(lib().b200_kzg_open(self.ctx.h, poly.dev, C.c_int(point.shape[0]), _p(point)), "open")

    def batch_open(self, polys, points, evals):
        """evals: list of (poly idx, point idx, value) — `Evaluation` (pb/pcs.rs:132-155)."""
        points = _fr(np.stack(points))
        nv = points.shape[1]
        ptrs = (C.c_void_p * len(polys))(*[p.dev for p in polys])
        ep = (C.c_int * len(evals))(*[e[0] for e in evals])
        ept = (C.c_int * len(evals))(*[e[1] for e in evals])
        ev = _fr(np.stack([e[2] for e in evals]))
        _chk(lib().b200_kzg_batch_open(self.ctx.h, C.c_int(nv), ptrs, C.c_int(len(polys)), _p(points),
                                       C.c_int(points.shape[0]), ep, ept, _p(ev), C.c_int(len(evals))), "batch_open")


TABLE_RANGE, TABLE_AND, TABLE_XOR = 0, 1, 2


def fractional_sum_check_prove(ctx, ps, qs, claimed_p=None, claimed_q=None):
    """`prove_fractional_sum_check(claimed_p_0s, claimed_q_0s, ps, qs, transcript)`
    (pb/piop/gkr/fractional_sum_check.rs:87-190) on the context's transcript. ps / qs: MultilinearPolynomial lists;
    claimed_*: per element None (the layer-0 value is written) or anything else (Some: it is absorbed).
    Returns (p_xs, q_xs, x, p_0s, q_0s)."""
    B = len(ps)
    n = ps[0].num_vars
    assert len(qs) == B and all(t.num_vars == n for t in list(ps) + list(qs))
    mask = 0
    for b in range(B):
        mask |= (claimed_p is not None and claimed_p[b] is not None) << b
        mask |= (claimed_q is not None and claimed_q[b] is not None) << (16 + b)
    pp = (C.c_void_p * B)(*[t.dev for t in ps])
    qp = (C.c_void_p * B)(*[t.dev for t in qs])
    z = lambda k: np.zeros((k, 4), dtype=np.uint64)  # noqa: E731
    p_xs, q_xs, x, p0, q0 = z(B), z(B), z(n), z(B), z(B)
    _chk(lib().b200_fractional_sum_check_prove(ctx.h, C.c_int(B), C.c_int(n), pp, qp, C.c_uint32(mask), _p(p_xs), _p(q_xs),
                                               _p(x), _p(p0), _p(q0)), "fractional_sum_check_prove")
    return p_xs, q_xs, x, p0, q0


class _LassoTableC(C.Structure):
    _fields_ = [("chunks", C.c_int), ("num_operands", C.c_int), ("operand_bits", C.c_int), ("out_bits", C.c_int),
                ("subtable", C.c_void_p)]


class LassoTable:
    """A decomposable table given as DATA — the role of a `DecomposableTable` implementation in the Lasso frontend
    (chunk bits, subtable values, the combiner g): `chunks` chunks, each addressing ONE 2^16-entry subtable `values`;
    one operand (dim_t = chunk t of x) or two (dim_t = chunk t of x << operand_bits | chunk t of y);
    lookup output g(E) = Σ_t 2^(out_bits t) E_t. `subtable_fn(address) -> value` may be given instead of `values`."""

    def __init__(self, chunks, num_operands, operand_bits, out_bits, values=None, subtable_fn=None):
        if values is None:
            values = [subtable_fn(x) for x in range(1 << 16)]
        self.chunks, self.num_operands, self.operand_bits, self.out_bits = chunks, num_operands, operand_bits, out_bits
        self.values = np.ascontiguousarray(values, dtype=np.uint32)
        if self.values.shape != (1 << 16,):
            raise ValueError("a subtable has 2^16 entries")
        self._handles = {}

    def handle(self, ctx):
        """the table uploaded to `ctx`'s device (b200_lasso_table_create), cached per context"""
        if ctx not in self._handles:
            desc = _LassoTableC(self.chunks, self.num_operands, self.operand_bits, self.out_bits, self.values.ctypes.data)
            h = C.c_void_p()
            _chk(lib().b200_lasso_table_create(ctx.h, C.byref(desc), C.byref(h)), "lasso_table_create")
            self._handles[ctx] = h
        return self._handles[ctx]

    def free(self):
        for h in self._handles.values():
            lib().b200_lasso_table_free(h)
        self._handles = {}


class LassoProver:
    """Lasso / Surge lookup prover (north_star; DESIGN.md "Lasso protocol"). The table is either a built-in kind
    (range / and / xor) + number of 16-bit-addressed chunks, or `table=LassoTable(...)` — a table given as data."""

    def __init__(self, ctx, kzg, kind=None, chunks=None, table=None):
        self.ctx, self.kzg, self.kind, self.table = ctx, kzg, kind, table
        self.chunks = table.chunks if table is not None else chunks
        assert (table is None) != (kind is None), "give either a built-in kind or a table"

    def prove_dev(self, mu, dev_xs, dev_ys=None):
        """Operands already resident on the device (raw pointers to u64 arrays); fully asynchronous."""
        if self.table is not None:
            _chk(lib().b200_lasso_prove_table_dev(self.ctx.h, self.table.handle(self.ctx), C.c_int(mu), C.c_void_p(dev_xs),
                                                  C.c_void_p(dev_ys) if dev_ys else None), "lasso_prove_table_dev")
            return
        _chk(lib().b200_lasso_prove_dev(self.ctx.h, C.c_int(self.kind), C.c_int(self.chunks), C.c_int(mu),
                                        C.c_void_p(dev_xs), C.c_void_p(dev_ys) if dev_ys else None), "lasso_prove_dev")

    def prove(self, xs, ys=None):
        """Appends the whole proof for the 2^mu lookups to the context transcript."""
        xs = np.ascontiguousarray(xs, dtype=np.uint64)
        mu = int(xs.shape[0]).bit_length() - 1
        assert xs.shape[0] == 1 << mu
        if ys is not None:
            ys = np.ascontiguousarray(ys, dtype=np.uint64)
        if self.table is not None:
            _chk(lib().b200_lasso_prove_table(self.ctx.h, self.table.handle(self.ctx), C.c_int(mu), _p(xs),
                                              _p(ys) if ys is not None else None), "lasso_prove_table")
            return
        _chk(lib().b200_lasso_prove(self.ctx.h, C.c_int(self.kind), C.c_int(self.chunks), C.c_int(mu), _p(xs),
                                    _p(ys) if ys is not None else None), "lasso_prove")

    def witness(self, xs, ys=None):
        """(a | dim | E | read_ts) tables [1+3c, 2^mu, 4] and final_cts [c, 2^16, 4] as computed on the device."""
        xs = np.ascontiguousarray(xs, dtype=np.uint64)
        mu = int(xs.shape[0]).bit_length() - 1
        if ys is not None:
            ys = np.ascontiguousarray(ys, dtype=np.uint64)
        c = self.chunks
        mt, st = C.c_void_p(), C.c_void_p()
        _chk(lib().b200_poly_alloc(self.ctx.h, C.c_uint64((1 + 3 * c) << mu), C.byref(mt)), "alloc")
        _chk(lib().b200_poly_alloc(self.ctx.h, C.c_uint64(c << 16), C.byref(st)), "alloc")
        _chk(lib().b200_lasso_witness(self.ctx.h, C.c_int(self.kind), C.c_int(c), C.c_int(mu), _p(xs),
                                      _p(ys) if ys is not None else None, mt, st), "lasso_witness")
        m_out = np.zeros((1 + 3 * c, 1 << mu, 4), dtype=np.uint64)
        s_out = np.zeros((c, 1 << 16, 4), dtype=np.uint64)
        _chk(lib().b200_poly_download(self.ctx.h, mt, C.c_uint64((1 + 3 * c) << mu), _p(m_out)), "download")
        _chk(lib().b200_poly_download(self.ctx.h, st, C.c_uint64(c << 16), _p(s_out)), "download")
        lib().b200_poly_free(self.ctx.h, mt)
        lib().b200_poly_free(self.ctx.h, st)
        return m_out, s_out


# ---- multi-GPU (one process per GPU) ---------------------------------------------------------
def dist_init(ctx, rank=None, world=None):
    """Map every rank's mailbox and bulk arena over CUDA IPC / NVLink. Needs an initialised torch.distributed group
    (any backend: only 128-byte handles are exchanged)."""
    import torch.distributed as dist

    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    mine = (C.c_uint8 * 128)()
    _chk(lib().b200_dist_mailbox_handle(ctx.h, mine), "dist_mailbox_handle")
    handles = exchange_handles(bytes(mine), world)
    blob = b"".join(handles)
    _chk(lib().b200_dist_init(ctx.h, C.c_int(rank), C.c_int(world), blob), "dist_init")
    dist.barrier()
    return rank, world


def dist_init_local(ctxs):
    """The ranks of a group as several contexts of THIS process (same GPU or peer-accessible GPUs); every context must
    then be driven from its own host thread (`run_ranks`). This is how the sharded provers are tested on one GPU."""
    arr = (C.c_void_p * len(ctxs))(*[c.h for c in ctxs])
    _chk(lib().b200_dist_init_local(arr, C.c_int(len(ctxs))), "dist_init_local")


def run_ranks(ctxs, fn):
    """fn(rank, ctx) on one host thread per context (collective calls wait for each other on the device); returns the
    results in rank order and re-raises the first exception."""
    import threading

    out, err = [None] * len(ctxs), [None] * len(ctxs)

    def work(r):
        try:
            out[r] = fn(r, ctxs[r])
        except BaseException as e:  # noqa: BLE001 - re-raised below
            err[r] = e

    th = [threading.Thread(target=work, args=(r,)) for r in range(len(ctxs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    bad = [(r, e) for r, e in enumerate(err) if e is not None]
    if bad:
        # a rank that left early makes every other rank time out in its next collective: report the root cause (the
        # first error that is not a peer timeout) and name what every rank saw
        root = next((e for _, e in bad if not (isinstance(e, B200Error) and e.code == 6)), bad[0][1])
        summary = "; ".join(f"rank {r}: {e!r}" for r, e in bad)
        raise RuntimeError(f"run_ranks: {summary}") from root
    return out


def dist_check(ctx):
    """Raises B200Error(6) if an in-kernel wait for a peer timed out since the last check."""
    _chk(lib().b200_dist_check(ctx.h), "dist_check")


def dist_shard_lasso(ctx, k0=16):
    """Fully sharded Lasso prover: witness tables, fingerprints, product-tree layers >= k0 and all sum-checks / openings
    over them on the rank's 1/world slice (index window [k0 - log2 world, k0)); 0 switches it off."""
    _chk(lib().b200_dist_shard_lasso(ctx.h, C.c_int(int(k0))), "dist_shard_lasso")


def dist_shard_min_items(ctx, items):
    _chk(lib().b200_dist_shard_min_items(ctx.h, C.c_int(int(items))), "dist_shard_min_items")


def dist_shard_commits(ctx, on=True):
    """Split every commitment MSM of the KZG / Lasso provers by point range over the ranks (collective calls from then
    on: all ranks run the same prover on the same inputs and obtain the identical proof)."""
    _chk(lib().b200_dist_shard_commits(ctx.h, C.c_int(int(on))), "dist_shard_commits")


def dist_shard_sumchecks(ctx, min_vars=16):
    """Evaluate every sum-check of the Lasso prover with at least `min_vars` variables on the rank's 1/world slice of
    the hypercube (round partials exchanged over NVLink inside the round kernels); 0 switches it off. Collective, like
    dist_shard_commits: all ranks run the same prover on the same inputs and obtain the identical proof."""
    _chk(lib().b200_dist_shard_sumchecks(ctx.h, C.c_int(int(min_vars))), "dist_shard_sumchecks")


def exchange_handles(mine: bytes, world: int):
    """all-gather of fixed-size byte strings in rank order (host-side plumbing, testable with gloo)."""
    import torch.distributed as dist

    out = [None] * world
    dist.all_gather_object(out, mine)
    assert all(isinstance(h, (bytes, bytearray)) and len(h) == len(mine) for h in out)
    return [bytes(h) for h in out]


def shard_slice(n_total_vars: int, rank: int, world: int):
    """[lo, hi) of the hypercube slice owned by `rank` when sharding on the TOP log2(world) variables."""
    g = world.bit_length() - 1
    assert 1 << g == world and n_total_vars > g
    size = 1 << (n_total_vars - g)
    return rank * size, (rank + 1) * size


def shard_window_slice(evals, n_total_vars: int, window_pos: int, rank: int, world: int):
    """The rank's compact slice of a table sharded on the index bits [window_pos, window_pos + log2 world)."""
    g = world.bit_length() - 1
    assert 1 << g == world and 0 <= window_pos <= n_total_vars - g
    a = np.asarray(evals)
    return np.ascontiguousarray(a.reshape((1 << (n_total_vars - g - window_pos), world, 1 << window_pos) + a.shape[1:])[:, rank]
                                .reshape((1 << (n_total_vars - g),) + a.shape[1:]))


def sumcheck_prove_evals_sharded(ctx, num_vars_total, local_polys, weights, y, claimed_sum, np_per_term=2, window_pos=-1,
                                 sharded_rounds=-1):
    nterms = len(local_polys) // np_per_term
    ptrs = (C.c_void_p * len(local_polys))(*[p.dev for p in local_polys])
    ch = np.zeros((num_vars_total, 4), dtype=np.uint64)
    ev = np.zeros((len(local_polys), 4), dtype=np.uint64)
    _chk(lib().b200_sumcheck_prove_evals_windowed(ctx.h, C.c_int(num_vars_total), C.c_int(window_pos), C.c_int(sharded_rounds),
                                                  C.c_int(nterms), C.c_int(np_per_term), ptrs, _p(_fr(weights)), _p(_fr(y)),
                                                  _p(_fr(claimed_sum)), _p(ch), _p(ev)),
         "sumcheck_prove_evals_sharded")
    return ch, ev


def variable_base_msm_sharded(ctx, local_scalars, local_bases):
    local_scalars = _fr(local_scalars).reshape(-1, 4)
    local_bases = np.ascontiguousarray(local_bases, dtype=np.uint64).reshape(-1, 8)
    out = np.zeros(8, dtype=np.uint64)
    _chk(lib().b200_variable_base_msm_sharded(ctx.h, _p(local_scalars), _p(local_bases),
                                              C.c_uint64(local_scalars.shape[0]), _p(out)), "msm_sharded")
    return out


# ---- generic expression sum-check (pb/piop/sum_check/classic/eval.rs for any Expression) -----------
def _mont_consts(ctx, ints):
    """canonical ints -> Montgomery limbs, converted on the device (no host field arithmetic in the product)."""
    n = max(1, len(ints))
    k = 1 << (n - 1).bit_length()
    raw = np.zeros((k, 4), dtype=np.uint64)
    for i, v in enumerate(ints):
        for j in range(4):
            raw[i, j] = (v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    p = MultilinearPolynomial.new(ctx, raw)
    _chk(lib().b200_fr_convert(ctx.h, p.dev, p.dev, C.c_uint64(k), C.c_int(1)), "fr_convert")
    return p.evals()[:n]


def _leaf_tables(ctx, num_vars, leaves, polys, ys=()):
    """Every leaf of a compiled expression as a dense device table (rotated queries gathered through the
    BooleanHypercube map, eq_xy tables, the identity polynomial, one-hot Lagrange tables)."""
    from .expression import BooleanHypercube

    tables = []
    for leaf in leaves:
        kind = leaf[0]
        if kind == "poly":
            _, p, rot = leaf
            if rot == 0:
                tables.append(polys[p])
            else:
                t = MultilinearPolynomial.alloc(ctx, num_vars)
                _chk(lib().b200_poly_rotate(ctx.h, polys[p].dev, C.c_int(num_vars), C.c_int(rot), t.dev), "poly_rotate")
                tables.append(t)
        elif kind == "eq":
            tables.append(MultilinearPolynomial.eq_xy(ctx, ys[leaf[1]]))
        elif kind == "identity":
            t = MultilinearPolynomial.alloc(ctx, num_vars)
            _chk(lib().b200_poly_iota(ctx.h, C.c_int(num_vars), t.dev), "poly_iota")
            tables.append(t)
        else:  # lagrange(i): one-hot at the i-th row in BooleanHypercube order (classic.rs:44-55)
            idx = BooleanHypercube(num_vars).nth(leaf[1])
            t = MultilinearPolynomial.alloc(ctx, num_vars)
            _chk(lib().b200_poly_onehot(ctx.h, C.c_int(num_vars), C.c_uint64(idx), t.dev), "poly_onehot")
            tables.append(t)
    return tables


def expression_rows(ctx, num_vars, expression, polys, challenges=()):
    """`Expression::evaluate` on every hypercube row (prover.rs:96-117) -> device polynomial. challenges: canonical
    Python ints. `lookup_compressed_poly` is this applied to distribute_powers(columns, beta)."""
    from .expression import compile_expression

    leaves, consts, prog = compile_expression(expression, list(challenges))
    if not leaves:
        raise B200Error("expression_rows: the expression has no polynomial leaf")
    tables = _leaf_tables(ctx, num_vars, leaves, polys)
    cm = _mont_consts(ctx, consts)
    ops = np.asarray(prog, dtype=np.int32).reshape(-1, 4)
    ptrs = (C.c_void_p * len(tables))(*[t.dev for t in tables])
    out = MultilinearPolynomial.alloc(ctx, num_vars)
    _chk(lib().b200_expression_rows(ctx.h, C.c_int(num_vars), C.c_int(len(tables)), ptrs, C.c_int(len(consts)), _p(cm),
                                    C.c_int(ops.shape[0]), _p(ops), out.dev), "expression_rows")
    return out


def lookup_m_poly(ctx, num_vars, compressed_input, compressed_table):
    """`lookup_m_poly` (prover.rs:143-192); raises B200Error (code B200_ERR_LOOKUP) for an input missing from the table."""
    m = MultilinearPolynomial.alloc(ctx, num_vars)
    _chk(lib().b200_lookup_m(ctx.h, C.c_int(num_vars), compressed_input.dev, compressed_table.dev, m.dev), "lookup_m")
    return m


def lookup_h_poly(ctx, num_vars, compressed_input, compressed_table, m, gamma):
    """`lookup_h_poly` (prover.rs:206-250); gamma: Montgomery (4,) uint64"""
    h = MultilinearPolynomial.alloc(ctx, num_vars)
    _chk(lib().b200_lookup_h(ctx.h, C.c_int(num_vars), compressed_input.dev, compressed_table.dev, m.dev,
                             _p(_fr(gamma)), h.dev), "lookup_h")
    return h


def prove_expression_native(ctx, num_vars, expression, polys, challenges, ys, claimed_sum):
    """`ClassicSumCheck::<EvaluationsProver>::prove` for an arbitrary expression, compiled INSIDE the library
    (b200_sumcheck_prove_expression): the expression crosses the boundary as prefix tokens. challenges: canonical
    Python ints; ys: list of (num_vars, 4) Montgomery arrays. Returns (challenges, evals)."""
    from .expression import serialize_expression

    tokens, consts = serialize_expression(expression, [], [])
    tokens = np.asarray(tokens, dtype=np.int32)
    cm = _mont_consts(ctx, consts)
    ch_in = _mont_consts(ctx, list(challenges))
    ys_in = np.ascontiguousarray(np.concatenate([_fr(y).reshape(-1, 4) for y in ys])) if len(ys) else np.zeros((1, 4), dtype=np.uint64)
    ptrs = (C.c_void_p * len(polys))(*[p.dev for p in polys])
    ch = np.zeros((num_vars, 4), dtype=np.uint64)
    ev = np.zeros((len(polys), 4), dtype=np.uint64)
    _chk(lib().b200_sumcheck_prove_expression(ctx.h, C.c_int(num_vars), _p(tokens), C.c_int(len(tokens)), _p(cm),
                                              C.c_int(len(consts)), ptrs, C.c_int(len(polys)), _p(ch_in),
                                              C.c_int(len(challenges)), _p(ys_in), C.c_int(len(ys)), _p(_fr(claimed_sum)),
                                              _p(ch), _p(ev)), "sumcheck_prove_expression")
    return ch, ev


def compile_expression_native(expression, const_mont):
    """`b200_expression_compile` (host only, no GPU): -> (leaves [(kind, a, b)], consts (n, 4) uint64 Montgomery,
    const_chal [challenge index or -1], ops [(opcode, dst, a, b)], ntemps, degree). const_mont: the expression's
    constants (serialize_expression order) already in Montgomery form, (n, 4) uint64."""
    from .expression import serialize_expression

    tokens, consts = serialize_expression(expression, [], [])
    tokens = np.asarray(tokens, dtype=np.int32)
    cm = np.ascontiguousarray(const_mont, dtype=np.uint64).reshape(-1, 4) if len(consts) else np.zeros((1, 4), dtype=np.uint64)
    assert cm.shape[0] >= len(consts)
    cap = 4096
    leaves = np.zeros((cap, 3), dtype=np.int32)
    cout = np.zeros((cap, 4), dtype=np.uint64)
    cchal = np.zeros(cap, dtype=np.int32)
    ops = np.zeros((cap, 4), dtype=np.int32)
    nl, nc, no, nt, deg = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _chk(lib().b200_expression_compile(_p(tokens), C.c_int(len(tokens)), _p(cm), C.c_int(len(consts)), _p(leaves),
                                       C.c_int(cap), C.byref(nl), _p(cout), _p(cchal), C.c_int(cap), C.byref(nc), _p(ops),
                                       C.c_int(cap), C.byref(no), C.byref(nt), C.byref(deg)), "expression_compile")
    return ([tuple(int(v) for v in r) for r in leaves[: nl.value]], cout[: nc.value].copy(), [int(v) for v in cchal[: nc.value]],
            [tuple(int(v) for v in r) for r in ops[: no.value]], nt.value, deg.value)


def compose_native(k, constraints, num_poly, permutation_polys, num_challenges=0, max_degree=4, lookups=()):
    """`b200_expression_compose` (host only, no GPU): preprocessor.rs:25-60 inside the library. Constants cross as
    Montgomery limbs (converted here with Python ints). Returns (num_permutation_z_polys, tokens int32, consts (n, 4)
    uint64 Montgomery) — what `b200v_hyperplonk_new` and `b200_sumcheck_prove_expression` take."""
    from .expression import R_MOD, serialize_expression

    ctok, ltok, consts = [], [], []
    for c in constraints:
        serialize_expression(c, ctok, consts)
    for lookup in lookups:
        ltok.append(len(lookup))
        for a, t in lookup:
            serialize_expression(a, ltok, consts)
            serialize_expression(t, ltok, consts)
    cm = np.zeros((max(1, len(consts)), 4), dtype=np.uint64)
    for i, c in enumerate(consts):
        v = c * (1 << 256) % R_MOD
        cm[i] = [(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)]
    ctok = np.asarray(ctok, dtype=np.int32)
    ltok = np.asarray(ltok if ltok else [0], dtype=np.int32)
    pidx = np.asarray(list(permutation_polys) if len(permutation_polys) else [0], dtype=np.int32)
    cap = 1 << 16
    tout, cout = np.zeros(cap, dtype=np.int32), np.zeros((cap, 4), dtype=np.uint64)
    nt, nc, nz = C.c_int(), C.c_int(), C.c_int()
    _chk(lib().b200_expression_compose(C.c_int(k), C.c_int(num_poly), C.c_int(num_challenges), C.c_int(len(constraints)),
                                       _p(ctok), C.c_int(len(ctok)), C.c_int(len(lookups)), _p(ltok),
                                       C.c_int(len(ltok) if lookups else 0), _p(cm), C.c_int(len(consts)),
                                       C.c_int(len(permutation_polys)), _p(pidx), C.c_int(max_degree), _p(tout), C.c_int(cap),
                                       C.byref(nt), _p(cout), C.c_int(cap), C.byref(nc), C.byref(nz)), "expression_compose")
    return nz.value, tout[: nt.value].copy(), cout[: nc.value].copy()


def prove_expression(ctx, num_vars, expression, polys, challenges, ys, claimed_sum):
    """`ClassicSumCheck::<EvaluationsProver>::prove(num_vars, VirtualPolynomial::new(expression, polys, challenges,
    ys), sum, transcript)`. challenges: canonical Python ints; ys: list of (num_vars, 4) Montgomery arrays.
    Returns (challenges, evals) with evals = every input polynomial bound at the challenges."""
    from .expression import compile_expression

    leaves, consts, prog = compile_expression(expression, challenges)
    tables = _leaf_tables(ctx, num_vars, leaves, polys, ys)
    # polynomials that the expression never queries at rotation 0 are still bound and returned (classic.rs:143-149)
    extra = [p for p in range(len(polys)) if ("poly", p, 0) not in leaves]
    K = len(tables)
    if extra:
        # append them as tables the program never reads; slots shift by len(extra)
        shift = len(extra)
        prog = [(op, d + shift if d >= K else d, a + shift if a >= K else a, b + shift if b >= K else b) for op, d, a, b in prog]
        tables += [polys[p] for p in extra]
    cm = _mont_consts(ctx, consts)
    ops = np.asarray(prog, dtype=np.int32).reshape(-1, 4)
    ptrs = (C.c_void_p * len(tables))(*[t.dev for t in tables])
    ch = np.zeros((num_vars, 4), dtype=np.uint64)
    ev = np.zeros((len(tables), 4), dtype=np.uint64)
    _chk(lib().b200_sumcheck_prove_generic(ctx.h, C.c_int(num_vars), C.c_int(expression.degree()), C.c_int(len(tables)), ptrs,
                                           C.c_int(len(consts)), _p(cm), C.c_int(ops.shape[0]), _p(ops), _p(_fr(claimed_sum)),
                                           _p(ch), _p(ev)), "sumcheck_prove_generic")
    pos = {leaf: i for i, leaf in enumerate(leaves)}
    for j, p in enumerate(extra):
        pos[("poly", p, 0)] = K + j
    return ch, np.stack([ev[pos[("poly", p, 0)]] for p in range(len(polys))])
