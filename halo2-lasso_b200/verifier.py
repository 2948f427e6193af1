"""CPU verifier binding (libb200verify.so, include/b200_verify.h): the verifying halves of the reference's traits —
`SumCheck::verify` (pb/piop/sum_check.rs:39-58), `MultilinearKzg::verify` / `batch_verify` (pb/pcs/multilinear/kzg.rs:
330-361, pb/pcs/multilinear.rs:237-275), `HyperPlonk::verify` (pb/backend/hyperplonk.rs:293-363) — and the verifier of
the Lasso argument. Host code only: it runs without a GPU. Field elements are (..., 4) uint64 Montgomery limbs, G1
points (..., 8) uint64, as everywhere in this package."""
import ctypes as C
import os

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_DIR, "libb200verify.so")
HEADER_PATH = os.path.join(os.path.dirname(_DIR), "include", "b200_verify.h")
ACCEPT, REJECT, ERR_ARG = 0, 1, 2
_lib = None


class VerifierArgError(ValueError):
    """B200V_ERR_ARG: the statement / parameters are malformed (not a property of the proof)"""


def lib():
    """Load libb200verify.so; fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(f"{LIB_PATH} missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        _lib = C.CDLL(LIB_PATH)
    return _lib


def declared_symbols():
    import re

    src = re.sub(r"/\*.*?\*/", "", open(HEADER_PATH).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(b200v_[a-z0-9_]+)\s*\(", src)))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _fr(a, n=None):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    assert n is None or a.shape[0] == n
    return a


def _ok(rc):
    if rc == ERR_ARG:
        raise VerifierArgError("invalid argument")
    return rc == ACCEPT


class ProofTranscript:
    """`Keccak256Transcript::from_proof` (pb/util/transcript.rs:113-123): the reading side of the Fiat-Shamir transcript"""

    def __init__(self, proof: bytes):
        self.h = C.c_void_p()
        buf = (C.c_uint8 * max(1, len(proof))).from_buffer_copy(proof if proof else b"\0")
        _ok(lib().b200v_transcript_new(buf, C.c_uint64(len(proof)), C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().b200v_transcript_free(self.h)
            self.h = None

    def common_field_elements(self, fes):
        fes = _fr(fes)
        _ok(lib().b200v_transcript_common_field_elements(self.h, _p(fes), C.c_int(fes.shape[0])))

    def read_field_elements(self, n):
        out = np.zeros((n, 4), dtype=np.uint64)
        if not _ok(lib().b200v_transcript_read_field_elements(self.h, _p(out), C.c_int(n))):
            raise ValueError("Invalid field element encoding in proof")
        return out

    def read_commitments(self, n):
        out = np.zeros((n, 8), dtype=np.uint64)
        if not _ok(lib().b200v_transcript_read_commitments(self.h, _p(out), C.c_int(n))):
            raise ValueError("Invalid elliptic curve point encoding in proof")
        return out

    def squeeze_challenges(self, n):
        out = np.zeros((n, 4), dtype=np.uint64)
        _ok(lib().b200v_transcript_squeeze_challenges(self.h, _p(out), C.c_int(n)))
        return out

    def done(self):
        return _ok(lib().b200v_transcript_done(self.h))


def sumcheck_verify(tr, num_vars, degree, claimed_sum, coefficients_form=False):
    """`ClassicSumCheck::verify` (classic.rs:242-263): (final claim, challenges) or None when a round is inconsistent"""
    fin, x = np.zeros(4, dtype=np.uint64), np.zeros((num_vars, 4), dtype=np.uint64)
    s = np.ascontiguousarray(claimed_sum, dtype=np.uint64).reshape(4)
    ok = _ok(lib().b200v_sumcheck_verify(tr.h, C.c_int(num_vars), C.c_int(degree), _p(s), C.c_int(int(coefficients_form)),
                                         _p(fin), _p(x)))
    return (fin, x) if ok else None


class MultilinearKzgVerifier:
    """`MultilinearKzgVerifierParam` + `verify` / `batch_verify` (kzg.rs:79-84, 330-361)"""

    def __init__(self, handle, num_vars):
        self.h, self.num_vars = handle, num_vars

    @classmethod
    def setup(cls, ss):
        """the verifier half of the seeded test setup (`MultilinearKzg.setup(ctx, ss)` is the prover half)"""
        ss = _fr(ss)
        h = C.c_void_p()
        _ok(lib().b200v_kzg_setup(_p(ss), C.c_int(ss.shape[0]), C.byref(h)))
        return cls(h, ss.shape[0])

    @classmethod
    def from_g2_powers(cls, ss_g2):
        a = np.ascontiguousarray(ss_g2, dtype=np.uint64).reshape(-1, 16)
        h = C.c_void_p()
        _ok(lib().b200v_kzg_import(_p(a), C.c_int(a.shape[0]), C.byref(h)))
        return cls(h, a.shape[0])

    def g2_powers(self):
        out = np.zeros((self.num_vars, 16), dtype=np.uint64)
        _ok(lib().b200v_kzg_export(self.h, _p(out)))
        return out

    def __del__(self):
        if getattr(self, "h", None):
            lib().b200v_kzg_free(self.h)
            self.h = None

    def verify(self, tr, comm, point, evaluation):
        point = _fr(point)
        comm = np.ascontiguousarray(comm, dtype=np.uint64).reshape(8)
        ev = np.ascontiguousarray(evaluation, dtype=np.uint64).reshape(4)
        return _ok(lib().b200v_kzg_verify(self.h, tr.h, _p(comm), _p(point), C.c_int(point.shape[0]), _p(ev)))

    def batch_verify(self, tr, comms, points, evals):
        """evals: list of (poly, point, value), mirroring `Evaluation` (pb/pcs.rs:132-155)"""
        comms = np.ascontiguousarray(np.stack([np.asarray(c, dtype=np.uint64).reshape(8) for c in comms]))
        pts = np.ascontiguousarray(np.stack([_fr(p) for p in points]))
        nv = pts.shape[1]
        ep = np.asarray([e[0] for e in evals], dtype=np.int32)
        ept = np.asarray([e[1] for e in evals], dtype=np.int32)
        ev = np.ascontiguousarray(np.stack([np.asarray(e[2], dtype=np.uint64).reshape(4) for e in evals]))
        return _ok(lib().b200v_kzg_batch_verify(self.h, tr.h, C.c_int(nv), _p(comms), C.c_int(comms.shape[0]), _p(pts),
                                                C.c_int(pts.shape[0]), _p(ep), _p(ept), _p(ev), C.c_int(len(evals))))

    def lasso_verify(self, tr, kind, chunks, mu, expect_a=None, expect_dims=None, want_commitments=False):
        """the proof of `LassoProver(ctx, kzg, kind, chunks).prove(...)` for 2^mu lookups. Without `expect_a` /
        `expect_dims` (commitments to the lookup outputs / to the c chunked operands) ACCEPT only means that SOME
        committed a decomposes into table entries: bind the proof to your statement, or take the proof's commitments
        (`want_commitments=True` -> (ok, [1 + 4c points])) and link them yourself."""
        ea = np.ascontiguousarray(np.asarray(expect_a, dtype=np.uint64).reshape(8)) if expect_a is not None else None
        ed = (np.ascontiguousarray(np.asarray(expect_dims, dtype=np.uint64).reshape(chunks, 8))
              if expect_dims is not None else None)
        out = np.zeros((1 + 4 * chunks, 8), dtype=np.uint64) if want_commitments else None
        ok = _ok(lib().b200v_lasso_verify_statement(self.h, tr.h, C.c_int(kind), C.c_int(chunks), C.c_int(mu),
                                                    _p(ea) if ea is not None else None, _p(ed) if ed is not None else None,
                                                    _p(out) if out is not None else None))
        return (ok, out) if want_commitments else ok


    def lasso_verify_table(self, tr, table, mu, expect_a=None, expect_dims=None, want_commitments=False):
        """the proof of `LassoProver(ctx, kzg, table=table).prove(...)`: `table` is a `LassoTable` (a table given as data,
        part of the statement). Statement binding as in `lasso_verify`."""
        c = table.chunks
        ea = np.ascontiguousarray(np.asarray(expect_a, dtype=np.uint64).reshape(8)) if expect_a is not None else None
        ed = np.ascontiguousarray(np.asarray(expect_dims, dtype=np.uint64).reshape(c, 8)) if expect_dims is not None else None
        out = np.zeros((1 + 4 * c, 8), dtype=np.uint64) if want_commitments else None
        vals = np.ascontiguousarray(table.values, dtype=np.uint32)
        ok = _ok(lib().b200v_lasso_verify_table(self.h, tr.h, C.c_int(c), C.c_int(table.num_operands),
                                                C.c_int(table.operand_bits), C.c_int(table.out_bits), _p(vals), C.c_int(mu),
                                                _p(ea) if ea is not None else None, _p(ed) if ed is not None else None,
                                                _p(out) if out is not None else None))
        return (ok, out) if want_commitments else ok


class HyperPlonkVerifier:
    """`HyperPlonkVerifierParam` + `HyperPlonk::verify` (hyperplonk.rs:58-74, 293-363). `expression` is the composed
    zero-check expression (expression.py::compose), the commitments are `HyperPlonk.commitments()` of the prover side."""

    def __init__(self, kzg, k, num_instances, num_witness_polys, num_challenges, num_lookups, num_permutation_z_polys,
                 expression, preprocess_comms, permutation_comms):
        from .expression import R_MOD, serialize_expression

        self.kzg = kzg
        cols = [num_instances] if isinstance(num_instances, int) else list(num_instances)
        phases = [num_witness_polys] if isinstance(num_witness_polys, int) else list(num_witness_polys)
        chals = [0] * len(phases) if num_challenges is None else list(num_challenges)
        if isinstance(expression, tuple):  # (tokens, Montgomery constants), e.g. from b200_expression_compose
            tok = np.ascontiguousarray(expression[0], dtype=np.int32)
            consts = np.ascontiguousarray(expression[1], dtype=np.uint64).reshape(-1, 4)
            cm = consts if len(consts) else np.zeros((1, 4), dtype=np.uint64)
        else:
            tokens, consts = serialize_expression(expression, [], [])
            R = 1 << 256
            cm = np.zeros((max(1, len(consts)), 4), dtype=np.uint64)
            for i, c in enumerate(consts):  # canonical ints -> Montgomery limbs
                v = c * R % R_MOD
                cm[i] = [(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)]
            tok = np.asarray(tokens, dtype=np.int32)
        a, b, c = (np.asarray(v if v else [0], dtype=np.int32) for v in (cols, phases, chals))
        pre = np.ascontiguousarray(np.asarray(preprocess_comms, dtype=np.uint64).reshape(-1, 8))
        perm = np.ascontiguousarray(np.asarray(permutation_comms, dtype=np.uint64).reshape(-1, 8))
        self.h = C.c_void_p()
        _ok(lib().b200v_hyperplonk_new(kzg.h, C.c_int(k), C.c_int(len(cols)), _p(a), C.c_int(len(phases)), _p(b), _p(c),
                                       C.c_int(num_lookups), C.c_int(num_permutation_z_polys), _p(tok), C.c_int(len(tok)),
                                       _p(cm), C.c_int(len(consts)), _p(pre) if pre.size else None, C.c_int(pre.shape[0]),
                                       _p(perm) if perm.size else None, C.c_int(perm.shape[0]), C.byref(self.h)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().b200v_hyperplonk_free(self.h)
            self.h = None

    def verify(self, tr, instances):
        """instances: all instance columns back to back, as Montgomery field elements"""
        inst = np.ascontiguousarray(instances, dtype=np.uint64).reshape(-1, 4)
        return _ok(lib().b200v_hyperplonk_verify(self.h, tr.h, _p(inst) if inst.size else None, C.c_int(inst.shape[0])))


def fractional_sum_check_verify(tr, num_vars, claimed_p, claimed_q):
    """`verify_fractional_sum_check(num_vars, claimed_p_0s, claimed_q_0s, transcript)` (fractional_sum_check.rs:192-265) on a
    `ProofTranscript`; claimed_*: list with None (read from the proof) or a field element (Some: absorbed). Returns None on
    REJECT, else (p_xs, q_xs, x, p_0s, q_0s) — the caller still checks p_xs / q_xs against its polynomials at x."""
    B = len(claimed_p)
    mask = 0
    cp, cq = np.zeros((B, 4), dtype=np.uint64), np.zeros((B, 4), dtype=np.uint64)
    for b in range(B):
        if claimed_p[b] is not None:
            mask |= 1 << b
            cp[b] = np.asarray(claimed_p[b], dtype=np.uint64).reshape(4)
        if claimed_q[b] is not None:
            mask |= 1 << (16 + b)
            cq[b] = np.asarray(claimed_q[b], dtype=np.uint64).reshape(4)
    z = lambda k: np.zeros((k, 4), dtype=np.uint64)  # noqa: E731
    p_xs, q_xs, x, p0, q0 = z(B), z(B), z(num_vars), z(B), z(B)
    ok = _ok(lib().b200v_fractional_sum_check_verify(tr.h, C.c_int(B), C.c_int(num_vars), C.c_uint32(mask), _p(cp), _p(cq),
                                                     _p(p_xs), _p(q_xs), _p(x), _p(p0), _p(q0)))
    return (p_xs, q_xs, x, p0, q0) if ok else None


class HyperPlonkLassoVerifier:
    """Verifier of `hyperplonk.HyperPlonkLasso` proofs: HyperPlonk::verify, then the Lasso verifier BOUND to the witness
    commitment the HyperPlonk section carries (the first phase's witness commitments open the proof: 64 bytes each, big-
    endian coordinates, transcript.rs:216-227), then nothing may be left over."""

    def __init__(self, hp_verifier, kind, chunks, lookup_witness):
        self.hpv, self.kind, self.chunks, self.lookup_witness = hp_verifier, kind, chunks, lookup_witness

    def witness_commitment(self, proof):
        off = 64 * self.lookup_witness
        x = int.from_bytes(proof[off:off + 32], "big")
        y = int.from_bytes(proof[off + 32:off + 64], "big")
        q_mod = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
        mont = [(v << 256) % q_mod for v in (x, y)]  # the ABI carries Montgomery limbs (b200_verify.h)
        return np.array([(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for v in mont for i in range(4)], dtype=np.uint64)

    def verify(self, proof, instances, k):
        tr = ProofTranscript(proof)
        if not self.hpv.verify(tr, instances):
            return False
        return bool(self.hpv.kzg.lasso_verify(tr, self.kind, self.chunks, k, expect_a=self.witness_commitment(proof)) and tr.done())
