"""CPU check of the product's 8x32-bit limb arithmetic (csrc/ff32.cuh) compiled for the host with an
emulated carry flag: the exact mad.lo.cc/madc.hi.cc chain algorithm vs Python big ints."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
R = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
Q = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47


@pytest.fixture(scope="module")
def shim(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ff32") / "ff32_host.so")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "host_shim", "ff32_host.cpp")])
    return C.CDLL(so)


def pack(v):
    a = np.zeros((len(v), 8), dtype=np.uint32)
    for i, x in enumerate(v):
        for k in range(8):
            a[i, k] = (x >> (32 * k)) & 0xFFFFFFFF
    return a


def unpack(a):
    return [sum(int(a[i, k]) << (32 * k) for k in range(8)) for i in range(a.shape[0])]


def _p(x):
    return x.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("name,p", [("h32_fr", R), ("h32_fq", Q)])
def test_limb_arithmetic_matches_bigint(shim, name, p):
    rng = random.Random(3)
    Rm = (1 << 256) % p
    Ri = pow(Rm, -1, p)
    edge = [0, 1, p - 1, p - 2, Rm, (p - 1) // 2, 2 ** 253, (1 << 32) - 1, ((1 << 256) - 1) % p]
    a = [rng.randrange(p) for _ in range(20000)] + edge + edge
    b = [rng.randrange(p) for _ in range(20000)] + edge + edge[::-1]
    A, B = pack(a), pack(b)
    O = np.zeros_like(A)
    n = C.c_long(len(a))
    getattr(shim, name + "_mul")(_p(A), _p(B), _p(O), n)
    assert unpack(O) == [x * y * Ri % p for x, y in zip(a, b)]
    getattr(shim, name + "_add")(_p(A), _p(B), _p(O), n)
    assert unpack(O) == [(x + y) % p for x, y in zip(a, b)]
    getattr(shim, name + "_sub")(_p(A), _p(B), _p(O), n)
    assert unpack(O) == [(x - y) % p for x, y in zip(a, b)]
    k = 200
    getattr(shim, name + "_inv")(_p(A), _p(O), C.c_long(k))
    assert unpack(O[:k]) == [pow(x * Ri % p, -1, p) * Rm % p for x in a[:k]]
    # binary (Kaliski) inverse on the edge values too (0 -> 0), and against the Fermat ladder it replaced
    tail = a[-2 * len(edge):]
    T = pack(tail)
    Ot, Of = np.zeros_like(T), np.zeros_like(T)
    getattr(shim, name + "_inv")(_p(T), _p(Ot), C.c_long(len(tail)))
    getattr(shim, name + "_inv_fermat")(_p(T), _p(Of), C.c_long(len(tail)))
    assert unpack(Ot) == [0 if x == 0 else pow(x * Ri % p, -1, p) * Rm % p for x in tail]
    assert (Ot == Of).all()
    small = [(v * Rm) % p for v in (1, 2, 3, 4, 2 ** 200, p - 1)] + [1, 2, 4, 2 ** 31, 2 ** 32, 2 ** 253]
    Sm = pack(small)
    Os = np.zeros_like(Sm)
    getattr(shim, name + "_inv")(_p(Sm), _p(Os), C.c_long(len(small)))
    assert unpack(Os) == [pow(x * Ri % p, -1, p) * Rm % p for x in small]
    getattr(shim, name + "_to_canonical")(_p(A), _p(O), C.c_long(k))
    assert unpack(O[:k]) == [x * Ri % p for x in a[:k]]
    # from_canonical must reduce ANY 256-bit integer (transcript challenges are raw hashes)
    big = [rng.randrange(1 << 256) for _ in range(2000)] + [(1 << 256) - 1, p, p + 1, 2 * p, 5 * p]
    Bg = pack(big)
    Ob = np.zeros_like(Bg)
    getattr(shim, name + "_from_canonical")(_p(Bg), _p(Ob), C.c_long(len(big)))
    assert unpack(Ob) == [x * Rm % p for x in big]


def test_device_transcript_code_matches_oracle_on_host(shim):
    """csrc/transcript.cuh (Keccak-f, absorb/squeeze, BE stream) compiled for the host vs the oracle."""
    import oracle as O

    for n in (0, 1, 3, 4, 5, 9, 40):  # crosses the 136-byte rate boundary at different offsets
        fes = O.rand_fr(77 + n, max(n, 1))[:n]
        g5 = O.g1_mul(O.g1_generator(), O.fr_from_ints([5])[0])
        proof = np.zeros(32 * n + 64, dtype=np.uint8)
        ch = np.zeros((3, 4), dtype=np.uint64)
        fes_c = np.ascontiguousarray(fes) if n else np.zeros((1, 4), dtype=np.uint64)
        plen = shim.h32_transcript_run(_p(fes_c), C.c_int(n), _p(g5), _p(proof), C.c_int(proof.size), _p(ch))
        tr = O.Transcript()
        for i in range(n):
            tr.write_fe(fes[i])
        c0 = tr.squeeze()
        tr.common_fe(c0)
        tr.write_comm(g5)
        c1, c2 = tr.squeeze(), tr.squeeze()
        assert plen == len(tr.proof())
        assert bytes(proof[:plen]) == tr.proof()
        assert (ch == np.stack([c0, c1, c2])).all()


def test_g1_xyzz_formulas_match_oracle_on_host(shim):
    """csrc/g1.cuh (XYZZ add / mixed add / double / to_affine incl. the P+P, P-P, identity branches)."""
    import oracle as O

    g = O.g1_generator()
    base = [O.g1_mul(g, O.fr_from_ints([k])[0]) for k in (1, 2, 3, 7, 11)]
    cases = [
        ([0, 1, 2], [1, 1, 1]),            # mixed adds
        ([0, 0], [1, 1]),                  # mixed add hits doubling (P + P)
        ([0, 0], [1, -1]),                 # P - P = identity
        ([3, 4, 1], [5, -3, 1000003]),     # mul_small + full add
        ([1, 0, 0], [1, 1, 1]),            # 2G + G + G: full/mixed doubling via equal points
        ([2], [0]),                        # 0 * P = identity
        ([0, 1, 2, 3, 4], [-1, 2, -3, 4, -5]),
    ]
    for idx, ks in cases:
        pts = np.ascontiguousarray(np.stack([base[i] for i in idx]))
        kk = np.asarray(ks, dtype=np.int32)
        out = np.zeros(8, dtype=np.uint64)
        shim.h32_g1_lincomb(_p(pts), _p(kk), C.c_int(len(ks)), _p(out))
        # expected with the oracle: Σ k_i P_i via scalar field arithmetic
        exp = np.zeros(8, dtype=np.uint64)
        for i, k in zip(idx, ks):
            exp = O.g1_add(exp, O.g1_mul(base[i], O.fr_from_ints([k % O.R_MOD])[0]))
        assert (out == exp).all(), (idx, ks)
