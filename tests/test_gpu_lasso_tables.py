"""Lasso over tables given as DATA (b200_lasso_table / LassoTable — the DecomposableTable role): GPU proofs against the
CPU oracle byte for byte, against the committed fixture of the independent pure-Python model, through the product's own
CPU verifier, and the error paths (operand outside the table, invalid descriptor)."""
import json
import os

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
NV = 16
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hl():
    import halo2_lasso_b200 as m

    return m


@pytest.fixture(scope="module")
def env(hl):
    ctx = hl.Context(0)
    okzg = O.Kzg(O.rand_fr(7, NV))
    kzg = hl.MultilinearKzg(ctx, [okzg.eqs(k) for k in range(NV + 1)])
    yield ctx, okzg, kzg
    ctx.close()


def or8(x):
    return (x >> 8) | (x & 0xFF)


def lt8(x):  # 8|8-bit "less than" flag — a 1-bit output per chunk
    return int((x >> 8) < (x & 0xFF))


def popcount16(x):
    return bin(x).count("1")


def sbox12(x):  # one 12-bit operand per chunk, addresses >= 2^12 unused (but part of the table)
    return (x * 2654435761 >> 7) & 0xFFF if x < (1 << 12) else 0


TABLES = {  # name: (chunks, num_operands, operand_bits, out_bits, rule)
    "or8_c8": (8, 2, 8, 8, or8),
    "or8_overlap_c3": (3, 2, 8, 5, or8),
    "lt8_c4": (4, 2, 8, 1, lt8),
    "popcount16_c4": (4, 1, 16, 5, popcount16),
    "sbox12_c5": (5, 1, 12, 12, sbox12),
}


def operands(chunks, nops, bits, mu, seed):
    xs, ys = O.rand_u64s(seed, 1 << mu), O.rand_u64s(seed + 1, 1 << mu)
    tot = bits * chunks
    if tot < 64:
        xs &= np.uint64((1 << tot) - 1)
        ys &= np.uint64((1 << tot) - 1)
    xs[1::4] = xs[0::4]
    ys[1::4] = ys[0::4]
    return xs, (ys if nops == 2 else None)


@pytest.mark.parametrize("name,mu", [("or8_c8", 6), ("or8_overlap_c3", 9), ("lt8_c4", 5), ("popcount16_c4", 13),
                                     ("sbox12_c5", 10), ("or8_c8", 14)])
def test_table_proof_parity_and_verifies(hl, env, name, mu):
    ctx, okzg, kzg = env
    chunks, nops, bits, out_bits, rule = TABLES[name]
    tab = hl.LassoTable(chunks, nops, bits, out_bits, subtable_fn=rule)
    otab = O.CustomTable(chunks, nops, bits, out_bits, tab.values)
    xs, ys = operands(chunks, nops, bits, mu, 700 + mu)
    to = O.Transcript()
    assert O.lasso_prove_custom(okzg, to, otab, mu, xs, ys) == 0
    tr = hl.Keccak256Transcript(ctx)
    hl.LassoProver(ctx, kzg, table=tab).prove(xs, ys)
    proof = tr.into_proof()
    assert proof == to.proof(), f"{name}: GPU proof differs from the oracle"
    assert O.lasso_verify_custom(okzg, O.Transcript(proof), otab, mu)
    # the product's own CPU verifier (pairing form), bound to the same table
    from halo2_lasso_b200 import verifier as V

    vk = V.MultilinearKzgVerifier.setup(O.rand_fr(7, NV))
    vt = V.ProofTranscript(proof)
    assert vk.lasso_verify_table(vt, tab, mu) and vt.done()
    other = hl.LassoTable(chunks, nops, bits, out_bits, values=np.where(np.arange(1 << 16) == 77, tab.values + 1, tab.values))
    assert not vk.lasso_verify_table(V.ProofTranscript(proof), other, mu)
    tab.free()


def test_golden_bytes_of_the_python_model(hl, env):
    ctx, okzg, kzg = env
    gold = json.load(open(os.path.join(HERE, "golden", "lasso_golden.json")))
    case = [c for c in gold["cases"] if c["kind"] == 3][0]
    t = case["table"]
    tab = hl.LassoTable(case["chunks"], t["num_operands"], t["operand_bits"], t["out_bits"], subtable_fn=or8)
    xs = np.asarray([int(v) for v in case["xs"]], dtype=np.uint64)
    ys = np.asarray([int(v) for v in case["ys"]], dtype=np.uint64)
    tr = hl.Keccak256Transcript(ctx)
    hl.LassoProver(ctx, kzg, table=tab).prove(xs, ys)
    assert tr.into_proof().hex() == case["proof"]
    tab.free()


def test_operands_outside_the_table_and_bad_descriptors(hl, env):
    ctx, okzg, kzg = env
    tab = hl.LassoTable(3, 2, 8, 8, subtable_fn=or8)
    xs, ys = operands(3, 2, 8, 6, 55)
    bad = xs.copy()
    bad[5] = np.uint64(1 << 24)  # 3 x 8 bits
    hl.Keccak256Transcript(ctx)
    with pytest.raises(hl.B200Error) as e:
        hl.LassoProver(ctx, kzg, table=tab).prove(bad, ys)
    assert e.value.code == 5  # Invalid lookup input
    with pytest.raises(hl.B200Error):
        hl.LassoProver(ctx, kzg, table=tab).prove(xs, None)  # a two-operand table needs ys
    tab.free()
    for args in ((1, 2, 8, 8), (3, 2, 9, 8), (3, 3, 4, 8), (8, 1, 16, 8), (8, 2, 8, 9)):  # c < 2, 18 address bits, 3 operands,
        with pytest.raises(hl.B200Error):                                                 # 128 operand bits, output > 64 bits
            hl.LassoTable(*args, subtable_fn=or8).handle(ctx)
