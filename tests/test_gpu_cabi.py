"""The C ABI used from a plain C99 host program (examples/cabi_demo.c: no Python, no torch, no C++ on the caller's
side — what a cgo / Rust-FFI binding links against): built with gcc on the box, run, and its proof files compared
byte for byte with the oracle on the same splitmix64 inputs."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_host_program_matches_oracle(tmp_path):
    lib_dir = os.path.join(ROOT, "halo2-lasso_b200")
    exe = str(tmp_path / "cabi_demo")
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "cabi_demo.c"), "-L", lib_dir, "-lb200lasso",
                           f"-Wl,-rpath,{lib_dir}", "-o", exe])
    mu = 8
    out = subprocess.run([exe, str(tmp_path), str(mu)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "kernel launches" in out.stdout

    n = 12
    a, b, y = O.rand_fr(1, 1 << n), O.rand_fr(2, 1 << n), O.rand_fr(3, n)
    one = O.fr_from_ints([1])[0]
    to = O.Transcript()
    O.sumcheck_prove_evals(to, n, [a, b], y, [(one, [0, 1])], one)
    assert (tmp_path / "sumcheck.bin").read_bytes() == to.proof()

    okzg = O.Kzg(O.rand_fr(7, 16))
    xs = O.rand_u64s(5, 1 << mu)
    xs[(1 << mu) // 2:] = xs[: (1 << mu) // 2]
    to = O.Transcript()
    assert O.lasso_prove(okzg, to, O.TABLE_RANGE, 4, mu, xs, None)
    proof = (tmp_path / "lasso.bin").read_bytes()
    assert proof == to.proof()
    assert O.lasso_verify(okzg, O.Transcript(proof), O.TABLE_RANGE, 4, mu)
