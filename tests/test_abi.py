"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/b200_lasso.h declares,
and fails loudly (no CPU fallback) when no GPU is present."""
import ctypes as C
import os

import pytest

import halo2_lasso_b200 as hl


def test_library_exports_every_declared_symbol():
    syms = hl.declared_symbols()
    assert len(syms) >= 30
    lib = hl.lib()
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_library_exports_nothing_undeclared():
    """every exported b200_* symbol is declared (and documented) in include/b200_lasso.h"""
    import shutil
    import subprocess

    if not shutil.which("nm"):
        pytest.skip("nm not available")
    out = subprocess.run(["nm", "-D", "--defined-only", hl.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[2] for l in out.splitlines() if len(l.split()) == 3 and l.split()[1] == "T" and l.split()[2].startswith("b200_")}
    assert exported == set(hl.declared_symbols())


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hl.B200Error) as e:
        hl.Context(0)
    assert e.value.code in (hl.B200_ERR_CUDA, hl.B200_ERR_ARG)


def test_product_code_never_touches_the_oracle():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "halo2-lasso_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in src and "import oracle" not in src and "oracle/" not in src.replace("oracle/lasso.hpp", ""), f
