"""The C-ABI library loads and exports every symbol include/mandalorion_poa.h declares; without a
GPU the product path fails loudly (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mandalorion_poa.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpoa_[a-z_]+)\s*\(", text)))


def test_header_symbols_are_exported(built):
    from mandalorion_b200 import library_path
    from mandalorion_b200.poa import ABI_SYMBOLS
    lib = ctypes.CDLL(library_path())
    syms = declared_symbols()
    assert len(syms) >= 10
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert sorted(ABI_SYMBOLS) == syms
    lib.mpoa_abi_version.restype = ctypes.c_int
    assert lib.mpoa_abi_version() == 3


def test_default_params_are_the_reference_command_line(built):
    # abpoa -M 5 -r 0 (reference utils/SpliceDefineConsensus.py:917): M=5, X=4, O=4,24, E=2,1, b=10, f=0.01
    from mandalorion_b200 import library_path
    from mandalorion_b200.poa import _Params
    lib = ctypes.CDLL(library_path())
    p = _Params()
    lib.mpoa_default_params(ctypes.byref(p))
    assert (p.match, p.mismatch, p.gap_open1, p.gap_ext1, p.gap_open2, p.gap_ext2, p.wb) == (5, 4, 4, 2, 24, 1, 10)
    assert abs(p.wf - 0.01) < 1e-9 and (p.simd_pn_i16, p.simd_pn_i32) == (16, 8)


def test_struct_layouts_match_the_header(built):
    from mandalorion_b200.poa import _Params, _Stats, _Trace
    assert ctypes.sizeof(_Params) == 4 * 16
    assert ctypes.sizeof(_Stats) == 8 * 28
    assert ctypes.sizeof(_Trace) == 8 * 5


def test_no_cpu_fallback(built):
    import torch
    from mandalorion_b200 import PoaContext, PoaError
    if torch.cuda.is_available():
        pytest.skip("GPU present: the loud-failure path is for CPU-only boxes")
    with pytest.raises(PoaError):
        PoaContext(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mandalorion_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                for needle in ("import oracle", "from oracle", "mpoa_oracle", "liboracle", "oracle/"):
                    assert needle not in text, f"{f} refers to the oracle ({needle})"
