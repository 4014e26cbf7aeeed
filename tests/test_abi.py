"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/iswm_b200.h declares, and the ctypes table matches the header."""
import ctypes
import os
import re

import pytest

from iswm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "iswm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(iswm_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_symbols():
    syms = header_symbols()
    assert "iswm_conv_igemm" in syms and "iswm_wce_fwd_bwd" in syms and len(syms) >= 30


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_ctypes_table_matches_header():
    syms = set(header_symbols())
    table = set(_lib.SIGNATURES)
    assert syms == table, f"header-only: {sorted(syms - table)}; table-only: {sorted(table - syms)}"


def test_argument_counts_match_header():
    src = open(os.path.join(ROOT, "include", "iswm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, args in re.findall(r"\b(iswm_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        args = args.strip()
        n = 0 if args in ("", "void") else len(args.split(","))
        assert n == len(_lib.SIGNATURES[name][1]), f"{name}: header has {n} args, ctypes table {len(_lib.SIGNATURES[name][1])}"


def test_version_and_error_text_callable_without_gpu():
    lib = _lib.lib()
    assert lib.iswm_version() == 100
    assert isinstance(lib.iswm_last_error(), bytes)


def test_ops_refuse_cpu_tensors():
    import torch
    from iswm_b200 import ops
    with pytest.raises(RuntimeError):
        ops.class_hist(torch.zeros(16, dtype=torch.int64), 2, out=torch.zeros(2, dtype=torch.int64))
