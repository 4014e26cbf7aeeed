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


def test_job_structs_match_the_header_layout():
    """ctypes mirrors of the job records the batched kernels read from device memory (sizes asserted in the .cu too)."""
    assert ctypes.sizeof(_lib.UnpackJob) == 40 and _lib.UnpackJob.blk_begin.offset == 32
    arr, blocks = _lib.fill_unpack_jobs([(0x1000, 0x2000, 64, 64, 9), (0x3000, 0x4000, 256, 2048, 9), (0x5000, 0x6000, 256, 304, 9)])
    assert [a.chunks for a in arr] == [1, 4, 1] and [a.blk_begin for a in arr] == [0, 64, 64 + 1024] and blocks == 64 + 1024 + 256
    parr, pblocks = _lib.fill_pack_jobs([(1, 2, 64, 64, 9, 64, 576, 0), (3, 4, 256, 2048, 9, 2048, 18432, 0)])
    assert parr[0].blk_begin == 0 and parr[1].blk_begin == parr[0].blk_count and pblocks == parr[1].blk_begin + parr[1].blk_count


def test_device_transform_draw_is_reproducible_and_in_range():
    import torch
    from iswm_b200.data import DeviceTransform
    tf = DeviceTransform(crop_size=(32, 48), hflip=True, generator=torch.Generator().manual_seed(3))
    a = tf.draw(16, 40, 64)
    tf.generator = torch.Generator().manual_seed(3)
    b = tf.draw(16, 40, 64)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    assert a[0].dtype == torch.int32 and a[1].dtype == torch.uint8
    assert int(a[0][:, 0].max()) <= 64 - 48 and int(a[0][:, 1].max()) <= 40 - 32 and int(a[0].min()) >= 0
    with pytest.raises(ValueError):
        tf.draw(2, 16, 16)
    assert DeviceTransform().draw(4, 8, 8) == (None, None)
