"""CPU suite: the C-ABI library builds/loads and exports every symbol include/iblnerf_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "iblnerf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ibln_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from ibl_nerf_b200 import build, _lib
    build.build_library()
    h = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(h, n), "missing export: " + n
    assert set(names) == set(_lib.exported_names()), set(names) ^ set(_lib.exported_names())
    hh = _lib.lib()
    assert hh.ibln_abi_version() == _lib.ABI_VERSION == 4
    assert hh.ibln_mlp_packed_bytes() == (98 + 92) * 16384 + 4 * 6296      # fwd chunks + consts + transposed (dgrad) chunks


def test_invalid_arguments_are_rejected_without_a_gpu():
    from ibl_nerf_b200 import _lib
    h = _lib.lib()
    assert h.ibln_stratified_z(None, None, None, 4, 64, 0, None, 0, None) == -1
    assert h.ibln_composite_fwd(None, None, None, None, 1, 64, 18, 3, 1, None, None, None, 0, None) == -1
    assert b"invalid" in h.ibln_error_string(-1)


def test_product_path_has_no_cpu_fallback():
    import torch
    import ibl_nerf_b200 as ib
    with pytest.raises(ib._lib.IblnError):
        ib.ops.stratified_z(torch.zeros(4), torch.ones(4), 8)
    with pytest.raises(ib._lib.IblnError):
        ib.get_embedder(10)[0](torch.zeros(5, 3))
