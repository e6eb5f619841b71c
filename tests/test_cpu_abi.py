"""CPU: the C-ABI library builds, loads without a CUDA driver and exports every symbol include/densebox_b200.h
declares; argument validation that does not need a device; the product path refuses to run without CUDA."""
import ctypes
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def lib():
    from densebox_b200.build import build
    return ctypes.CDLL(build())


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "densebox_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dbx_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_version_and_error_strings(lib):
    assert lib.dbx_version() >= 100
    lib.dbx_error_string.restype = ctypes.c_char_p
    assert lib.dbx_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert lib.dbx_error_string(code).startswith(b"dbx:")


def test_workspace_query_is_host_only(lib):
    n = ctypes.c_size_t(0)
    for variant, lo in ((0, 3.0e9), (1, 3.3e9), (2, 3.5e9)):
        assert lib.dbx_net_workspace_bytes(variant, 32, 240, 240, 1, ctypes.byref(n)) == 0
        assert lo < n.value < 3 * lo, n.value
    small = ctypes.c_size_t(0)
    assert lib.dbx_net_workspace_bytes(0, 32, 240, 240, 0, ctypes.byref(small)) == 0
    assert small.value < n.value
    assert lib.dbx_net_workspace_bytes(0, 1, 100, 240, 1, ctypes.byref(n)) == -1   # H % 8
    assert lib.dbx_net_workspace_bytes(7, 1, 240, 240, 1, ctypes.byref(n)) == -1   # bad variant
    assert lib.dbx_net_workspace_bytes(1, 1, 48, 48, 1, ctypes.byref(n)) == -1     # refine branch needs >= 56 px


def test_null_arguments_are_rejected(lib):
    assert lib.dbx_net_forward(None, None, 0, 0, 0, None) == -1
    assert lib.dbx_conv_fprop(None, 1, 8, 8, 64, 64, 0, None, 1, 1, 0, 64, None, 0, None, 0, 0, 0, None, 64, 0, 0, 0,
                              None) == -1
    assert lib.dbx_count_positives(None, None, 4, None, None) == -1


def test_product_path_has_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import densebox_b200
    from oracle import densebox_oracle as O
    net = densebox_b200.DenseBox(O.seeded_vgg19(0))
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 240, 240))
    with pytest.raises(RuntimeError):
        densebox_b200.densebox_loss(torch.zeros(1, 1, 60, 60), torch.zeros(1, 4, 60, 60), [[1, 1, 9, 9]])
    with pytest.raises(Exception):
        densebox_b200.NetEngine("densebox", 1, 240, 240)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "densebox_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f
