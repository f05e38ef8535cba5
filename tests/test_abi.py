"""The C-ABI shared library loads on a GPU-less host and exports every symbol include/irp_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT
from irp_b200 import _lib

HEADER = os.path.join(ROOT, "include", "irp_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(irp_[a-z0-9_]+)\s*\(", text)))


def test_library_is_built_in_tree():
    assert os.path.exists(_lib.LIB_PATH), "run __graft_entry__.build() first"
    assert os.path.commonpath([_lib.LIB_PATH, ROOT]) == ROOT


def test_every_declared_symbol_is_exported_and_bound():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} declared in irp_b200.h but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes prototype in irp_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_abi_version_and_error_string():
    lib = _lib.load()
    assert lib.irp_abi_version() == 1
    assert isinstance(lib.irp_last_error(), bytes)


def test_host_only_calls_report_errors_without_a_gpu():
    lib = _lib.load()
    # argument validation happens before any CUDA call
    assert lib.irp_preprocess_workspace_bytes(0, 5) == 0
    assert lib.irp_preprocess_workspace_bytes(4, 5) > 0
    assert lib.irp_lof_workspace_bytes(1000, 50, 30) > 1000 * 30 * 12
    st = lib.irp_preprocess_geometry(0, 10, None, None, None, None, None)
    assert st == 1 and b"geometry" in lib.irp_last_error()
    cout = ctypes.c_int()
    assert lib.irp_resnet50_conv_shape(0, ctypes.byref(cout), None, None, None, None) == 0 and cout.value == 64
    assert lib.irp_resnet50_conv_shape(53, None, None, None, None, None) == 1


def test_conv_index_order_matches_torchvision():
    import torchvision
    from irp_b200.stage import conv_bn_pairs
    lib = _lib.load()
    m = torchvision.models.resnet50(weights=None)
    for i, (conv, _) in enumerate(conv_bn_pairs(m)):
        vals = [ctypes.c_int() for _ in range(5)]
        assert lib.irp_resnet50_conv_shape(i, *[ctypes.byref(v) for v in vals]) == 0
        cout, cin, kh, kw, stride = (v.value for v in vals)
        assert tuple(conv.weight.shape) == (cout, cin, kh, kw)
        assert conv.stride == (stride, stride)


def test_library_build_id_matches_the_sources_on_disk():
    """irp_build_id() is the hash `make` stamped at build time; _lib.source_build_id() recomputes it from csrc/ and
    include/ -- a stale .so (sources edited, library not rebuilt) must not pass silently."""
    lib = _lib.load()
    assert lib.irp_build_id().decode() == _lib.source_build_id() == _lib.assert_fresh()
