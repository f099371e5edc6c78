"""CPU checks of the C-ABI boundary: the library loads, exports every symbol include/a3gc_b200.h
declares, and rejects bad arguments with a status + message (no compute calls without a GPU)."""
import ctypes as C
import os
import re

import pytest

from a3gc_ip_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "a3gc_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(a3gc_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = header_functions()
    assert len(names) >= 12
    l = C.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/a3gc_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == names, "ctypes table out of sync with the header"


def test_abi_version_and_error_string():
    L = _lib.lib()
    assert L.a3gc_abi_version() == _lib.ABI_VERSION == 3
    assert isinstance(L.a3gc_last_error(), bytes)


def test_invalid_arguments_are_rejected_without_compute():
    L = _lib.lib()
    assert L.a3gc_gc_forward(None, None, None, 4, 12, 8, 0, None) == -1
    assert b"invalid" in L.a3gc_last_error()
    p = _lib.NetParams()
    assert L.a3gc_net_forward(7, C.byref(p), None, None, None, None, None, None, 1, 1, 12, 8, 3, 0, 0, None, 0, None) == -1
    assert L.a3gc_net_forward(1, C.byref(p), None, None, None, None, None, None, 1, 1, 12, 8, 3, 0, 0, None, 0, None) == -1   # NULL x/y or params
    assert L.a3gc_prepare_input(None, None, None, None, None, None, None, 3, 12, None) == -1
    # engine selection errors are reported, not silently rerouted
    assert L.a3gc_layer_workspace_bytes(1, 4, 4, 10, 10, 2, 0, 2) == 0      # TC engine needs hidden % 64 == 0
    assert b"tensor-core" in L.a3gc_last_error()
    assert L.a3gc_layer_workspace_bytes(1, 4, 4, 10, 10, 2, 1, 1) == 0      # SIMT engine is fp32 only


def test_workspace_sizes():
    L = _lib.lib()
    n = L.a3gc_net_workspace_bytes(1, 2, 3, 12, 16, 3, 0, 1)
    frames = 2 * 3
    assert n >= frames * 15 * (16 + 32 + 32) * 4
    assert L.a3gc_layer_workspace_bytes(3, 2, 3, 16, 16, 2, 0, 1) > 0
    assert L.a3gc_launch_count() == 0
