"""The C-ABI library loads on a machine without a GPU and exports every symbol that
include/gwsem_b200.h declares (no compute is attempted here)."""
import ctypes
import os
import re

from conftest import ROOT
from gwsem_b200 import _lib


def header_functions():
    src = open(os.path.join(ROOT, "include", "gwsem_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gwsem_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    names = header_functions()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_covers_the_header():
    assert sorted(_lib.SYMBOLS) == header_functions()


def test_abi_version_and_loud_failure_without_gpu():
    lib = _lib.lib()
    assert lib.gwsem_abi_version() == 1
    if lib.gwsem_device_count() == 0:
        h = ctypes.c_void_p()
        assert lib.gwsem_engine_create(0, ctypes.byref(h)) != 0
        assert b"no CPU fallback" in lib.gwsem_last_error()
