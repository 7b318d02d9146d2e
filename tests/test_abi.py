"""The C-ABI library loads and exports every symbol include/qmpc.h declares (no compute calls: CPU only)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "qmpc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(q(?:mpc|rgp)_[a-zA-Z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from mpc_quad_ros_b200 import _capi
    lib = ctypes.CDLL(_capi.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 35
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    lib.qmpc_version.restype = ctypes.c_int
    assert lib.qmpc_version() == 100
    lib.qmpc_launch_count.restype = ctypes.c_longlong
    assert lib.qmpc_launch_count() == 0


def test_config_struct_layout_matches_header():
    from mpc_quad_ros_b200._capi import QmpcConfig
    # 8 ints + 3 doubles + 20 + 17 + 13 + 2 + 9 doubles + pointer + 10 policy ints
    assert ctypes.sizeof(QmpcConfig) == 8 * 4 + (3 + 20 + 17 + 13 + 2 + 9) * 8 + 8 + 10 * 4


def test_missing_library_fails_loudly(monkeypatch):
    from mpc_quad_ros_b200 import _capi
    monkeypatch.setattr(_capi, "_LIB", None)
    monkeypatch.setattr(_capi, "LIB_PATH", "/nonexistent/libqmpc.so")
    with pytest.raises(_capi.QmpcError):
        _capi.lib()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mpc_quad_ros_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("the oracle", "").lower() or f == "__init__.py", f
