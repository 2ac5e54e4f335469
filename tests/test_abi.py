"""The C ABI: every entry point include/kg.h declares is bound by kinovagrasping_b200/_lib.py and exported by the
nvcc-built product library (loaded here without a GPU: no compute call is made) and by the test-only emulation."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, _emu_lib
from kinovagrasping_b200 import _lib


def _declared():
    text = open(os.path.join(ROOT, "include", "kg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(kg_[a-z0-9_]+)\s*\(", text))


def test_header_and_binding_table_agree():
    declared = _declared()
    assert len(declared) >= 30
    assert declared == set(_lib.SIGNATURES), (sorted(declared - set(_lib.SIGNATURES)), sorted(set(_lib.SIGNATURES) - declared))


def test_emulation_exports_every_symbol():
    lib = _emu_lib()
    for name in _declared():
        assert hasattr(lib, name), name


def test_product_library_loads_and_exports_every_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), name
    # error paths that need no device: null handles are rejected with the documented code
    lib.kg_num_envs.restype = ctypes.c_int
    assert lib.kg_num_envs(None) < 0
    lib.kg_last_error.restype = ctypes.c_char_p
    assert lib.kg_last_error(None) == b"null handle"


def test_package_has_no_cpu_fallback(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "missing.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()
