"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/kb_b200.h declares,
and refuses to run without a CUDA device (no CPU fallback).  No compute calls."""
from __future__ import annotations

import ctypes
import os
import re

import pytest

from kubebrain_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "kb_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kb_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _declared() == sorted(_lib.ABI_SYMBOLS)


def test_library_exports_every_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), f"libkbb200.so does not export {name}"
    assert L.kb_abi_version() == 1


def test_no_cpu_fallback(have_gpu):
    if have_gpu:
        pytest.skip("a CUDA device is present")
    with pytest.raises(_lib.KbError) as ei:
        _lib.Engine(0)
    assert ei.value.code == _lib.KB_ECUDA


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.lib()


def test_product_does_not_touch_the_oracle():
    """the product package and csrc never import, link or mention the oracle"""
    pkg = os.path.join(ROOT, "kubebrain_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in text.lower() or f == "synth.py", os.path.join(dirpath, f)
    out = os.popen(f"ldd {_lib.LIB_PATH} 2>/dev/null").read() if os.path.exists(_lib.LIB_PATH) else ""
    assert "kboracle" not in out
