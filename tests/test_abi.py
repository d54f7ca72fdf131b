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
    assert L.kb_abi_version() == 2


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


def test_header_is_plain_c():
    """the boundary is a C ABI: the header must compile as C99 (what cgo feeds it to) and as C++"""
    import shutil
    import subprocess

    hdr = os.path.join(ROOT, "include", "kb_b200.h")
    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr])
    gxx = shutil.which("g++")
    if gxx:
        subprocess.check_call([gxx, "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr])


def test_ctypes_structs_match_the_header_layout():
    """sizes of the structs that cross the boundary by value or by array, as a C compiler lays them out"""
    import shutil
    import subprocess
    import tempfile

    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    from kubebrain_b200 import _lib

    src = '#include <stdio.h>\n#include "kb_b200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(kb_config),sizeof(kb_range_req),sizeof(kb_range_view),sizeof(kb_write_op),sizeof(kb_get_req),' \
          'sizeof(kb_get_view),sizeof(kb_compact_view),sizeof(kb_match_view));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c, exe = os.path.join(d, "s.c"), os.path.join(d, "s")
        with open(c, "w") as f:
            f.write(src)
        subprocess.check_call([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        sizes = [int(x) for x in subprocess.check_output([exe]).split()]
    import ctypes as C
    got = [C.sizeof(t) for t in (_lib.KbConfig, _lib.KbRangeReq, _lib.KbRangeView, _lib.KbWriteOp, _lib.KbGetReq,
                                 _lib.KbGetView, _lib.KbCompactView, _lib.KbMatchView)]
    assert got == sizes


def test_plain_c_program_links_and_calls_the_library(tmp_path):
    """a C99 translation unit links libkbb200.so and calls the host-side entry points (no device needed)"""
    import shutil
    import subprocess

    gcc = shutil.which("gcc")
    if not gcc:
        pytest.skip("no gcc")
    src = r'''
#include <stdio.h>
#include <string.h>
#include "kb_b200.h"
int main(void) {
    unsigned char b[64];
    unsigned long long n = kb_wire_range_head(5, b);
    if (n != 4 || memcmp(b, "\x0a\x02\x18\x05", 4) != 0) return 1;
    n = kb_wire_range_tail(1, 300, b);
    if (n != 5 || memcmp(b, "\x18\x01\x20\xac\x02", 5) != 0) return 2;
    n = kb_wire_watch_head(0, 1, (const unsigned char *)"eof", 3, b);
    if (n != 9 || memcmp(b, "\x0a\x00\x20\x01\x32\x03" "eof", 9) != 0) return 3;
    printf("%d\n", kb_abi_version());
    return 0;
}
'''
    c, exe = str(tmp_path / "link.c"), str(tmp_path / "link")
    with open(c, "w") as f:
        f.write(src)
    libdir = os.path.join(ROOT, "kubebrain_b200")
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), c, "-o", exe,
                           "-L", libdir, "-lkbb200", "-Wl,-rpath," + libdir])
    out = subprocess.check_output([exe]).decode().strip()
    assert int(out) == _lib.lib().kb_abi_version()


def test_entry_points_reject_null_arguments():
    """argument checks come before any CUDA call: the range halves, the prefetch and the waits answer KB_EINVAL on NULL
    handles (what the cgo shim's error path relies on); kb_pending_free / kb_result_free ignore NULL"""
    L = _lib.lib()
    h = ctypes.c_void_p()
    req = (_lib.KbRangeReq * 1)()
    assert L.kb_range_submit(None, req, 1, _lib.KB_OUT_HOST, ctypes.byref(h)) == _lib.KB_EINVAL
    assert L.kb_range_collect(None, None, ctypes.byref(h)) == _lib.KB_EINVAL
    assert L.kb_range_batch(None, req, 1, _lib.KB_OUT_HOST, ctypes.byref(h)) == _lib.KB_EINVAL
    assert L.kb_range_prefetch(None, req, 1) == _lib.KB_EINVAL
    assert L.kb_result_wait(None, None, None) == _lib.KB_EINVAL
    L.kb_pending_free(None, None)
    L.kb_result_free(None, None)
    assert h.value is None
