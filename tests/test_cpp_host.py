"""Runs the C++ host-mirror test driver (tests/cpp/host_mirror_test.cpp over kubebrain_b200/host/kubebrain.hpp):
the reference's table tests written in a compiled language against the same C ABI."""
from __future__ import annotations

import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


def _build():
    if not os.path.exists(os.path.join(ROOT, "kubebrain_b200", "libkbb200.so")):
        import __graft_entry__

        __graft_entry__.build()
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "tests", "cpp")], stdout=subprocess.DEVNULL,
                          stderr=subprocess.DEVNULL)


def test_cpp_host_mirror_cpu():
    _build()
    out = subprocess.run([EXE, "cpu"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0 and "cpu ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_go_shim_call_sequences_replayed():
    """tests/cpp/shim_replay_test.cpp: every method of the Go shim (storage adaptor, scanner, event slab) as the exact C
    call sequence it issues, against the oracle"""
    _build()
    exe = os.path.join(ROOT, "tests", "cpp", "shim_replay_test")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "shim replay ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
def test_cpp_host_mirror_gpu():
    _build()
    out = subprocess.run([EXE, "gpu"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "gpu ok" in out.stdout, out.stdout + out.stderr
