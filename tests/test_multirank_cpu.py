"""world_size-2 gloo test of the N>1 host logic (CPU): hash sharding by namespace partitions the store, the
revision-cursor all-gather yields min over ranks, and the union of the per-shard scans equals the global scan."""
from __future__ import annotations

import os
import socket
import subprocess
import sys

import numpy as np

from kubebrain_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ns_shard_is_a_partition():
    ns = np.arange(5000)
    for world in (1, 2, 4, 8):
        s = synth.ns_shard(ns, world)
        assert s.min() >= 0 and s.max() < world
        counts = np.bincount(s, minlength=world)
        assert counts.sum() == 5000
        if world > 1:
            assert counts.min() > 0.7 * 5000 / world  # fnv1a spreads the namespaces evenly


def test_world_size_2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_rank_worker.py")], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert any(o.strip().startswith("OK") or "\nOK" in o for o in outs), outs
