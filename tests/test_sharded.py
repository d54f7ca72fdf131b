"""Host logic of the N > 1 deployment (kubebrain_b200/sharded.py): routing and merging, CPU only.  The end-to-end
version with two gloo ranks and the oracle's scans is tests/test_multirank_cpu.py."""
from __future__ import annotations

import numpy as np

from kubebrain_b200 import sharded, synth


def test_shard_of_key_matches_the_generator():
    """the synthetic stores are sharded by synth.ns_shard(ns id); sharded.shard_of_key must agree on the keys"""
    for world in (1, 2, 4, 8):
        ids = np.arange(200)
        exp = synth.ns_shard(ids, world)
        for i in ids:
            key = b"/registry/pods/ns-%05d/some-object" % i
            assert sharded.shard_of_key(key, world) == int(exp[i])
            assert sharded.owner_of_prefix(b"/registry/pods/ns-%05d/" % i, world) == int(exp[i])


def test_prefix_routing():
    assert sharded.owner_of_prefix(b"/registry/", 4) is None
    assert sharded.owner_of_prefix(b"/registry/pods/", 4) is None
    assert sharded.owner_of_prefix(b"/registry/pods/ns-000", 4) is None  # the namespace segment is not complete
    assert sharded.owner_of_prefix(b"/registry/pods/ns-00012/", 4) is not None
    assert sharded.owner_of_prefix(b"/registry/pods/ns-00012/web-", 4) == sharded.owner_of_prefix(b"/registry/pods/ns-00012/", 4)
    assert sharded.owner_of_prefix(b"/registry/", 1) == 0
    # cluster-scoped objects are placed by their resource
    assert sharded.shard_of_key(b"/registry/nodes/node-17", 8) == sharded.shard_of_key(b"/registry/nodes/node-99", 8)


def test_merge_list_runs_and_limit():
    runs = [[(b"a", b"1", 5), (b"d", b"4", 6)], [(b"b", b"2", 7)], [], [(b"c", b"3", 8), (b"e", b"5", 9)]]
    full, more = sharded.merge_list_runs(runs)
    assert [k for k, _, _ in full] == [b"a", b"b", b"c", b"d", b"e"] and more is False
    got, more = sharded.merge_list_runs(runs, limit=3)
    assert [k for k, _, _ in got] == [b"a", b"b", b"c"] and more is True
    got, more = sharded.merge_list_runs(runs, limit=5)
    assert len(got) == 5 and more is False
    assert sharded.merge_list_runs([], 10) == ([], False)


def test_merge_watch_streams_and_cursor():
    a = [(3, b"x"), (9, b"y")]
    b = [(1, b"p"), (4, b"q"), (10, b"r")]
    assert [r for r, _ in sharded.merge_watch_streams([a, b, []])] == [1, 3, 4, 9, 10]
    assert sharded.readable_revision([1007, 1000, 1012]) == 1000
    assert sharded.fnv1a64(b"") == 14695981039346656037
