"""Pins the CPU oracle (oracle/kb_oracle.c) against the reference's own golden vectors G1..G7
(SURVEY.md section 8c).  Each test names the reference test it reproduces.  CPU only."""
from __future__ import annotations

import struct

import numpy as np
import pytest

from kubebrain_b200.packed import PackedEvents, PackedStore, PackedWatchers, Slab
from oracle import binding as ko
from tests.refmodel import CREATE, DELETE, MAGIC, PUT, TOMBSTONE, MiniBackend, ikey

MAXREV = 2**64 - 1


# ---- G1: pkg/backend/coder/normal_test.go:23-32 TestCompatible ---------------------------------------
def test_g1_coder_known_answer():
    bs = bytes([87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 36,
                0, 0, 0, 0, 0, 0, 0, 0])
    uk, rev, err = ko.decode(bs)
    assert err == 0
    assert uk == b"/registry/test"
    assert rev == 0
    assert ko.encode_object_key(b"/registry/test", 0) == bs
    # encode/decode round trip with a non-zero revision
    k = ko.encode_object_key(b"/registry/pods/ns/a", 0x0102030405060708)
    assert k == MAGIC + b"/registry/pods/ns/a$" + bytes([1, 2, 3, 4, 5, 6, 7, 8])
    assert ko.decode(k) == (b"/registry/pods/ns/a", 0x0102030405060708, 0)


def test_coder_errors_and_parse_revision():
    # normal.go:59-65
    assert ko.decode(b"\x00\x00\x00\x00abc$" + b"\x00" * 8)[2] == -1  # bad magic
    assert ko.decode(MAGIC + b"abcd" + b"\x00" * 9)[2] == -2  # bad split byte
    assert ko.decode(MAGIC + b"$")[2] == -3  # Go would panic (index out of range)
    # rev.go:32-47
    assert ko.parse_revision(struct.pack(">Q", 77)) == (77, False, 0)
    assert ko.parse_revision(struct.pack(">Q", 77) + b"\x00") == (77, True, 0)
    assert ko.parse_revision(b"\x01\x02")[2] == -4


def test_prefix_end():
    # pkg/backend/util.go:70-83 ; etcd semantics
    assert ko.prefix_end(b"/registry/test") == b"/registry/tesu"
    assert ko.prefix_end(b"/registry/test/") == b"/registry/test0"
    assert ko.prefix_end(b"a\xff\xff") == b"b"
    assert ko.prefix_end(b"\xff\xff") == b"\x00"
    assert ko.prefix_end(b"") == b"\x00"


# ---- G2: pkg/backend/scanner/scanner_test.go:27-77 TestAdjustPartitionBorders ---------------------
G2_KEYS = [
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 36, 0, 0, 0, 0, 0, 0, 0, 0],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 101, 118, 101, 110, 116, 115, 47, 98, 100, 101, 102, 97, 117, 108, 116, 47, 118, 107, 45, 116, 101, 115, 116, 45, 112, 111, 100, 45, 118, 113, 115, 114, 106, 46, 49, 54, 98, 101, 101, 51, 101, 55, 56, 52, 98, 50, 101, 48, 101, 57],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 101, 118, 101, 110, 116, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 116, 101, 115, 116, 45, 115, 105, 100, 101, 99, 97, 114, 45, 116, 101, 115, 116, 45, 55, 52, 57, 54, 53, 100, 55, 98, 55, 57, 45, 99, 120, 99, 104, 112, 46, 49, 54, 99, 49, 51, 55, 100, 49, 54, 57, 98, 48, 56, 54, 99, 98],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 101, 118, 101, 110, 116, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 116, 101, 115, 116, 45, 115, 105, 100, 101, 99, 97, 114, 45, 116, 101, 115, 116, 45, 55, 52, 57, 54, 53, 100, 55, 98, 55, 57, 45, 108, 103, 119, 112, 99, 46, 49, 54, 99, 49, 52, 54, 101, 99, 48, 53, 54, 53, 51, 100, 97, 102],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 101, 118, 101, 110, 116, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 116, 101, 115, 116, 45, 115, 105, 100, 101, 99, 97, 114, 45, 116, 101, 115, 116, 45, 55, 52, 57, 54, 53, 100, 55, 98, 55, 57, 45, 115, 55, 107, 57, 120, 46, 49, 54, 99, 49, 51, 57, 100, 48, 57, 101, 50, 54, 49, 97, 57, 102],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 101, 118, 101, 110, 116, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 116, 101, 115, 116, 45, 115, 105, 100, 101, 99, 97, 114, 45, 116, 101, 115, 116, 45, 57, 55, 98, 98, 57, 53, 55, 52, 55, 45, 50, 108, 102, 52, 104, 46, 49, 54, 99, 49, 51, 56, 55, 54, 100, 49, 49, 49, 52, 56, 49, 99],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 101, 118, 101, 110, 116, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 118, 107, 45, 112, 101, 114, 102, 111, 114, 109, 97, 99, 101, 45, 112, 111, 100, 45, 114, 120, 120, 115, 52, 46, 49, 54, 98, 101, 98, 98, 97, 101, 55, 97, 51, 56, 54, 54, 52, 57],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 112, 111, 100, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 116, 101, 115, 116, 45, 115, 105, 100, 101, 99, 97, 114, 45, 116, 101, 115, 116, 45, 55, 52, 57, 54, 53, 100, 55, 98, 55, 57, 45, 100, 108, 122, 108, 50],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 112, 111, 100, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 116, 101, 115, 116, 45, 115, 105, 100, 101, 99, 97, 114, 45, 116, 101, 115, 116, 45, 55, 52, 57, 54, 53, 100, 55, 98, 55, 57, 45, 110, 122, 98, 106, 99],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 112, 111, 100, 115, 47, 116, 101, 115, 116, 47, 99, 114, 45, 56, 53, 53, 53, 55, 102, 99, 100, 45, 108, 112, 118, 45, 116, 101, 115, 116, 45, 118, 107, 54, 45, 104, 108, 45, 100, 114, 105, 118, 101, 114, 45, 100, 57, 122, 100, 110],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 112, 111, 100, 115, 47, 116, 101, 115, 116, 47, 99, 114, 45, 56, 53, 53, 53, 55, 102, 99, 100, 45, 116, 101, 115, 116, 45, 115, 116, 97, 116, 117, 115, 45, 99, 97, 99, 104, 101, 45, 116, 101, 115, 116, 45, 118, 107, 54, 45, 104, 108, 45, 116, 101, 115, 116, 45, 115, 116, 97, 116, 117, 115, 45, 99, 97, 99, 104, 101, 45, 108, 116, 55, 103, 113],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 114, 101, 112, 108, 105, 99, 97, 115, 101, 116, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 100, 112, 45, 57, 55, 49, 55, 50, 57, 54, 50, 53, 98, 45, 54, 98, 100, 52, 52, 57, 57, 52, 102, 56],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 116, 101, 115, 116, 47, 115, 116, 97, 116, 101, 102, 117, 108, 115, 101, 116, 101, 120, 116, 101, 110, 115, 105, 111, 110, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 100, 112, 45, 49, 98, 56, 54, 56, 51, 51, 57, 53, 97, 45, 48],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 116, 101, 115, 116, 47, 115, 116, 97, 116, 101, 102, 117, 108, 115, 101, 116, 101, 120, 116, 101, 110, 115, 105, 111, 110, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 100, 112, 45, 52, 97, 49, 99, 51, 97, 56, 56, 57, 57, 48, 45, 48],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 116, 101, 115, 116, 47, 115, 116, 97, 116, 101, 102, 117, 108, 115, 101, 116, 101, 120, 116, 101, 110, 115, 105, 111, 110, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 100, 112, 45, 53, 49, 97, 54, 49, 97, 97, 102, 50, 100, 45, 48],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 116, 101, 115, 116, 47, 115, 116, 97, 116, 101, 102, 117, 108, 115, 101, 116, 101, 120, 116, 101, 110, 115, 105, 111, 110, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 100, 112, 45, 55, 56, 53, 101, 50, 55, 101, 53, 56, 50, 45, 48],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 116, 101, 115, 116, 47, 115, 116, 97, 116, 101, 102, 117, 108, 115, 101, 116, 101, 120, 116, 101, 110, 115, 105, 111, 110, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 100, 112, 45, 57, 100, 52, 51, 57, 50, 101, 55, 51, 52, 45, 48],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 116, 101, 115, 116, 47, 115, 116, 97, 116, 101, 102, 117, 108, 115, 101, 116, 101, 120, 116, 101, 110, 115, 105, 111, 110, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 100, 112, 45, 99, 57, 50, 50, 51, 101, 49, 50, 102, 98, 45, 48],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 47, 116, 101, 115, 116, 47, 115, 116, 97, 116, 101, 102, 117, 108, 115, 101, 116, 101, 120, 116, 101, 110, 115, 105, 111, 110, 115, 47, 100, 101, 102, 97, 117, 108, 116, 47, 100, 112, 45, 100, 57, 102, 97, 49, 56, 51, 55, 101, 54, 45, 48],
    [87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 48, 36, 0, 0, 0, 0, 0, 0, 0, 0],
]


def test_g2_adjust_partition_borders():
    keys = [bytes(k) for k in G2_KEYS]
    assert keys == sorted(keys)
    adjusted = ko.adjust_partition_borders(keys)
    assert len(adjusted) == len(keys)
    # the reference's assertion: no partition may be empty (scanner_test.go:75)
    for i in range(1, len(adjusted)):
        assert adjusted[i - 1] != adjusted[i]
    # first and last borders are never touched; interior borders that fail Decode stay as they are,
    # interior borders that decode with rev != 0 move to the revision key of their user key
    assert adjusted[0] == keys[0] and adjusted[-1] == keys[-1]
    for i in range(1, len(keys) - 1):
        uk, rev, err = ko.decode(keys[i])
        if err == 0 and rev != 0:
            assert adjusted[i] == ko.encode_object_key(uk, 0)
        else:
            assert adjusted[i] == keys[i]


def test_adjust_moves_object_key_border_to_revision_key():
    a = ko.encode_object_key(b"/registry/a", 0)
    mid = ko.encode_object_key(b"/registry/k", 17)
    z = ko.encode_object_key(b"/registry/z", 0)
    assert ko.adjust_partition_borders([a, mid, z]) == [a, ko.encode_object_key(b"/registry/k", 0), z]
    # the first and last borders are left alone even if they are object keys (scanner.go:212-221)
    assert ko.adjust_partition_borders([mid, z]) == [mid, z]


# ---- G3: pkg/backend/ring_test.go:26-108 TestRing -------------------------------------------------
def test_g3_ring_find_events_table():
    r = ko.Ring(10)
    ret, revs, _ = r.find(5)
    assert ret.empty
    r.add(1)
    ret, revs, _ = r.find(1)
    assert (ret.oldest_rev, ret.newest_rev) == (1, 1)
    r.reset()
    for i in range(1, 21):
        r.add(i, i)
    table = [
        (9, False, True, []),
        (10, False, True, []),
        (11, False, False, list(range(11, 21))),
        (12, False, False, list(range(12, 21))),
        (15, False, False, list(range(15, 21))),
        (19, False, False, [19, 20]),
        (20, False, False, [20]),
        (21, True, False, []),
        (30, True, False, []),
    ]
    for rev, high, low, events in table:
        ret, revs, pay = r.find(rev)
        assert bool(ret.high) == high and bool(ret.low) == low
        assert (ret.oldest_rev, ret.newest_rev) == (11, 20)
        assert revs.tolist() == events
        assert pay.tolist() == events
    r2 = ko.Ring(10)
    for i in range(1, 8):
        r2.add(i)
    ret, revs, _ = r2.find(7)
    assert ret.newest_rev == 7 and revs.tolist() == [7]


def test_ring_wraparound_every_phase():
    # every fill level / wrap phase against a trivial list model
    for cap in (1, 3, 7):
        r = ko.Ring(cap)
        model = []
        for i in range(1, 4 * cap + 2):
            r.add(i * 2, i)
            model = (model + [i * 2])[-cap:]
            for q in range(0, 2 * i + 3):
                ret, revs, _ = r.find(q)
                if q > model[-1]:
                    assert ret.high
                elif q < model[0]:
                    assert ret.low
                else:
                    assert revs.tolist() == [x for x in model if x >= q]


# ---- G4: pkg/backend/compact_test.go:36-81 ------------------------------------------------------
def test_g4_compact_borders():
    enc = lambda s: ko.encode_object_key(s, 0)
    got = ko.compact_borders(b"/registry/test", [b"/registry/test/pods", b"/registry/test/events"])
    assert got == [enc(b"/registry/test/"), enc(b"/registry/test/events/"), enc(b"/registry/test/events0"),
                   enc(b"/registry/test/pods/"), enc(b"/registry/test/pods0"), enc(b"/registry/test0")]
    assert ko.compact_borders(b"/registry/test") == [enc(b"/registry/test/"), enc(b"/registry/test0")]


# ---- G5: pkg/backend/backend_test.go:740-901 testBackendRange -------------------------------------
PREFIX = b"/registry/test"  # backend_test.go prefix const


def _list(st, store, key: bytes, end: bytes, rev: int, limit: int, cur_rev: int):
    """backend.List (range.go:124-174) on top of the oracle's scanner.Range"""
    if len(end) == 0:
        return None
    req = rev or cur_rev
    if key >= end:
        return None  # "invalid range end"
    lim = limit + 1 if limit > 0 else limit
    r = ko.range_(st, ko.encode_object_key(key, 0), ko.encode_object_key(end, 0), req, lim)
    assert r.rc == 0
    kvs = r.kvs(store)
    more = False
    if lim > 0 and len(kvs) > limit:
        more, kvs = True, kvs[:limit]
    return kvs, more


def test_g5_backend_range_table():
    inject = 10
    init0 = 1_700_000_000
    b = MiniBackend(init0)
    test_key = PREFIX + b"/" + b"key"  # path.Join(prefix, testKey)
    end_key = ko.prefix_end(test_key)
    fmt = lambda p, i: p + b"/" + (b"%05d" % i)
    invalid_rev = b.rev
    kv_list = []
    for i in range(inject):
        rev, ok = b.create(fmt(test_key, i), fmt(b"val", i))
        assert ok
        kv_list.append((fmt(test_key, i), fmt(b"val", i), rev))
    init = b.rev
    store = b.snapshot()
    st = ko.OracleStore(store)

    # get (range.go:34-121)
    idx, mod = ko.get(st, fmt(test_key, inject - 1), 0)
    assert idx >= 0 and mod == init and store.vals[idx] == fmt(b"val", inject - 1)
    idx, mod = ko.get(st, fmt(test_key, inject - 2), init)
    assert idx >= 0 and mod == init - 1
    idx, mod = ko.get(st, fmt(test_key, inject - 1), invalid_rev)
    assert idx == -1
    idx, mod = ko.get(st, test_key + b"/-0001", 0)
    assert idx == -1

    # list with prefix
    assert _list(st, store, test_key, end_key, 0, 0, init) == (kv_list, False)
    # list with range end
    assert _list(st, store, test_key, fmt(test_key, inject - 2), 0, 0, init) == (kv_list[: inject - 2], False)
    # list with range end & limit  -> More
    assert _list(st, store, test_key, fmt(test_key, inject - 2), 0, inject - 4, init) == (kv_list[: inject - 4], True)
    # list with invalid prefix -> empty
    assert _list(st, store, end_key, fmt(end_key, inject - 2), 0, 0, init) == ([], False)
    # list with invalid range end -> error
    assert _list(st, store, fmt(end_key, inject - 2), end_key, 0, 0, init) is None
    # list with range end & limit & revision
    assert _list(st, store, fmt(test_key, 1), fmt(test_key, inject - 1), init - 2, inject - 5, init) == (
        kv_list[1 : inject - 4], True)
    # list with dir prefix
    assert _list(st, store, test_key, ko.prefix_end(test_key), 0, inject - 5, init) == (kv_list[0 : inject - 5], True)
    # count (scanner.Count, scanner.go:121-126)
    r = ko.scan(st, [ko.encode_object_key(test_key, 0), ko.encode_object_key(end_key, 0)], init, collect=False)
    assert r.count == inject
    r = ko.scan(st, [ko.encode_object_key(end_key, 0), ko.encode_object_key(ko.prefix_end(end_key), 0)], init, collect=False)
    assert r.count == 0
    # partitions + ListByStream == List (backend_test.go:883-900): any split of the interval gives the same kvs
    start, end = ko.encode_object_key(test_key, 0), ko.encode_object_key(end_key, 0)
    for cut in range(1, store.n):
        borders = sorted({start, store.keys[cut], end})
        r = ko.scan(st, borders, init)
        assert r.kvs(store) == kv_list, cut
        assert r.count == inject


def test_scan_quirks_q1_q6():
    """Q1..Q6 of SURVEY.md 8a, each on a hand-built store."""
    b = MiniBackend(100)
    r1, _ = b.create(b"/r/a", b"a1")  # 101
    r2, _ = b.update(b"/r/a", b"a2", r1)  # 102
    r3, _ = b.create(b"/r/b", b"b1")  # 103
    r4, _ = b.delete(b"/r/b")  # 104
    r5, _ = b.create(b"/r/c", b"c1")  # 105
    r6, _ = b.update(b"/r/a", b"a3", r2)  # 106
    store = b.snapshot()
    st = ko.OracleStore(store)
    lo, hi = ko.encode_object_key(b"/r/", 0), ko.encode_object_key(b"/r0", 0)
    # latest: a@106, b deleted, c@105
    assert ko.range_(st, lo, hi, 200).kvs(store) == [(b"/r/a", b"a3", 106), (b"/r/c", b"c1", 105)]
    # Q1 time travel: at 103 a@102, b@103 visible, c not yet created (its revision record (rev 0) is visible
    # but is never emitted: prevRevision == 0)
    assert ko.range_(st, lo, hi, 103).kvs(store) == [(b"/r/a", b"a2", 102), (b"/r/b", b"b1", 103)]
    assert ko.range_(st, lo, hi, 100).kvs(store) == []
    # Q3/Q4: limit stops the loop; count is 0 and limit_stop set
    r = ko.range_(st, lo, hi, 200, limit=1)
    assert r.kvs(store) == [(b"/r/a", b"a3", 106)] and r.limit_stop and r.count == 0
    # the trailing object is only emitted when the receiver still needs more
    r = ko.range_(st, lo, hi, 200, limit=2)
    assert [k for k, _, _ in r.kvs(store)] == [b"/r/a", b"/r/c"] and not r.limit_stop and r.count == 2
    # Q2: undecodable records are skipped and do not disturb prev
    items = [(store.keys[i], store.vals[i]) for i in range(store.n)]
    items.append((MAGIC + b"/r/a$" + struct.pack(">Q", 102) + b"x", b"junk"))  # split byte in the wrong place
    items.append((b"\x00\x00\x00\x01zzzz$" + b"\x00" * 8, b"junk"))  # bad magic (sorts first)
    store2 = PackedStore.from_items(items)
    st2 = ko.OracleStore(store2)
    assert ko.range_(st2, b"\x00", b"\xff", 200).kvs(store2) == [(b"/r/a", b"a3", 106), (b"/r/c", b"c1", 105)]
    # Q6: an empty user key compares equal to the initial nil prevUserKey
    store3 = PackedStore.from_items([(ikey(b"", 0), struct.pack(">Q", 5)), (ikey(b"", 5), b"v")])
    st3 = ko.OracleStore(store3)
    assert ko.range_(st3, b"\x00", b"\xff", 200).kvs(store3) == [(b"", b"v", 5)]


# ---- G6: backend_test.go:1134-1251 event streams + watch registration -----------------------------------
def test_g6_events_and_watch_registration():
    init = 5000
    b = MiniBackend(init)
    times = 10
    p = PREFIX + b"/create/and/watch"
    for i in range(times):
        rev, ok = b.create(p + b"/%d" % i, b"val")
        assert ok and rev == init + i + 1
    for i in range(times):
        rev, ok = b.delete(p + b"/%d" % i, init + i + 1)
        assert ok and rev == init + i + times + 1
    ev = b.packed_events()
    # event contents (backend.go:237-256)
    for i in range(times):
        assert b.events[i] == (CREATE, init + i + 1, p + b"/%d" % i, b"val", init + i + 1)
        assert b.events[times + i] == (DELETE, init + times + i + 1, p + b"/%d" % i, b"val", init + i + 1)
    ring = ko.Ring(200000)
    for i, e in enumerate(b.events):
        ring.add(e[1], i)
    cur = b.rev
    # watch from a revision older than the oldest cached event must fail (backend_test.go:1208-1211)
    mode, live, cu, err_rev = ko.watch_register(ring, ev, PREFIX, init, cur)
    assert mode == 2 and err_rev == init + 1
    # getEventsFromRev(initRevision+1, 20): all events, in order
    mode, live, cu, _ = ko.watch_register(ring, ev, PREFIX, init + 1, cur)
    assert mode == 3 and cu.tolist() == list(range(2 * times)) and live == cur + 1
    # from the first delete
    mode, live, cu, _ = ko.watch_register(ring, ev, PREFIX, init + times + 1, cur)
    assert mode == 3 and cu.tolist() == list(range(times, 2 * times))
    # a prefix that matches nothing: catch-up empty, live filter starts at the requested revision (watch.go:91-96)
    mode, live, cu, _ = ko.watch_register(ring, ev, b"/nothing", init + 3, cur)
    assert mode == 3 and len(cu) == 0 and live == init + 3
    # revision 0 => live only; revision beyond newest => live from that revision
    assert ko.watch_register(ring, ev, PREFIX, 0, cur)[:2] == (0, 0)
    assert ko.watch_register(ring, ev, PREFIX, cur + 5, cur)[:2] == (0, cur + 5)
    # empty cache: ok iff revision > current (watch.go:61-72)
    empty = ko.Ring(16)
    assert ko.watch_register(empty, ev, PREFIX, cur + 1, cur)[0] == 0
    assert ko.watch_register(empty, ev, PREFIX, cur, cur)[0] == 1


def test_catchup_chunks():
    # watch.go:102-117
    assert ko.catchup_chunks(1) == [1]
    assert ko.catchup_chunks(300) == [300]
    assert ko.catchup_chunks(301) == [300, 1]
    assert ko.catchup_chunks(30000) == [300] * 100
    n = 30001
    bs = n // 99
    got = ko.catchup_chunks(n)
    assert sum(got) == n and all(x == bs for x in got[:-1]) and 0 < got[-1] <= bs


def test_fanout_semantics():
    # filterByRevision strips only LEADING events below min_rev (watch.go:153-159); filterByPrefix keeps order
    keys = Slab.from_list([b"/a/1", b"/b/1", b"/a/2", b"/a/3", b"/b/2", b"/a/4"])
    rev = np.array([10, 11, 9, 12, 13, 14], dtype=np.uint64)  # note the out-of-order 9
    ev = PackedEvents(keys, rev, np.array([0, 3, 6], dtype=np.uint64))
    w = PackedWatchers(Slab.from_list([b"/a/", b"/b/", b"/", b"/zzz", b""]),
                       np.array([0, 12, 11, 0, 13], dtype=np.uint64))
    start, idx, msgs = ko.fanout(ev, w)
    lists = [idx[int(start[i]) : int(start[i + 1])].tolist() for i in range(w.n)]
    assert lists[0] == [0, 2, 3, 5]
    assert lists[1] == [4]  # batch 0 fully stripped (10,11,9 all < 12); batch 1: 12 passes
    assert lists[2] == [1, 2, 3, 4, 5]  # batch 0: strip rev 10, keep 11 and the later 9
    assert lists[3] == []
    assert lists[4] == [4, 5]
    assert msgs == 2 + 1 + 2 + 0 + 1
    # threads do not change the result
    s2, i2, m2 = ko.fanout(ev, w, threads=3, alloc_per_batch=True)
    assert s2.tolist() == start.tolist() and i2.tolist() == idx.tolist() and m2 == msgs


# ---- G7: compact_test.go:134-284 + backend_test.go:903-965 ------------------------------------------
def _g7_backend():
    init = 9000
    b = MiniBackend(init)
    s = lambda i: PREFIX + b"/compact-consistence/%d" % i
    r, _ = b.create(s(1), s(1))
    for _ in range(3):
        r, ok = b.update(s(1), s(1), r)
        assert ok
    assert b.delete(s(1))[1]
    r, _ = b.create(s(2), s(2))
    assert b.delete(s(2))[1]
    assert not b.delete(s(2))[1]  # delete2-2 fails but consumes a revision
    assert b.rev == init + 8
    return b


def test_g7_compaction_yields_empty_range():
    b = _g7_backend()
    store = b.snapshot()
    st = ko.OracleStore(store)
    lo, hi = ko.compact_borders(PREFIX)
    compact_rev = b.rev - 1
    r = ko.scan(st, [lo, hi], compact_rev, compact=True, collect=False)
    assert r.rc == 0
    # every record of both (fully deleted) objects is a victim: 5+2 object versions (4 superseded + tombstone
    # twice-classified ...) and the two deleted-flag revision records
    by_class = {c: sorted(int(v) for v, cc in zip(r.victims, r.vclass) if cc == c) for c in (1, 2, 3)}
    keys = store.keys.tolist()
    obj1 = [i for i, k in enumerate(keys) if b"/1$" in k]
    obj2 = [i for i, k in enumerate(keys) if b"/2$" in k]
    assert by_class[1] == obj1[1:-1] + obj2[1:-1]  # every version that has a later visible version
    assert by_class[2] == [obj1[-1], obj2[-1]]  # the tombstones
    assert by_class[3] == [obj1[0], obj2[0]]  # the deleted-flag revision records
    # order of the delete calls inside one object: v1, v2, v3, v4(superseded by tombstone), tombstone ; and the
    # revision record is handled when it is read, i.e. FIRST
    order1 = [int(v) for v in r.victims if int(v) in obj1]
    assert order1 == [obj1[0]] + obj1[1:]
    b.apply_victims(store, r.victims)
    after = b.snapshot()
    assert after.n == 0
    # range after compaction returns nothing, twice (compact_test.go:270-283)
    st2 = ko.OracleStore(after)
    assert ko.range_(st2, lo, hi, b.rev).kvs(after) == []
    # range BEFORE applying the victims also returns nothing (both objects are deleted)
    assert ko.range_(st, lo, hi, b.rev).kvs(store) == []
    # a range below the compact revision is refused (scanner.go:618-624)
    assert ko.range_(st, lo, hi, compact_rev - 1, compact_rev=compact_rev).rc == -6
    assert ko.range_(st, lo, hi, compact_rev, compact_rev=compact_rev).rc == 0


def test_g7_compaction_keeps_live_objects_and_q5():
    init = 100
    b = MiniBackend(init)
    ra, _ = b.create(b"/registry/test/a", b"a1")  # 101
    ra2, _ = b.update(b"/registry/test/a", b"a2", ra)  # 102
    rb, _ = b.create(b"/registry/test/b", b"b1")  # 103
    rb2, _ = b.update(b"/registry/test/b", b"b2", rb)  # 104
    rdel, _ = b.delete(b"/registry/test/b")  # 105
    store = b.snapshot()
    st = ko.OracleStore(store)
    lo, hi = ko.compact_borders(PREFIX)
    # compact at 104: b's deleted-flag revision record embeds 105 > 104 -> skipped WITHOUT updating prev (Q5):
    # 'a' is counted twice, b@103 is not superseded-deleted by the first b version it meets... the oracle is the law
    r = ko.scan(st, [lo, hi], 104, compact=True, collect=False)
    keys = store.keys.tolist()
    vic = [(keys[int(v)], int(c)) for v, c in zip(r.victims, r.vclass)]
    assert vic == [(ikey(b"/registry/test/a", 101), 1), (ikey(b"/registry/test/b", 103), 1)]
    assert r.count == 3  # a counted at b's revision record AND at b@103 (double count), plus trailing b@104
    # get at an old revision returns nothing after that version is compacted (backend_test.go:903-965)
    b.apply_victims(store, r.victims)
    after = b.snapshot()
    st2 = ko.OracleStore(after)
    assert ko.get(st2, b"/registry/test/a", 101)[0] == -1
    idx, mod = ko.get(st2, b"/registry/test/a", 0)
    assert after.vals[idx] == b"a2" and mod == 102
    assert ko.get(st2, b"/registry/test/b", 0) == (-2, 105)  # tombstone


# ---- G8: expire_test.go:32-97 (TTL path, only when the store has no native TTL) ------------------------
def test_g8_ttl_expiry_classification():
    b = MiniBackend(10)
    e1, _ = b.create(b"/registry/test/events/ns/e1", b"x")  # 11
    e2, _ = b.create(b"/registry/test/events/ns/e2", b"y")  # 12
    p1, _ = b.create(b"/registry/test/pods/ns/p1", b"z")  # 13
    e3, _ = b.create(b"/registry/test/events/ns/e3", b"w")  # 14
    store = b.snapshot()
    st = ko.OracleStore(store)
    lo, hi = ko.compact_borders(PREFIX)
    keys = store.keys.tolist()
    r = ko.scan(st, [lo, hi], 14, compact=True, collect=False, timeout_rev=12, support_ttl=False)
    vic = [(keys[int(v)], int(c)) for v, c in zip(r.victims, r.vclass)]
    assert vic == [
        (ikey(b"/registry/test/events/ns/e1", 0), 4), (ikey(b"/registry/test/events/ns/e1", 11), 5),
        (ikey(b"/registry/test/events/ns/e2", 0), 4), (ikey(b"/registry/test/events/ns/e2", 12), 5),
    ]
    assert r.count == 2  # e3 and p1 survive
    # with native TTL support (badger) nothing is expired by the scanner
    r = ko.scan(st, [lo, hi], 14, compact=True, collect=False, timeout_rev=12, support_ttl=True)
    assert len(r.victims) == 0 and r.count == 4
