"""etcd wire encoding: the oracle's encoder and the library's host-side framing against golden
bytes produced by the protobuf runtime from the restated etcd v3.5.2 schema (tests/golden/etcd_schema.py,
make_wire_golden.py).  CPU only; the device encoder is compared with the oracle in tests/test_gpu_parity.py."""
from __future__ import annotations

import json
import os
import struct

import numpy as np
import pytest

from kubebrain_b200 import wire
from kubebrain_b200.packed import PackedStore
from oracle import binding as ko

MAGIC = b"\x57\xfb\x80\x8b"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wire_golden.json")


def _cases():
    with open(GOLD) as f:
        return json.load(f)


def _store_and_recs(kvs):
    """a store holding exactly the records of `kvs` + the record index of every kv (emission order != key order)"""
    items = {MAGIC + k + b"$" + struct.pack(">Q", rev): v for k, v, rev in kvs}
    store = PackedStore.from_items(list(items.items()))
    index = {k: i for i, k in enumerate(store.keys.tolist())}
    recs = [index[MAGIC + k + b"$" + struct.pack(">Q", rev)] for k, v, rev in kvs]
    return store, recs


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c["name"])
def test_oracle_encoder_matches_protobuf_runtime(case):
    kvs = [(bytes.fromhex(k), bytes.fromhex(v), rev) for k, v, rev in case["kvs"]]
    n = len(kvs)
    store, recs = _store_and_recs(kvs)
    st = ko.OracleStore(store)
    elems, off = ko.wire_encode(st, recs, ko.WIRE_KVS)
    assert len(off) == n + 1 and int(off[-1]) == len(elems)
    for i, (k, v, rev) in enumerate(kvs):
        assert int(off[i + 1] - off[i]) == ko.wire_elem_size(len(k), len(v), rev, ko.WIRE_KVS)
    got = ko.wire_range_head(case["header_rev"]) + elems + ko.wire_range_tail(case["more"], n + (1 if case["more"] else 0))
    assert got.hex() == case["range_response"]
    ev, eoff = ko.wire_encode(st, recs, ko.WIRE_EVENTS)
    batches = [ko.wire_watch_head(0) + ev[int(eoff[i]) : int(eoff[min(i + 300, n)])] for i in range(0, n, 300)]
    assert [b.hex() for b in batches] == case["watch_batches"]
    assert ko.wire_watch_head(case["header_rev"], True).hex() == case["watch_end"]
    assert ko.wire_watch_head(case["header_rev"], True, b"context deadline exceeded").hex() == case["watch_end_err"]


def test_library_framing_matches_oracle():
    """kb_wire_range_head / _tail / kb_wire_watch_head are host code of libkbb200.so (no device needed)"""
    for rev in (0, 1, 127, 128, 1700000010, 2**63 - 1, 2**63, 2**64 - 1):
        assert wire.range_head(rev) == ko.wire_range_head(rev)
        assert wire.watch_head(rev) == ko.wire_watch_head(rev)
        assert wire.watch_head(rev, True) == ko.wire_watch_head(rev, True)
        assert wire.watch_head(rev, True, b"x" * 200) == ko.wire_watch_head(rev, True, b"x" * 200)
    for more in (False, True):
        for count in (0, 1, 300, 10001, 2**40):
            assert wire.range_tail(more, count) == ko.wire_range_tail(more, count)


def test_live_protobuf_runtime_if_present():
    """the committed goldens are reproducible: regenerate two cases with the protobuf runtime of this image"""
    pb = pytest.importorskip("google.protobuf")
    from tests.golden import etcd_schema as es
    M = es.build()
    for case in _cases()[:3]:
        kvs = [(bytes.fromhex(k), bytes.fromhex(v), rev) for k, v, rev in case["kvs"]]
        n = len(kvs)
        assert es.range_response(M, case["header_rev"], kvs, case["more"], n + (1 if case["more"] else 0)).hex() == \
            case["range_response"]
        assert es.watch_cancel(M, case["header_rev"], "").hex() == case["watch_end"]
