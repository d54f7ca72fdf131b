"""GPU parity of the device-side etcd wire encoder (KB_WIRE_ETCD_KVS / KB_WIRE_ETCD_EVENTS, csrc/kb_wire.cuh) against
the oracle's encoder (itself pinned to the protobuf runtime by tests/test_wire.py): byte-exact element streams, element
offsets, key / value offsets inside the elements, and whole framed messages."""
from __future__ import annotations

import json
import os
import struct

import numpy as np
import pytest

from kubebrain_b200 import synth, wire
from kubebrain_b200._lib import (KB_OUT_DEVICE, KB_OUT_HOST, KB_WIRE_ETCD_EVENTS, KB_WIRE_ETCD_KVS, Engine, KbError)
from kubebrain_b200.coder import NormalCoder, prefix_end
from kubebrain_b200.packed import PackedStore
from oracle import binding as ko
from tests import fuzz

pytestmark = pytest.mark.gpu
CODER = NormalCoder()
MAGIC = b"\x57\xfb\x80\x8b"


@pytest.fixture(scope="module")
def eng():
    e = Engine(0)
    yield e
    e.close()


def check_wire(eng, store, st, reqs):
    plain = eng.range_batch(reqs, KB_OUT_HOST)
    for mode, omode in ((KB_WIRE_ETCD_KVS, ko.WIRE_KVS), (KB_WIRE_ETCD_EVENTS, ko.WIRE_EVENTS)):
        res = eng.range_batch(reqs, KB_OUT_HOST | mode)
        assert res.wire == mode
        assert res.req_first.tolist() == plain.req_first.tolist()
        assert res.req_count.tolist() == plain.req_count.tolist()
        assert res.req_examined.tolist() == plain.req_examined.tolist()
        assert res.rec_idx.tolist() == plain.rec_idx.tolist()
        assert res.rev.tolist() == plain.rev.tolist()
        exp, off = ko.wire_encode(st, res.rec_idx.astype(np.uint64), omode)
        assert res.n_bytes == len(exp)
        assert res.elem_off.tolist() == off.tolist()
        assert res.arena.tobytes() == exp
        for q in range(len(reqs)):
            assert res.kvs(q) == plain.kvs(q), (q, "key/value offsets inside the elements")
        res.close()
    plain.close()


@pytest.mark.parametrize("seed", range(8))
def test_wire_fuzz(eng, seed):
    store = fuzz.fuzz_store(500 + seed, n_keys=40 + 20 * seed)
    st = ko.OracleStore(store)
    eng.load_sorted(store)
    eng.set_compact_revision(None)
    reqs = []
    for s, e in fuzz.fuzz_bounds(store, seed):
        for rev in (0, 23, 60, 2**64 - 1):
            for lim in (0, 1, 5):
                reqs.append((s, e, rev, lim))
    check_wire(eng, store, st, reqs)


def test_wire_varint_boundaries(eng):
    """revisions / lengths around every varint boundary, empty user key, empty value, rev >= 2^63"""
    items = {}
    revs = [1, 127, 128, 16383, 16384, 2**21 - 1, 2**21, 2**28, 2**35, 2**42, 2**49, 2**56, 2**63 - 1, 2**63, 2**64 - 2]
    for i, rev in enumerate(revs):
        uk = b"/k/%03d" % i + b"x" * [0, 1, 100, 113, 114, 120, 127, 128, 300][i % 9]
        items[MAGIC + uk + b"$" + b"\x00" * 8] = struct.pack(">Q", rev)
        items[MAGIC + uk + b"$" + struct.pack(">Q", rev)] = bytes([i]) * [0, 1, 126, 127, 128, 16383, 16384, 70000][i % 8]
    items[MAGIC + b"$" + b"\x00" * 8] = struct.pack(">Q", 9)  # empty user key
    items[MAGIC + b"$" + struct.pack(">Q", 9)] = b""
    store = PackedStore.from_items(list(items.items()))
    st = ko.OracleStore(store)
    eng.load_sorted(store)
    eng.set_compact_revision(None)
    check_wire(eng, store, st, [(b"\x00", b"\xff" * 4, 2**64 - 1, 0), (b"\x00", b"\xff" * 4, 2**64 - 1, 4),
                                (b"\x00", b"\xff" * 4, 2**40, 0)])


def test_wire_messages_config2_shape(eng):
    """whole framed messages for a List and a range stream at the config-2 record shape, via kubebrain_b200.wire"""
    store, meta = synth.gen_store(4000, 4, 256, 2048, 50, config_id=2)
    st = ko.OracleStore(store)
    eng.load_sorted(store)
    eng.set_compact_revision(None)
    p = b"/registry/pods/"
    lo, hi = CODER.encode_object_key(p, 0), CODER.encode_object_key(prefix_end(p), 0)
    limit = 501
    res = eng.range_batch([(lo, hi, meta.read_rev, limit)], KB_OUT_HOST | KB_WIRE_ETCD_KVS)
    n = int(res.req_first[1])
    more = n > limit - 1  # backend.List asks for limit+1 and reports More (range.go:150-190); here: raw scanner answer
    exp_elems, _ = ko.wire_encode(st, res.rec_idx.astype(np.uint64), ko.WIRE_KVS)
    assert wire.range_response(res, 0, meta.read_rev, more) == \
        ko.wire_range_head(meta.read_rev) + exp_elems + ko.wire_range_tail(more, n + (1 if more else 0))
    res.close()
    full = (CODER.encode_object_key(b"/registry/", 0), CODER.encode_object_key(b"/registry0", 0), meta.read_rev, 0)
    res = eng.range_batch([full], KB_OUT_HOST | KB_WIRE_ETCD_EVENTS)
    n = int(res.req_first[1])
    assert n > 600
    ev, eoff = ko.wire_encode(st, res.rec_idx.astype(np.uint64), ko.WIRE_EVENTS)
    exp = [ko.wire_watch_head(0) + ev[int(eoff[i]) : int(eoff[min(i + 300, n)])] for i in range(0, n, 300)]
    exp.append(ko.wire_watch_head(meta.read_rev, True))
    assert list(wire.stream_messages(res, 0, meta.read_rev)) == exp
    res.close()
    # device-resident variant: same bytes, left in HBM
    dres = eng.range_batch([full], KB_OUT_DEVICE | KB_WIRE_ETCD_EVENTS)
    assert dres.on_device and dres.n_bytes == len(ev) and dres.n_kvs == n
    dres.close()


def test_wire_golden_messages(eng):
    """the committed protobuf-runtime goldens, end to end through the device encoder"""
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wire_golden.json")) as f:
        cases = json.load(f)
    for case in cases:
        kvs = [(bytes.fromhex(k), bytes.fromhex(v), rev) for k, v, rev in case["kvs"]]
        if not kvs:
            continue
        # one object per (key, rev): a revision record naming that revision + the object record; user keys that repeat
        # with several revisions would collapse to one winner, so give every kv its own key prefix
        items = {}
        order = []
        for i, (k, v, rev) in enumerate(kvs):
            if v == b"tombstone" or rev == 0:  # never emitted by a scan (rev 0 names a revision record)
                continue
            uk = b"%06d/" % i + k
            items[MAGIC + uk + b"$" + b"\x00" * 8] = struct.pack(">Q", rev)
            items[MAGIC + uk + b"$" + struct.pack(">Q", rev)] = v
            order.append((uk, v, rev))
        store = PackedStore.from_items(list(items.items()))
        eng.load_sorted(store)
        eng.set_compact_revision(None)
        res = eng.range_batch([(b"\x00", b"\xff" * 4, 2**64 - 1, 0)], KB_OUT_HOST | KB_WIRE_ETCD_KVS)
        assert res.kvs(0) == order
        from tests.golden import etcd_schema as es
        M = es.build()
        n = len(order)
        assert wire.range_response(res, 0, case["header_rev"], case["more"]) == \
            es.range_response(M, case["header_rev"], order, case["more"], n + (1 if case["more"] else 0))
        res.close()


def test_wire_rejects_count_mode(eng):
    from kubebrain_b200._lib import KB_OUT_COUNT
    store = fuzz.fuzz_store(1, n_keys=10)
    eng.load_sorted(store)
    with pytest.raises(KbError):
        eng.range_batch([(b"\x00", b"\xff", 0, 0)], KB_OUT_COUNT | KB_WIRE_ETCD_KVS)
    with pytest.raises(KbError):
        eng.range_batch([(b"\x00", b"\xff", 0, 0)], KB_OUT_HOST | KB_WIRE_ETCD_KVS | KB_WIRE_ETCD_EVENTS)


def test_device_resident_results_match_host_results(eng):
    """KB_OUT_DEVICE returns before the copy into the arena has finished (stream-ordered results): after kb_sync the
    arena holds exactly the bytes KB_OUT_HOST returns, in the arena and in both wire modes, call after call"""
    store, meta = synth.gen_store(20000, 4, 256, 2048, 50, config_id=2)
    eng.load_sorted(store)
    eng.set_compact_revision(None)
    lo, hi = CODER.encode_object_key(b"/registry/", 0), CODER.encode_object_key(b"/registry0", 0)
    p = b"/registry/pods/"
    reqs = [(lo, hi, meta.read_rev, 0), (CODER.encode_object_key(p, 0), CODER.encode_object_key(prefix_end(p), 0),
                                         meta.read_rev, 501)]
    for mode in (0, KB_WIRE_ETCD_KVS, KB_WIRE_ETCD_EVENTS):
        host = eng.range_batch(reqs, KB_OUT_HOST | mode)
        want = host.arena.tobytes()
        # several device-resident calls back to back: each one starts while the previous gather may still be running
        devs = [eng.range_batch(reqs, KB_OUT_DEVICE | mode) for _ in range(4)]
        devs[0].wait()  # the first answer alone (host-blocking form of kb_result_wait)
        assert eng.read_device(devs[0].bytes_ptr, devs[0].n_bytes, sync=False) == want
        for d in devs:
            assert (d.n_kvs, d.n_bytes) == (host.n_kvs, host.n_bytes)
            assert d.req_count.tolist() == host.req_count.tolist()
            assert eng.read_device(d.bytes_ptr, d.n_bytes) == want
            d.close()
        host.close()
