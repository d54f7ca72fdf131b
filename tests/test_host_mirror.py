"""CPU tests of the host-side mirror (coder, Ring, catch-up chunking, compact borders) against the oracle and the
reference's golden vectors."""
from __future__ import annotations

import random

import pytest

from kubebrain_b200.backend import Backend, Event, Ring
from kubebrain_b200.coder import DecodeError, ErrInvalidRevFormat, NormalCoder, parse_revision, prefix_end
from kubebrain_b200.scanner import KeyValue
from oracle import binding as ko


def test_coder_matches_oracle_and_g1():
    c = NormalCoder()
    bs = bytes([87, 251, 128, 139, 47, 114, 101, 103, 105, 115, 116, 114, 121, 47, 116, 101, 115, 116, 36,
                0, 0, 0, 0, 0, 0, 0, 0])
    assert c.decode(bs) == (b"/registry/test", 0)  # coder/normal_test.go:23-32
    rng = random.Random(1)
    for _ in range(200):
        uk = bytes(rng.randrange(256) for _ in range(rng.randint(0, 40)))
        rev = rng.getrandbits(64)
        k = c.encode_object_key(uk, rev)
        assert k == ko.encode_object_key(uk, rev)
        assert c.decode(k) == (uk, rev)
        assert c.encode_revision_key(uk) == ko.encode_object_key(uk, 0)
        assert prefix_end(uk) == ko.prefix_end(uk)
    for bad in (b"\x00" * 20, b"\x57\xfb\x80\x8babc" + b"\x00" * 9, b"\x57\xfb\x80\x8b$"):
        with pytest.raises(DecodeError):
            c.decode(bad)
        assert ko.decode(bad)[2] != 0
    assert parse_revision(b"\x00" * 7 + b"\x05") == (5, False)
    assert parse_revision(b"\x00" * 7 + b"\x05\x00") == (5, True)
    with pytest.raises(ErrInvalidRevFormat):
        parse_revision(b"abc")


def _ev(rev):
    return Event(0, rev, KeyValue(b"k%d" % rev, b"v", rev))


def test_ring_g3_and_oracle():
    r = Ring(10)
    assert r.find_events(5)[0]  # empty
    for i in range(1, 21):
        r.add(_ev(i))
    table = [(9, False, True, []), (10, False, True, []), (11, False, False, list(range(11, 21))),
             (15, False, False, list(range(15, 21))), (20, False, False, [20]), (21, True, False, []),
             (30, True, False, [])]
    for rev, high, low, evs in table:  # ring_test.go:61-107
        empty, h, l, newest, oldest, events = r.find_events(rev)
        assert (h, l) == (high, low) and (oldest.revision, newest.revision) == (11, 20)
        assert [e.revision for e in events] == evs
    for cap in (1, 3, 7):
        a, b = Ring(cap), ko.Ring(cap)
        for i in range(1, 4 * cap + 2):
            a.add(_ev(2 * i))
            b.add(2 * i, i)
            for q in range(0, 2 * i + 3):
                empty, h, l, newest, oldest, events = a.find_events(q)
                ret, revs, _ = b.find(q)
                assert (h, l) == (bool(ret.high), bool(ret.low))
                assert [e.revision for e in events] == revs.tolist()


def test_catchup_chunks_match_oracle():
    for n in (1, 299, 300, 301, 30000, 30001, 123457):
        evs = list(range(n))
        assert [len(c) for c in Backend._catch_up_chunks(evs)] == ko.catchup_chunks(n)


def test_compact_borders_g4():
    b = Backend.__new__(Backend)
    b.coder = NormalCoder()
    b.prefix = "/registry/test"
    b.skipped_prefixes = ["/registry/test/pods", "/registry/test/events"]
    assert b.get_compact_borders() == ko.compact_borders(b"/registry/test", [b"/registry/test/pods", b"/registry/test/events"])
    b.skipped_prefixes = []
    assert b.get_compact_borders() == ko.compact_borders(b"/registry/test")


# ---- range stream host logic with a stub engine (no device): receiver.go:105-166, scanner.go:129-145,179-192 -----------
class _StubResult:
    def __init__(self, kvs):
        self._kvs = kvs

    def kvs(self, q):
        return self._kvs

    def close(self):
        pass


class _StubEngine:
    """answers every range with n synthetic kvs, or raises the error a compacted revision raises"""

    def __init__(self, n, fail=None):
        self.n, self.fail = n, fail

    def range_batch(self, reqs, mode):
        if self.fail:
            raise self.fail
        return _StubResult([(b"/k/%05d" % i, b"v%d" % i, 100 + i) for i in range(self.n)])


def test_range_stream_batches_and_end_marker():
    from kubebrain_b200._lib import KbError
    from kubebrain_b200.scanner import Scanner

    for n in (0, 1, 299, 300, 301, 650):
        msgs = list(Scanner(_StubEngine(n)).range_stream(b"a", b"b", 777))
        batches, end = msgs[:-1], msgs[-1]
        assert [len(m.kvs) for m in batches] == [300] * (n // 300) + ([n % 300] if n % 300 else [])
        assert all(m.more for m in batches)
        # Q7: batches come from forked receivers whose readRev is never set (receiver.go:162-166) -> revision 0
        assert all(m.revision == 0 for m in batches)
        assert (end.revision, end.kvs, end.more, end.err) == (777, [], False, "")  # getListStreamEnd, scanner.go:179-192
        assert [kv.key for m in batches for kv in m.kvs] == [b"/k/%05d" % i for i in range(n)]
    err = KbError(-5, "range stream revision 3 less than compact revision 9")
    msgs = list(Scanner(_StubEngine(5, fail=err)).range_stream(b"a", b"b", 3))
    assert len(msgs) == 1 and msgs[0].more is False and "compact revision 9" in msgs[0].err and msgs[0].revision == 3
