"""Test scaffolding: a tiny in-memory model of the reference's WRITE path, used only to build the
stores and event streams that the reference's own table tests (backend_test.go) build through
Backend.Create/Update/Delete.  It restates the record formats only:

* create  (pkg/backend/creator/naive.go:53-105): PutIfNotExist(revKey, BE64(rev)) + Put(objKey@rev, val);
  re-create over a deleted-flag revision record CASes the revision record instead.
* update  (pkg/backend/txn.go:249-265): CAS(revKey, BE64(new), BE64(old)) + Put(objKey@new, val)
* delete  (pkg/backend/txn.go:145-190): CAS(revKey, BE64(new)+0x00, BE64(old)) + Put(objKey@new, "tombstone")
* events  (pkg/backend/backend.go:237-256): CREATE/PUT carry the new revision in Kv.Revision; DELETE carries the
  previous value and previous revision, Event.Revision = delete revision.
"""
from __future__ import annotations

import struct
from typing import Dict, List, Optional, Tuple

import numpy as np

from kubebrain_b200.packed import PackedEvents, PackedStore, Slab

MAGIC = b"\x57\xfb\x80\x8b"
TOMBSTONE = b"tombstone"
CREATE, PUT, DELETE = 0, 1, 2


def ikey(user_key: bytes, rev: int) -> bytes:
    return MAGIC + user_key + b"$" + struct.pack(">Q", rev)


class MiniBackend:
    def __init__(self, init_rev: int):
        self.kv: Dict[bytes, bytes] = {}
        self.rev = init_rev
        self.events: List[Tuple[int, int, bytes, bytes, int]] = []  # (type, rev, key, value, kv_rev)

    def _latest(self, key: bytes) -> Tuple[Optional[bytes], int]:
        rv = self.kv.get(ikey(key, 0))
        if rv is None:
            return None, 0
        rev = struct.unpack(">Q", rv[:8])[0]
        val = self.kv.get(ikey(key, rev))
        if val is None or val == TOMBSTONE:
            return None, rev
        return val, rev

    def create(self, key: bytes, val: bytes) -> Tuple[int, bool]:
        self.rev += 1
        rev = self.rev
        rk = ikey(key, 0)
        old = self.kv.get(rk)
        if old is not None and not (len(old) == 9 and struct.unpack(">Q", old[:8])[0] < rev):
            return rev, False
        self.kv[rk] = struct.pack(">Q", rev)
        self.kv[ikey(key, rev)] = val
        self.events.append((CREATE, rev, key, val, rev))
        return rev, True

    def update(self, key: bytes, val: bytes, prev_rev: int) -> Tuple[int, bool]:
        if prev_rev == 0:
            return self.create(key, val)
        self.rev += 1
        rev = self.rev
        rk = ikey(key, 0)
        if self.kv.get(rk) != struct.pack(">Q", prev_rev):
            return rev, False
        self.kv[rk] = struct.pack(">Q", rev)
        self.kv[ikey(key, rev)] = val
        self.events.append((PUT, rev, key, val, rev))
        return rev, True

    def delete(self, key: bytes, expected_rev: int = 0) -> Tuple[int, bool]:
        old_val, mod_rev = self._latest(key)
        self.rev += 1
        rev = self.rev
        if old_val is None:
            return rev, False
        if expected_rev > 0 and expected_rev != mod_rev:
            return rev, False
        self.kv[ikey(key, 0)] = struct.pack(">Q", rev) + b"\x00"
        self.kv[ikey(key, rev)] = TOMBSTONE
        self.events.append((DELETE, rev, key, old_val, mod_rev))
        return rev, True

    def apply_victims(self, store: PackedStore, victims) -> None:
        """store.Del / DelCurrent of every victim record (scanner.go:538-564)"""
        for i in victims:
            self.kv.pop(store.keys[int(i)], None)

    def snapshot(self) -> PackedStore:
        return PackedStore.from_items(list(self.kv.items()))

    def packed_events(self, batch: int = 300) -> PackedEvents:
        keys = Slab.from_list([e[2] for e in self.events])
        rev = np.array([e[1] for e in self.events], dtype=np.uint64)
        n = len(self.events)
        bo = np.array(list(range(0, n, batch)) + [n], dtype=np.uint64) if n else np.zeros(1, np.uint64)
        return PackedEvents(keys, rev, bo)
