"""A second, independent restatement of the reference's scan loop -- pure Python, written straight from
pkg/backend/scanner/scanner.go:389-516 (worker.run), :566-591 (compactIfExpired), scanner/receiver.go:62-103 and
pkg/backend/coder/normal.go:58-70 -- used only to cross-check the C oracle (tests/test_oracle_crosscheck.py): the two
were written separately, so an agreement on adversarial inputs is evidence for both."""
from __future__ import annotations

import struct
from bisect import bisect_left
from typing import List, Optional, Sequence, Tuple

MAGIC = b"\x57\xfb\x80\x8b"
TOMBSTONE = b"tombstone"
V_SUPERSEDED, V_TOMBSTONE, V_REVRECORD, V_TTL_REVREC, V_TTL_OBJECT = 1, 2, 3, 4, 5


class DecodeError(Exception):
    pass


def decode(key: bytes) -> Tuple[bytes, int]:
    """normal.go:58-70; a key shorter than 13 bytes makes Go index out of range: reported as undecodable here"""
    if len(key) < 13:
        raise DecodeError("short")
    if key[:4] != MAGIC:
        raise DecodeError("magic")
    if key[len(key) - 9] != 0x24:
        raise DecodeError("split")
    return key[4 : len(key) - 9], struct.unpack(">Q", key[-8:])[0]


class Run:
    def __init__(self):
        self.emit: List[int] = []      # record indices handed to receiver.append, in order
        self.victims: List[Tuple[int, int]] = []  # (record index, class) of every delete call, in order
        self.count = 0
        self.examined = 0              # successful it.Next() calls
        self.limit_stop = False        # the loop ended because the receiver was full, not at io.EOF
        self.error: Optional[str] = None


def worker_run(keys: Sequence[bytes], vals: Sequence[bytes], start: bytes, end: bytes, revision: int, limit: int = 0,
               compact: bool = False, timeout_revision: int = 0, support_ttl: bool = True) -> Run:
    """keys ascending and unique (the storage.Iter contract); the iterator covers [start, end)"""
    out = Run()
    lo, hi = bisect_left(keys, start), bisect_left(keys, end)
    hi = max(hi, lo)
    collected = 0  # len(receiver.result); the compaction receiver (emptyResultReceiver) never fills

    def need_more() -> bool:
        return not (limit > 0 and not compact and collected >= limit)

    prev_uk, prev_rev, prev_val, prev_idx = b"", 0, b"", -1  # Go: nil slices, zero revision
    i = lo
    eof = False
    while need_more():
        if i >= hi:
            eof = True
            break
        idx = i
        i += 1
        out.examined += 1
        key, value = keys[idx], vals[idx]
        try:
            cur_uk, cur_rev = decode(key)
        except DecodeError:
            continue
        # compactIfExpired (scanner.go:566-591)
        if not support_ttl and timeout_revision != 0 and b"/events/" in cur_uk:
            if cur_rev == 0:
                if len(value) < 8:
                    out.error = "value[:8] out of range"  # Go panics
                    return out
                if struct.unpack(">Q", value[:8])[0] <= timeout_revision:
                    out.victims.append((idx, V_TTL_REVREC))
                    continue
            elif cur_rev <= timeout_revision:
                out.victims.append((idx, V_TTL_OBJECT))
                continue
        if cur_rev > revision:
            continue
        if cur_uk != prev_uk:
            if prev_rev > 0 and prev_val != TOMBSTONE:
                if not compact:
                    out.emit.append(prev_idx)
                    collected += 1
                out.count += 1
        elif compact and prev_rev > 0:
            out.victims.append((prev_idx, V_SUPERSEDED))
        if compact and value == TOMBSTONE:
            out.victims.append((idx, V_TOMBSTONE))
        if compact and cur_rev == 0 and len(value) == 9:
            if struct.unpack(">Q", value[:8])[0] > revision:
                continue  # without updating prev
            out.victims.append((idx, V_REVRECORD))
        prev_uk, prev_rev, prev_val, prev_idx = cur_uk, cur_rev, value, idx
    if not eof:
        # `err != io.EOF` with err == nil: the worker reports (0, nil); what the receiver holds is still merged
        out.limit_stop = True
        out.count = 0
        return out
    if prev_rev > 0 and prev_val != TOMBSTONE and need_more():
        if not compact:
            out.emit.append(prev_idx)
        out.count += 1
    return out


def fanout(event_keys: Sequence[bytes], event_revs: Sequence[int], batch_off: Sequence[int],
           prefixes: Sequence[bytes], min_revs: Sequence[int]):
    """WatcherHub.Stream (watcherhub.go:78-92) hands every batch to every watcher; each watcher applies
    filterByRevision (drops only the LEADING events below its revision, watch.go:152-159) and then filterByPrefix
    (watch.go:139-149) and forwards the batch if anything is left (watch.go:126-131).  Returns the per-watcher ordered
    event index lists and the total number of forwarded messages."""
    lists, messages = [], 0
    for prefix, rev in zip(prefixes, min_revs):
        got: List[int] = []
        for b in range(len(batch_off) - 1):
            lo, hi = int(batch_off[b]), int(batch_off[b + 1])
            i = lo
            while i < hi and event_revs[i] < rev:
                i += 1
            kept = [j for j in range(i, hi) if event_keys[j].startswith(prefix)]
            if kept:
                got.extend(kept)
                messages += 1
        lists.append(got)
    return lists, messages


def get(keys: Sequence[bytes], vals: Sequence[bytes], user_key: bytes, revision: int) -> Tuple[int, int]:
    """backend.get / getInternalVal (pkg/backend/range.go:81-121): a REVERSE iterator (start > end: seek to the largest
    key <= start, in range while key > end; pkg/storage/badger/iter.go:39-75) from EncodeObjectKey(key, revision) down
    to EncodeObjectKey(key, 0), limit 1.  Returns (record index, mod revision); index -1 = ErrKeyNotFound, -2 = the
    record is a tombstone (ErrKeyNotFound with a non-zero mod revision)."""
    from bisect import bisect_right

    if revision == 0:
        revision = 2**64 - 1
    start = MAGIC + user_key + b"$" + struct.pack(">Q", revision)
    end = MAGIC + user_key + b"$" + struct.pack(">Q", 0)
    pos = bisect_right(keys, start) - 1
    if pos < 0 or not keys[pos] > end:
        return -1, 0
    try:
        uk, mod_rev = decode(keys[pos])
    except DecodeError:
        return -1, 0  # Decode's error is ignored by the reference: userKey nil, modRev 0 -> not found
    if mod_rev == 0 or uk != user_key:
        return -1, 0
    if vals[pos] == TOMBSTONE:
        return -2, mod_rev
    return pos, mod_rev
