"""Host logic of the N > 1 deployment: one process per GPU, every rank owns the namespaces that hash to it.

The data path has no collective: MVCC dedup is per user key and a user key never spans two shards (the reference
keeps the same invariant between its scan workers, pkg/backend/scanner/scanner.go:202-225), watcher predicates are
per event.  What remains on the host is routing and merging:

  * ``shard_of_key`` / ``owner_of_prefix``: where a key lives, and whether a List / Watch prefix touches one shard
    (prefix at or below namespace level) or all of them;
  * ``merge_list_runs``: the per-shard answers of a broad List are sorted runs of *whole namespaces*; the global answer
    is their merge by key followed by ``limit`` (commonResultReceiver semantics, scanner/receiver.go:82-103, and
    backend.List's More flag, pkg/backend/range.go:150-190);
  * ``merge_watch_streams``: a watcher whose prefix is shorter than the partition prefix is registered on every shard;
    its per-shard streams (each revision ordered) are merged by event revision;
  * the only exchange between ranks is the committed-revision cursor (``kb_cursor_allgather``): readable revision =
    min over shards (pkg/backend/tso/tso.go:47-49, pkg/server/service/revision/revision.go:219-259).
"""
from __future__ import annotations

import heapq
from typing import Iterable, List, Optional, Sequence, Tuple

FNV_OFFSET, FNV_PRIME, MASK64 = 14695981039346656037, 1099511628211, (1 << 64) - 1


def fnv1a64(data: bytes) -> int:
    h = FNV_OFFSET
    for b in data:
        h = ((h ^ b) * FNV_PRIME) & MASK64
    return h


def partition_token(user_key: bytes) -> bytes:
    """the path segment a key is sharded by: the namespace of ``/registry/<resource>/<namespace>/<name>``, the
    resource of a cluster-scoped ``/registry/<resource>/<name>``"""
    parts = user_key.split(b"/")
    if len(parts) >= 5:  # ['', 'registry', resource, namespace, name...]
        return parts[3]
    return parts[2] if len(parts) >= 3 else user_key


def shard_of_key(user_key: bytes, world: int) -> int:
    return fnv1a64(partition_token(user_key)) % world if world > 1 else 0


def owner_of_prefix(prefix: bytes, world: int) -> Optional[int]:
    """the single shard a List / Watch prefix can match, or None when it has to go to every shard (the prefix stops
    before the end of the namespace segment)"""
    if world <= 1:
        return 0
    parts = prefix.split(b"/")
    if len(parts) >= 5:  # the namespace segment is complete (terminated by '/')
        return fnv1a64(parts[3]) % world
    return None


def merge_list_runs(runs: Sequence[Sequence[tuple]], limit: int = 0) -> Tuple[List[tuple], bool]:
    """runs: per-shard answers, each sorted by key, entries (key, value, revision).  Returns (kvs, more): the merge,
    cut to ``limit`` (limit <= 0: unlimited); ``more`` says the cut dropped something."""
    merged = heapq.merge(*runs, key=lambda kv: kv[0])
    if limit <= 0:
        return list(merged), False
    out = []
    for kv in merged:
        if len(out) == limit:
            return out, True
        out.append(kv)
    return out, False


def merge_watch_streams(streams: Sequence[Iterable[tuple]]) -> List[tuple]:
    """streams: per-shard deliveries of one watcher, each ordered by revision, entries (revision, ...).  Revisions are
    unique across shards (one tso), so the merge is total."""
    return list(heapq.merge(*streams, key=lambda ev: ev[0]))


def readable_revision(cursors: Sequence[int]) -> int:
    """what ``kb_cursor_allgather`` computes on the device: the newest revision every shard has committed"""
    return min(cursors)
