"""Deterministic synthetic Kubernetes-shaped inputs.

One generator feeds the CPU oracle, the CUDA path and the CPU baseline timing, so every leg of a
comparison sees identical bytes.  All randomness comes from splitmix64 seeded with
0x6B75626562726169 ("kubebrai") xor a config id.

* user key  = ``/registry/<res>/ns-%05d/<name>`` right-padded with ``[a-z0-9]`` to exactly ``Lu`` bytes (only
  bytes greater than ``'$'``: the contiguity precondition of reference coder/normal.go:29);
* records   = per object one revision record (rev 0, value BE64(latest rev) [+0x00 when deleted]) followed by its
  versions at strictly increasing revisions, interleaved round-robin across objects
  (formats: reference creator/naive.go:53-105, txn.go:145-190,249-265);
* values    = ``Lv`` PRNG bytes keyed by (object, version); a deleted object's last version is the literal
  ``tombstone`` (reference backend/util.go:28).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

from .packed import PackedEvents, PackedStore, PackedWatchers, Slab

SEED = 0x6B75626562726169
MAGIC = np.array([0x57, 0xFB, 0x80, 0x8B], dtype=np.uint8)
TOMBSTONE = np.frombuffer(b"tombstone", dtype=np.uint8)
_ALNUM = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz0123456789", dtype=np.uint8)
RESOURCES = [(b"pods", 0.50), (b"configmaps", 0.15), (b"secrets", 0.10), (b"services", 0.10),
             (b"deployments", 0.10), (b"events", 0.05)]

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """one splitmix64 output per input counter value (vectorised, wraps mod 2^64)"""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def _stream(seed: int, n: int, salt: int = 0) -> np.ndarray:
    base = np.uint64((seed ^ (salt * 0x9E3779B97F4A7C15)) & 0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        return splitmix64(base + np.arange(n, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15))


def _alnum_matrix(seed: int, rows: int, cols: int, salt: int) -> np.ndarray:
    words = (cols + 7) // 8
    r = _stream(seed, rows * words, salt).reshape(rows, words)
    b = r.view(np.uint8).reshape(rows, words * 8)[:, :cols]
    return _ALNUM[b % 36]


def _ns_digits(ns: np.ndarray) -> np.ndarray:
    """'ns-%05d' as a (n, 8) uint8 matrix"""
    out = np.empty((ns.shape[0], 8), dtype=np.uint8)
    out[:, 0:3] = np.frombuffer(b"ns-", dtype=np.uint8)
    v = ns.astype(np.int64)
    for d in range(5):
        out[:, 7 - d] = (v % 10 + 48).astype(np.uint8)
        v //= 10
    return out


def _user_keys(seed: int, n: int, lu: int, n_namespaces: int, res_idx: np.ndarray, ns: np.ndarray, salt: int) -> np.ndarray:
    """(n, lu) uint8 user keys; lu must leave room for at least 4 name bytes after the longest prefix"""
    keys = _alnum_matrix(seed, n, lu, salt)
    for ri, (res, _) in enumerate(RESOURCES):
        sel = np.nonzero(res_idx == ri)[0]
        if sel.size == 0:
            continue
        pre = b"/registry/" + res + b"/"
        p = len(pre)
        assert lu >= p + 9 + 4, "Lu too small for /registry/<res>/ns-%05d/<name>"
        keys[sel, :p] = np.frombuffer(pre, dtype=np.uint8)
        keys[sel, p : p + 8] = _ns_digits(ns[sel])
        keys[sel, p + 8] = ord("/")
    return keys


def _pick_resources(seed: int, n: int, salt: int, only: Optional[bytes] = None) -> np.ndarray:
    if only is not None:
        idx = [i for i, (r, _) in enumerate(RESOURCES) if r == only][0]
        return np.full(n, idx, dtype=np.int64)
    u = (_stream(seed, n, salt) >> np.uint64(11)).astype(np.float64) / float(1 << 53)
    edges = np.cumsum([w for _, w in RESOURCES])
    return np.minimum(np.searchsorted(edges, u, side="right"), len(RESOURCES) - 1).astype(np.int64)


@dataclass
class StoreMeta:
    n_objects: int
    versions: int
    lu: int
    lv: int
    first_rev: int
    last_rev: int
    read_rev: int  # the 90th-percentile revision
    n_tombstoned: int


def fnv1a64_rows(mat: np.ndarray) -> np.ndarray:
    """FNV-1a 64 of every row of a uint8 matrix"""
    h = np.full(mat.shape[0], 14695981039346656037, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for j in range(mat.shape[1]):
            h = (h ^ mat[:, j].astype(np.uint64)) * np.uint64(1099511628211)
    return h


def ns_shard(ns: np.ndarray, world: int) -> np.ndarray:
    """shard of a namespace = fnv1a64("ns-%05d") mod G (hash of the object-key prefix)"""
    if world <= 1:
        return np.zeros(ns.shape[0], dtype=np.int64)
    return (fnv1a64_rows(_ns_digits(ns)) % np.uint64(world)).astype(np.int64)


def gen_store(n_objects: int, versions: int, lu: int, lv: int, n_namespaces: int, config_id: int = 2,
              tomb_frac: float = 0.05, only_resource: Optional[bytes] = None,
              first_rev: int = 1000, shard: Optional[Tuple[int, int]] = None) -> Tuple[PackedStore, StoreMeta]:
    """n_objects * (1 + versions) records sorted by internal key.  shard=(rank, world) keeps only the objects
    whose namespace hashes to `rank` (revisions stay those of the global data set)."""
    seed = SEED ^ config_id
    res_idx = _pick_resources(seed, n_objects, 1, only_resource)
    ns = (_stream(seed, n_objects, 2) % np.uint64(n_namespaces)).astype(np.int64)
    uk = _user_keys(seed, n_objects, lu, n_namespaces, res_idx, ns, 3)
    n_global = n_objects
    global_idx = np.arange(n_objects, dtype=np.uint64)
    tomb_all = (_stream(seed, n_objects, 4) % np.uint64(10000)) < np.uint64(int(tomb_frac * 10000))
    if shard is not None and shard[1] > 1:
        mine = np.nonzero(ns_shard(ns, shard[1]) == shard[0])[0]
        uk, global_idx, tomb_all = uk[mine], global_idx[mine], tomb_all[mine]
        n_objects = int(mine.shape[0])
    # unique + sorted user keys (fixed length, so user-key order == internal-key order)
    order = np.argsort(uk.view(f"S{lu}").reshape(-1), kind="stable")
    uk = uk[order]
    dup = np.nonzero((uk[1:] == uk[:-1]).all(axis=1))[0]
    assert dup.size == 0, "synthetic user keys collided; change the seed"
    creation = global_idx[order]  # creation index (in the global data set) of the object now at sorted position i

    per = versions + 1
    n = n_objects * per
    lk = lu + 13
    keys = np.empty((n_objects, per, lk), dtype=np.uint8)
    keys[:, :, 0:4] = MAGIC
    keys[:, :, 4 : 4 + lu] = uk[:, None, :]
    keys[:, :, 4 + lu] = 0x24
    # revisions: version v of the object created c-th gets first_rev + v*n_objects + c + 1
    vidx = np.arange(versions, dtype=np.uint64)
    revs = np.uint64(first_rev) + vidx[None, :] * np.uint64(n_global) + creation[:, None] + np.uint64(1)
    rev_all = np.zeros((n_objects, per), dtype=np.uint64)
    rev_all[:, 1:] = revs
    keys[:, :, 5 + lu :] = rev_all.astype(">u8").view(np.uint8).reshape(n_objects, per, 8)
    key_slab = Slab.from_fixed(keys.reshape(n, lk))

    tomb = tomb_all[order]
    # Values.  Per object: [revision record: BE64(latest) (+0x00 when deleted)] [v1] ... [v_last or "tombstone"].
    # Payload bytes come from a PCG64 raw stream seeded by splitmix64 (bit-stable across numpy versions).
    # Every object is first generated as a live row of A = 8 + versions*lv bytes; runs of consecutive live
    # objects are then block-copied, and the (few) deleted objects are patched one by one.
    assert lv >= 9
    A = 8 + versions * lv
    words = (A + 7) // 8
    pcg_seed = int(splitmix64(np.array([seed & 0xFFFFFFFFFFFFFFFF], dtype=np.uint64))[0])
    M2 = np.random.PCG64(pcg_seed).random_raw(n_objects * words).view(np.uint8).reshape(n_objects, words * 8)[:, :A]
    M2[:, 0:8] = revs[:, -1].astype(">u8").view(np.uint8).reshape(n_objects, 8)
    vlen = np.full((n_objects, per), lv, dtype=np.uint64)
    vlen[:, 0] = np.where(tomb, 9, 8)
    vlen[tomb, per - 1] = 9
    voff = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(vlen.reshape(-1), out=voff[1:])
    ooff = voff[::per].astype(np.int64)  # byte offset of every object's first value (+ total at the end)
    vals = np.empty(int(voff[-1]), dtype=np.uint8)
    tomb_idx = np.nonzero(tomb)[0]
    run_start = 0
    body = 8 + (versions - 1) * lv
    for ti in list(tomb_idx) + [n_objects]:
        ti = int(ti)
        if ti > run_start:
            vals[ooff[run_start] : ooff[ti]] = M2[run_start:ti].reshape(-1)
        if ti < n_objects:
            o = int(ooff[ti])
            vals[o : o + 8] = M2[ti, 0:8]
            vals[o + 8] = 0
            vals[o + 9 : o + 9 + body - 8] = M2[ti, 8:body]
            vals[o + 1 + body : o + 1 + body + 9] = TOMBSTONE
        run_start = ti + 1
    del M2
    val_slab = Slab(vals, voff)

    last_rev = first_rev + versions * n_global
    read_rev = first_rev + int(0.9 * versions * n_global)
    meta = StoreMeta(n_objects, versions, lu, lv, first_rev, last_rev, read_rev, int(tomb.sum()))
    return PackedStore(key_slab, val_slab), meta


def gen_natural_store(n_objects: int, lv: int, n_namespaces: int, config_id: int = 1) -> Tuple[PackedStore, StoreMeta]:
    """config 1: /registry/pods/ns-%03d-ish natural-length keys (variable Lu ~ 40..56), one version each."""
    seed = SEED ^ config_id
    ns = (_stream(seed, n_objects, 2) % np.uint64(n_namespaces)).astype(np.int64)
    extra = (_stream(seed, n_objects, 5) % np.uint64(17)).astype(np.int64)  # name length 12..28
    names = _alnum_matrix(seed, n_objects, 28, 3)
    items = {}
    for i in range(n_objects):
        uk = b"/registry/pods/ns-%05d/" % int(ns[i]) + names[i, : 12 + int(extra[i])].tobytes()
        items[uk] = i
    uks = sorted(items)
    recs = []
    first_rev = 1000
    vals_rng = _stream(seed, n_objects * ((lv + 7) // 8), 6).view(np.uint8)
    for j, uk in enumerate(uks):
        c = items[uk]
        rev = first_rev + c + 1
        revb = int(rev).to_bytes(8, "big")
        recs.append((MAGIC.tobytes() + uk + b"$" + b"\x00" * 8, revb))
        w = (lv + 7) // 8 * 8
        recs.append((MAGIC.tobytes() + uk + b"$" + revb, vals_rng[c * w : c * w + lv].tobytes()))
    store = PackedStore(Slab.from_list([k for k, _ in recs]), Slab.from_list([v for _, v in recs]))
    meta = StoreMeta(len(uks), 1, 0, lv, first_rev, first_rev + n_objects, first_rev + n_objects, 0)
    return store, meta


def gen_events(n_events: int, lu: int, n_namespaces: int, start_rev: int, config_id: int = 3,
               batch: int = 300) -> PackedEvents:
    """a revision-ordered burst on /registry/pods/<ns>/... keys, consecutive revisions, batches of <= 300"""
    seed = SEED ^ config_id
    res_idx = _pick_resources(seed, n_events, 11, b"pods")
    ns = (_stream(seed, n_events, 12) % np.uint64(n_namespaces)).astype(np.int64)
    uk = _user_keys(seed, n_events, lu, n_namespaces, res_idx, ns, 13)
    rev = np.uint64(start_rev) + np.arange(n_events, dtype=np.uint64)
    bo = np.array(list(range(0, n_events, batch)) + [n_events], dtype=np.uint64)
    return PackedEvents(Slab.from_fixed(uk), rev, bo)


def gen_watchers(n_ns_watchers: int, n_cluster: int, burst_start: int, burst_mid: int, config_id: int = 3,
                 ns_offset: int = 0) -> PackedWatchers:
    """namespace watchers /registry/pods/ns-%05d/ for namespaces [ns_offset, ns_offset+n) + cluster-wide
    /registry/pods/ ; min_rev = burst start for 90 %, burst midpoint for 10 %"""
    seed = SEED ^ config_id
    pre = b"/registry/pods/"
    n = n_ns_watchers + n_cluster
    mat = np.empty((n_ns_watchers, len(pre) + 9), dtype=np.uint8)
    mat[:, : len(pre)] = np.frombuffer(pre, dtype=np.uint8)
    mat[:, len(pre) : len(pre) + 8] = _ns_digits(np.arange(n_ns_watchers, dtype=np.int64) + ns_offset)
    mat[:, len(pre) + 8] = ord("/")
    lens = np.concatenate([np.full(n_ns_watchers, mat.shape[1], np.uint64), np.full(n_cluster, len(pre), np.uint64)])
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    data = np.concatenate([mat.reshape(-1), np.tile(np.frombuffer(pre, dtype=np.uint8), n_cluster)])
    late = (_stream(seed, n, 21) % np.uint64(10)) == np.uint64(0)
    min_rev = np.where(late, np.uint64(burst_mid), np.uint64(burst_start)).astype(np.uint64)
    return PackedWatchers(Slab(np.ascontiguousarray(data), off), min_rev)
