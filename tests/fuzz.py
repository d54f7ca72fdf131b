"""Adversarial random inputs for the parity tests (shared by CPU and GPU tests)."""
from __future__ import annotations

import random
import struct
from typing import List, Tuple

import numpy as np

from kubebrain_b200.packed import PackedEvents, PackedStore, PackedWatchers, Slab

MAGIC = b"\x57\xfb\x80\x8b"
TOMB = b"tombstone"


def fuzz_store(seed: int, n_keys: int = 60, max_rev: int = 60) -> PackedStore:
    """records with: shared prefixes, '$' and bytes below '$' inside user keys, user keys that are prefixes of each
    other, empty user key, missing revision records, 8/9/odd-length revision-record values, tombstones, 9-byte values
    that are not tombstones, undecodable keys (bad magic / bad split byte / shorter than 13 bytes)."""
    rng = random.Random(seed)
    alphabet = [b"a", b"b", b"/", b"$", b"#", b"\x00", b"\xff", b"events", b"/events/", b"zz"]
    user_keys = set()
    while len(user_keys) < n_keys:
        parts = rng.randint(0, 6)
        uk = b"".join(rng.choice(alphabet) for _ in range(parts))
        if rng.random() < 0.3 and user_keys:
            uk = rng.choice(sorted(user_keys)) + rng.choice(alphabet)  # extensions of existing keys
        if rng.random() < 0.1:
            uk = uk + b"x" * rng.randint(20, 300)  # long keys (several 16-byte chunks, > staging stride)
        user_keys.add(uk)
    items = {}
    for uk in sorted(user_keys):  # sorted: set order depends on the per-process hash seed
        revs = sorted(rng.sample(range(1, max_rev), rng.randint(0, 5)))
        if rng.random() < 0.85:
            r = rng.random()
            latest = revs[-1] if revs else rng.randint(1, max_rev)
            if r < 0.5:
                val = struct.pack(">Q", latest)
            elif r < 0.8:
                val = struct.pack(">Q", rng.choice([latest, rng.randint(1, max_rev)])) + b"\x00"
            elif r < 0.9:
                val = TOMB  # a 9-byte revision-record value that happens to be the tombstone literal
            else:
                val = bytes(rng.randrange(256) for _ in range(rng.choice([0, 3, 8, 9, 12])))
            if len(val) < 8 and b"/events/" in uk:
                val = struct.pack(">Q", latest)  # Go would panic on value[:8] in compactIfExpired
            items[MAGIC + uk + b"$" + b"\x00" * 8] = val
        for i, rev in enumerate(revs):
            r = rng.random()
            if r < 0.2:
                val = TOMB
            elif r < 0.3:
                val = bytes(rng.randrange(256) for _ in range(9))
            elif r < 0.35:
                val = b""
            else:
                val = bytes(rng.randrange(256) for _ in range(rng.randint(1, 40)))
            items[MAGIC + uk + b"$" + struct.pack(">Q", rev)] = val
    # undecodable records
    for _ in range(rng.randint(0, 6)):
        kind = rng.randint(0, 3)
        uk = rng.choice(sorted(user_keys))
        if kind == 0:
            k = b"\x57\xfb\x80\x8c" + uk + b"$" + struct.pack(">Q", rng.randint(0, max_rev))
        elif kind == 1:
            k = MAGIC + uk + b"%" + struct.pack(">Q", rng.randint(0, max_rev))
        elif kind == 2:
            k = MAGIC + bytes(rng.randrange(256) for _ in range(rng.randint(1, 8)))
        else:
            k = bytes(rng.randrange(256) for _ in range(rng.randint(1, 30)))
        items.setdefault(k, b"junk")
    return PackedStore.from_items(list(items.items()))


def fuzz_bounds(store: PackedStore, seed: int, n: int = 8) -> List[Tuple[bytes, bytes]]:
    rng = random.Random(seed)
    keys = store.keys.tolist()
    out = [(b"\x00", b"\xff" * 4), (MAGIC, MAGIC + b"\xff" * 8)]
    for _ in range(n):
        a, b = rng.choice(keys), rng.choice(keys)
        if rng.random() < 0.5:
            a = a[: rng.randint(1, max(len(a), 1))]
        if rng.random() < 0.5:
            b = b[: rng.randint(1, max(len(b), 1))] + b"\xff"
        if a > b:
            a, b = b, a
        out.append((a, b))
    out.append((keys[0], keys[0]))  # empty
    out.append((keys[-1], keys[-1] + b"\x00"))  # last record only
    return out


def fuzz_events(seed: int, n: int = 200, monotone: bool = True) -> PackedEvents:
    rng = random.Random(seed)
    alphabet = [b"a", b"b", b"/", b"/ns-1/", b"/ns-22/", b"pods", b"x" * 20]
    keys, revs = [], []
    rev = rng.randint(1, 50)
    for _ in range(n):
        keys.append(b"".join(rng.choice(alphabet) for _ in range(rng.randint(0, 6))))
        if monotone:
            rev += rng.randint(0, 2)
            revs.append(rev)
        else:
            revs.append(rng.randint(1, 100))
    cuts = sorted(set([0, n] + [rng.randint(0, n) for _ in range(rng.randint(0, 8))]))
    return PackedEvents(Slab.from_list(keys), np.array(revs, dtype=np.uint64), np.array(cuts, dtype=np.uint64))


def fuzz_watchers(ev: PackedEvents, seed: int, n: int = 40) -> PackedWatchers:
    rng = random.Random(seed)
    keys = ev.keys.tolist() or [b""]
    pref, mr = [], []
    maxrev = int(ev.rev.max()) if ev.n else 10
    for _ in range(n):
        r = rng.random()
        k = rng.choice(keys)
        if r < 0.1:
            p = b""
        elif r < 0.7:
            p = k[: rng.randint(0, len(k))]
        elif r < 0.8:
            p = k + b"more"  # longer than any key it could match
        elif r < 0.9 and pref:
            p = rng.choice(pref)  # duplicate prefix (several watchers in one group)
        else:
            p = bytes(rng.randrange(256) for _ in range(rng.randint(1, 5)))
        pref.append(p)
        mr.append(rng.choice([0, 0, rng.randint(0, maxrev + 2)]))
    return PackedWatchers(Slab.from_list(pref), np.array(mr, dtype=np.uint64))
