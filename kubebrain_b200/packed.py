"""Packed byte-array containers shared by the host mirror, the synthetic generator and the tests.

A *slab* is a list of byte strings stored back to back (``data``) with ``n+1`` uint64 offsets
(``off``).  This is the wire layout of the C ABI in ``include/kb_b200.h`` (``kb_load_sorted``,
``kb_events``): plain pointers and sizes, no torch types.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Sequence

import numpy as np


@dataclass
class Slab:
    data: np.ndarray  # uint8
    off: np.ndarray  # uint64, n+1

    @property
    def n(self) -> int:
        return int(self.off.shape[0] - 1)

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i: int) -> bytes:
        if i < 0:
            i += self.n
        return self.data[int(self.off[i]) : int(self.off[i + 1])].tobytes()

    def lengths(self) -> np.ndarray:
        return np.diff(self.off).astype(np.int64)

    def tolist(self) -> List[bytes]:
        return [self[i] for i in range(self.n)]

    @staticmethod
    def from_list(items: Sequence[bytes]) -> "Slab":
        lens = np.fromiter((len(x) for x in items), dtype=np.uint64, count=len(items))
        off = np.zeros(len(items) + 1, dtype=np.uint64)
        np.cumsum(lens, out=off[1:])
        data = np.frombuffer(b"".join(items), dtype=np.uint8).copy() if len(items) else np.zeros(0, np.uint8)
        return Slab(np.ascontiguousarray(data), off)

    @staticmethod
    def from_fixed(matrix: np.ndarray) -> "Slab":
        """rows of a 2-D uint8 matrix, all the same length"""
        n, w = matrix.shape
        off = (np.arange(n + 1, dtype=np.uint64) * np.uint64(w)).astype(np.uint64)
        return Slab(np.ascontiguousarray(matrix.reshape(-1)), off)

    def take(self, idx: Iterable[int]) -> "Slab":
        return Slab.from_list([self[int(i)] for i in idx])


@dataclass
class PackedStore:
    """A sorted snapshot of the storage engine: unique internal keys ascending (bytes.Compare order),
    as ``storage.Iter`` yields them (reference pkg/storage/badger/iter.go:39-83)."""

    keys: Slab
    vals: Slab

    @property
    def n(self) -> int:
        return self.keys.n

    @staticmethod
    def from_items(items: Sequence[tuple]) -> "PackedStore":
        items = sorted(items, key=lambda kv: kv[0])
        return PackedStore(Slab.from_list([k for k, _ in items]), Slab.from_list([v for _, v in items]))


@dataclass
class PackedEvents:
    """A revision-ordered slab of watch events (reference backend.go:237-256 builds them):
    ``keys`` = Event.Kv.Key (user keys), ``rev`` = Event.Revision, ``batch_off`` = the boundaries of
    the <=300-event batches the collector pushes to the hub (backend.go:41,262-266)."""

    keys: Slab
    rev: np.ndarray  # uint64
    batch_off: np.ndarray  # uint64, n_batches+1

    @property
    def n(self) -> int:
        return self.keys.n


@dataclass
class PackedWatchers:
    prefixes: Slab
    min_rev: np.ndarray  # uint64

    @property
    def n(self) -> int:
        return self.prefixes.n
