"""kubebrain_b200 -- B200-native MVCC range-scan / compaction-sweep / watch fan-out for KubeBrain.

The compute path is hand-written CUDA for sm_100a in ``csrc/`` behind the C ABI of ``include/kb_b200.h``
(``libkbb200.so``); this package is the host-side mirror of the reference's Go interfaces for the path
(``coder.Coder``, ``scanner.Scanner``, the read/watch half of ``backend.Backend``).  There is no CPU fallback:
importing the package is cheap, but every operation loads the shared library and raises if it (or a CUDA
device) is missing.
"""
from .packed import PackedEvents, PackedStore, PackedWatchers, Slab  # noqa: F401

__all__ = ["PackedEvents", "PackedStore", "PackedWatchers", "Slab"]
