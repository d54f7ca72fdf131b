"""ctypes binding of the CPU ORACLE (oracle/libkboracle.so) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package (kubebrain_b200) never does.
Parity status: pinned by the reference's golden vectors (tests/test_oracle_golden.py, G1..G7);
unpinned for bulk inputs (the Go reference cannot be built in this image).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from kubebrain_b200.packed import PackedEvents, PackedStore, PackedWatchers, Slab

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libkboracle.so")

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


class KoStore(C.Structure):
    _fields_ = [("keys", u8p), ("koff", u64p), ("vals", u8p), ("voff", u64p), ("n", C.c_uint64)]


class KoWorkerCfg(C.Structure):
    _fields_ = [
        ("read_rev", C.c_uint64),
        ("limit", C.c_int64),
        ("compact", C.c_int),
        ("timeout_rev", C.c_uint64),
        ("support_ttl", C.c_int),
        ("collect", C.c_int),
    ]


class KoResult(C.Structure):
    _fields_ = [
        ("emit", u64p), ("n_emit", C.c_uint64), ("cap_emit", C.c_uint64),
        ("victim", u64p), ("vclass", u8p), ("n_victim", C.c_uint64), ("cap_victim", C.c_uint64),
        ("count", C.c_int), ("limit_stop", C.c_int), ("examined", C.c_uint64), ("val_size", C.c_uint64),
    ]


class KoFindRet(C.Structure):
    _fields_ = [
        ("empty", C.c_int), ("high", C.c_int), ("low", C.c_int),
        ("newest_rev", C.c_uint64), ("oldest_rev", C.c_uint64), ("n_events", C.c_uint64),
    ]


class KoEvents(C.Structure):
    _fields_ = [
        ("keys", u8p), ("koff", u64p), ("rev", u64p), ("n", C.c_uint64),
        ("batch_off", u64p), ("n_batches", C.c_uint64),
    ]


class KoWatchers(C.Structure):
    _fields_ = [("prefixes", u8p), ("poff", u64p), ("min_rev", u64p), ("n", C.c_uint64)]


class KoFanout(C.Structure):
    _fields_ = [("start", u64p), ("event_idx", u32p), ("n_deliveries", C.c_uint64), ("n_messages", C.c_uint64)]


class KoWatchReg(C.Structure):
    _fields_ = [("mode", C.c_int), ("live_rev", C.c_uint64), ("n_catchup", C.c_uint64), ("err_rev", C.c_uint64)]


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, "kb_oracle.c"), os.path.join(_HERE, "kb_oracle.h")]
    if force or not os.path.exists(_LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "libkboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ko_encode_object_key.restype = C.c_size_t
        L.ko_encode_object_key.argtypes = [C.c_char_p, C.c_size_t, C.c_uint64, u8p]
        L.ko_decode.restype = C.c_int
        L.ko_decode.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), u64p]
        L.ko_parse_revision.restype = C.c_int
        L.ko_parse_revision.argtypes = [C.c_char_p, C.c_size_t, u64p, C.POINTER(C.c_int)]
        L.ko_prefix_end.restype = C.c_size_t
        L.ko_prefix_end.argtypes = [C.c_char_p, C.c_size_t, u8p]
        L.ko_lower_bound.restype = C.c_uint64
        L.ko_lower_bound.argtypes = [C.POINTER(KoStore), C.c_char_p, C.c_size_t]
        L.ko_result_init.argtypes = [C.POINTER(KoResult)]
        L.ko_result_free.argtypes = [C.POINTER(KoResult)]
        L.ko_worker_run.restype = C.c_int
        L.ko_worker_run.argtypes = [C.POINTER(KoStore), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                    C.POINTER(KoWorkerCfg), C.POINTER(KoResult)]
        L.ko_adjust_partition_borders.restype = C.c_int
        L.ko_adjust_partition_borders.argtypes = [u8p, u64p, C.c_uint64, u8p, u64p]
        L.ko_scan.restype = C.c_int
        L.ko_scan.argtypes = [C.POINTER(KoStore), u8p, u64p, C.c_uint64, C.POINTER(KoWorkerCfg), C.c_int,
                              C.c_uint64, C.c_int, C.POINTER(KoResult), C.POINTER(C.c_int)]
        L.ko_range.restype = C.c_int
        L.ko_range.argtypes = [C.POINTER(KoStore), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_uint64,
                               C.c_int64, C.c_int, C.c_uint64, C.POINTER(KoResult)]
        L.ko_get.restype = C.c_int64
        L.ko_get.argtypes = [C.POINTER(KoStore), C.c_char_p, C.c_size_t, C.c_uint64, u64p]
        L.ko_compact_borders.restype = C.c_int
        L.ko_compact_borders.argtypes = [u8p, u64p, C.c_uint64, u8p, u64p]
        L.ko_ring_new.restype = C.c_void_p
        L.ko_ring_new.argtypes = [C.c_int64]
        L.ko_ring_free.argtypes = [C.c_void_p]
        L.ko_ring_reset.argtypes = [C.c_void_p]
        L.ko_ring_add.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        L.ko_ring_find.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(KoFindRet), u64p, u64p]
        L.ko_fanout_free.argtypes = [C.POINTER(KoFanout)]
        L.ko_fanout_run.restype = C.c_int
        L.ko_fanout_run.argtypes = [C.POINTER(KoEvents), C.POINTER(KoWatchers), C.c_int, C.c_int, C.POINTER(KoFanout)]
        L.ko_watch_register.restype = C.c_int
        L.ko_watch_register.argtypes = [C.c_void_p, C.POINTER(KoEvents), C.c_char_p, C.c_size_t, C.c_uint64,
                                        C.c_uint64, C.POINTER(KoWatchReg), u64p, C.c_uint64]
        L.ko_catchup_chunks.restype = C.c_uint64
        L.ko_catchup_chunks.argtypes = [C.c_uint64, u64p, C.c_uint64]
        L.ko_bench_scan.restype = C.c_int64
        L.ko_bench_scan.argtypes = [C.POINTER(KoStore), C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_uint64,
                                    C.c_int64, C.c_int, C.c_int, u64p, u64p]
        _lib = L
    return _lib


def _p8(a: np.ndarray):
    return a.ctypes.data_as(u8p)


def _p64(a: np.ndarray):
    return a.ctypes.data_as(u64p)


def _slab_ptrs(s: Slab):
    data = s.data if s.data.size else np.zeros(1, np.uint8)
    return data, _p8(data), _p64(s.off)


# ---- coder ---------------------------------------------------------------------------------------

def encode_object_key(user_key: bytes, rev: int) -> bytes:
    out = (C.c_uint8 * (len(user_key) + 13))()
    n = lib().ko_encode_object_key(user_key, len(user_key), rev, out)
    return bytes(out[:n])


def decode(internal_key: bytes) -> Tuple[Optional[bytes], int, int]:
    """returns (user_key | None, revision, err)"""
    off, ln, rev = C.c_size_t(), C.c_size_t(), C.c_uint64()
    rc = lib().ko_decode(internal_key, len(internal_key), C.byref(off), C.byref(ln), C.byref(rev))
    if rc != 0:
        return None, 0, rc
    return internal_key[off.value : off.value + ln.value], rev.value, 0


def parse_revision(val: bytes) -> Tuple[int, bool, int]:
    rev, dele = C.c_uint64(), C.c_int()
    rc = lib().ko_parse_revision(val, len(val), C.byref(rev), C.byref(dele))
    return rev.value, bool(dele.value), rc


def prefix_end(prefix: bytes) -> bytes:
    out = (C.c_uint8 * max(len(prefix), 1))()
    n = lib().ko_prefix_end(prefix, len(prefix), out)
    return bytes(out[:n])


# ---- store / scan -----------------------------------------------------------------------------------

class OracleStore:
    def __init__(self, store: PackedStore):
        self.store = store
        self._kd, kp, kop = _slab_ptrs(store.keys)
        self._vd, vp, vop = _slab_ptrs(store.vals)
        self.c = KoStore(kp, kop, vp, vop, store.n)

    def lower_bound(self, key: bytes) -> int:
        return int(lib().ko_lower_bound(C.byref(self.c), key, len(key)))


@dataclass
class ScanResult:
    rc: int
    emit: np.ndarray  # record indices of the emitted kvs, in order
    victims: np.ndarray  # record indices of delete calls, in order
    vclass: np.ndarray
    count: int
    limit_stop: bool
    examined: int

    def kvs(self, store: PackedStore) -> List[Tuple[bytes, bytes, int]]:
        out = []
        for i in self.emit:
            uk, rev, _ = decode(store.keys[int(i)])
            out.append((uk, store.vals[int(i)], rev))
        return out


def _take_result(r: KoResult, rc: int) -> ScanResult:
    emit = np.ctypeslib.as_array(r.emit, shape=(r.n_emit,)).copy() if r.n_emit else np.zeros(0, np.uint64)
    vic = np.ctypeslib.as_array(r.victim, shape=(r.n_victim,)).copy() if r.n_victim else np.zeros(0, np.uint64)
    vcl = np.ctypeslib.as_array(r.vclass, shape=(r.n_victim,)).copy() if r.n_victim else np.zeros(0, np.uint8)
    res = ScanResult(rc, emit, vic, vcl, int(r.count), bool(r.limit_stop), int(r.examined))
    lib().ko_result_free(C.byref(r))
    return res


def worker_run(st: OracleStore, start: bytes, end: bytes, read_rev: int, limit: int = 0, compact: bool = False,
               timeout_rev: int = 0, support_ttl: bool = True, collect: bool = True) -> ScanResult:
    cfg = KoWorkerCfg(read_rev, limit, int(compact), timeout_rev, int(support_ttl), int(collect))
    r = KoResult()
    lib().ko_result_init(C.byref(r))
    rc = lib().ko_worker_run(C.byref(st.c), start, len(start), end, len(end), C.byref(cfg), C.byref(r))
    return _take_result(r, rc)


def scan(st: OracleStore, borders: Sequence[bytes], read_rev: int, limit: int = 0, compact: bool = False,
         timeout_rev: int = 0, support_ttl: bool = True, collect: bool = True,
         compact_rev: Optional[int] = None, threads: int = 1) -> ScanResult:
    sl = Slab.from_list(list(borders))
    data, bp, bop = _slab_ptrs(sl)
    cfg = KoWorkerCfg(read_rev, limit, int(compact), timeout_rev, int(support_ttl), int(collect))
    r = KoResult()
    lib().ko_result_init(C.byref(r))
    total = C.c_int()
    rc = lib().ko_scan(C.byref(st.c), bp, bop, sl.n, C.byref(cfg), int(compact_rev is not None),
                       compact_rev or 0, threads, C.byref(r), C.byref(total))
    return _take_result(r, rc)


def range_(st: OracleStore, start: bytes, end: bytes, read_rev: int, limit: int = 0,
           compact_rev: Optional[int] = None) -> ScanResult:
    r = KoResult()
    lib().ko_result_init(C.byref(r))
    rc = lib().ko_range(C.byref(st.c), start, len(start), end, len(end), read_rev, limit,
                        int(compact_rev is not None), compact_rev or 0, C.byref(r))
    return _take_result(r, rc)


def get(st: OracleStore, user_key: bytes, rev: int) -> Tuple[int, int]:
    mod = C.c_uint64()
    idx = lib().ko_get(C.byref(st.c), user_key, len(user_key), rev, C.byref(mod))
    return int(idx), int(mod.value)


def adjust_partition_borders(borders: Sequence[bytes]) -> List[bytes]:
    sl = Slab.from_list(list(borders))
    data, bp, bop = _slab_ptrs(sl)
    out = np.zeros(int(sl.off[-1]) + 16 * sl.n + 16, np.uint8)
    out_off = np.zeros(sl.n + 1, np.uint64)
    lib().ko_adjust_partition_borders(bp, bop, sl.n, _p8(out), _p64(out_off))
    return Slab(out, out_off).tolist()


def compact_borders(prefix: bytes, skipped: Sequence[bytes] = ()) -> List[bytes]:
    sl = Slab.from_list([prefix, *skipped])
    data, pp, pop = _slab_ptrs(sl)
    out = np.zeros(2 * (int(sl.off[-1]) + 16 * sl.n) + 16, np.uint8)
    out_off = np.zeros(2 * sl.n + 1, np.uint64)
    lib().ko_compact_borders(pp, pop, sl.n, _p8(out), _p64(out_off))
    return Slab(out, out_off).tolist()


# ---- ring -----------------------------------------------------------------------------------------

class Ring:
    def __init__(self, capacity: int):
        self.capacity = capacity
        self.h = lib().ko_ring_new(capacity)

    def __del__(self):
        try:
            lib().ko_ring_free(self.h)
        except Exception:
            pass

    def add(self, rev: int, payload: int = 0):
        lib().ko_ring_add(self.h, rev, payload)

    def reset(self):
        lib().ko_ring_reset(self.h)

    def find(self, rev: int):
        ret = KoFindRet()
        revs = np.zeros(self.capacity, np.uint64)
        pay = np.zeros(self.capacity, np.uint64)
        lib().ko_ring_find(self.h, rev, C.byref(ret), _p64(revs), _p64(pay))
        n = int(ret.n_events)
        return ret, revs[:n].copy(), pay[:n].copy()


# ---- watch ------------------------------------------------------------------------------------------

def _events_c(ev: PackedEvents):
    kd, kp, kop = _slab_ptrs(ev.keys)
    c = KoEvents(kp, kop, _p64(ev.rev), ev.n, _p64(ev.batch_off), len(ev.batch_off) - 1)
    return c, kd


def fanout(ev: PackedEvents, w: PackedWatchers, threads: int = 1, alloc_per_batch: bool = False):
    """returns (start[W+1], event_idx[D], n_messages)"""
    evc, keep1 = _events_c(ev)
    pd, pp, pop = _slab_ptrs(w.prefixes)
    wc = KoWatchers(pp, pop, _p64(w.min_rev), w.n)
    out = KoFanout()
    lib().ko_fanout_run(C.byref(evc), C.byref(wc), threads, int(alloc_per_batch), C.byref(out))
    start = np.ctypeslib.as_array(out.start, shape=(w.n + 1,)).copy()
    d = int(out.n_deliveries)
    idx = np.ctypeslib.as_array(out.event_idx, shape=(max(d, 1),))[:d].copy()
    msgs = int(out.n_messages)
    lib().ko_fanout_free(C.byref(out))
    return start, idx, msgs


def watch_register(ring: Ring, ev: PackedEvents, prefix: bytes, revision: int, current_rev: int):
    evc, keep = _events_c(ev)
    reg = KoWatchReg()
    cap = max(ring.capacity, 1)
    cu = np.zeros(cap, np.uint64)
    lib().ko_watch_register(ring.h, C.byref(evc), prefix, len(prefix), revision, current_rev, C.byref(reg),
                            _p64(cu), cap)
    return reg.mode, int(reg.live_rev), cu[: int(reg.n_catchup)].copy(), int(reg.err_rev)


def catchup_chunks(n: int) -> List[int]:
    sizes = np.zeros(256, np.uint64)
    k = lib().ko_catchup_chunks(n, _p64(sizes), 256)
    return [int(x) for x in sizes[: int(k)]]


def bench_scan(st: OracleStore, start: bytes, end: bytes, read_rev: int, limit: int, faithful: bool, threads: int):
    ex, cs = C.c_uint64(), C.c_uint64()
    n = lib().ko_bench_scan(C.byref(st.c), start, len(start), end, len(end), read_rev, limit, int(faithful),
                            threads, C.byref(ex), C.byref(cs))
    return int(n), int(ex.value), int(cs.value)


# ---- etcd wire encoding (kb_oracle.h "etcd wire encoding") ---------------------------------------------
WIRE_KVS, WIRE_EVENTS = 1, 2


def _wire_sigs():
    L = lib()
    if getattr(L, "_wire_ready", False):
        return L
    L.ko_wire_elem_size.restype = C.c_uint64
    L.ko_wire_elem_size.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int]
    L.ko_wire_encode.restype = C.c_uint64
    L.ko_wire_encode.argtypes = [C.POINTER(KoStore), u64p, C.c_uint64, C.c_int, u8p, u64p]
    L.ko_wire_range_head.restype = C.c_uint64
    L.ko_wire_range_head.argtypes = [C.c_uint64, u8p]
    L.ko_wire_range_tail.restype = C.c_uint64
    L.ko_wire_range_tail.argtypes = [C.c_int, C.c_int64, u8p]
    L.ko_wire_watch_head.restype = C.c_uint64
    L.ko_wire_watch_head.argtypes = [C.c_uint64, C.c_int, C.c_char_p, C.c_uint64, u8p]
    L._wire_ready = True
    return L


def wire_elem_size(uk_len: int, val_len: int, rev: int, mode: int) -> int:
    return int(_wire_sigs().ko_wire_elem_size(uk_len, val_len, rev, mode))


def wire_encode(st: OracleStore, rec: Sequence[int], mode: int) -> Tuple[bytes, np.ndarray]:
    """the emitted records `rec` as consecutive repeated-field elements + n+1 element offsets"""
    L = _wire_sigs()
    r = np.ascontiguousarray(np.asarray(rec, dtype=np.uint64))
    off = np.zeros(len(r) + 1, np.uint64)
    total = int(L.ko_wire_encode(C.byref(st.c), _p64(r), len(r), mode, None, _p64(off)))
    out = np.zeros(max(total, 1), np.uint8)
    L.ko_wire_encode(C.byref(st.c), _p64(r), len(r), mode, _p8(out), _p64(off))
    return out[:total].tobytes(), off


def _wire_small(fn, *a) -> bytes:
    buf = np.zeros(int(fn(*a, None)) + 1, np.uint8)
    n = int(fn(*a, _p8(buf)))
    return buf[:n].tobytes()


def wire_range_head(header_rev: int) -> bytes:
    return _wire_small(_wire_sigs().ko_wire_range_head, header_rev)


def wire_range_tail(more: bool, count: int) -> bytes:
    return _wire_small(_wire_sigs().ko_wire_range_tail, int(more), count)


def wire_watch_head(header_rev: int, canceled: bool = False, reason: bytes = b"") -> bytes:
    return _wire_small(_wire_sigs().ko_wire_watch_head, header_rev, int(canceled), reason, len(reason))
