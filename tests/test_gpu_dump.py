"""Durable dump / restore of the HBM snapshot: a restored engine answers every scan / get / sweep
like the engine that wrote the dump (and like the oracle); corrupt, truncated and foreign files are refused."""
from __future__ import annotations

import os
import struct

import numpy as np
import pytest

from kubebrain_b200 import synth
from kubebrain_b200._lib import KB_OUT_HOST, Engine, KbError
from kubebrain_b200.coder import NormalCoder
from kubebrain_b200.packed import PackedStore
from oracle import binding as ko
from tests import fuzz
from tests.test_gpu_parity import check_compact, check_gets, check_ranges

pytestmark = pytest.mark.gpu
CODER = NormalCoder()


def test_dump_restore_roundtrip(tmp_path):
    a, b = Engine(0), Engine(0)
    try:
        for seed in range(4):
            store = fuzz.fuzz_store(900 + seed, n_keys=50 + 40 * seed)
            st = ko.OracleStore(store)
            a.load_sorted(store)
            a.set_compact_revision(7 if seed & 1 else None)
            path = str(tmp_path / f"snap{seed}.kbd")
            a.dump(path)
            assert not os.path.exists(path + ".tmp")
            b.restore(path)
            assert b.store_info() == a.store_info()
            reqs = []
            for s, e in fuzz.fuzz_bounds(store, seed):
                for rev in (7, 23, 60, 2**64 - 1):
                    for lim in (0, 3):
                        reqs.append((s, e, rev, lim))
            check_ranges(b, store, st, reqs)
            if seed & 1:  # the compact-revision record travels with the dump (checkCompactRace, scanner.go:594-626)
                with pytest.raises(KbError) as ei:
                    b.range_batch([(b"\x00", b"\xff", 3, 0)], KB_OUT_HOST)
                assert ei.value.code == -5
            b.set_compact_revision(None)
            check_compact(b, store, st, b"\x00", b"\xff" * 4, 35)
    finally:
        a.close()
        b.close()


def test_dump_after_apply_batch_and_config_shape(tmp_path):
    """a snapshot maintained by kb_apply_batch dumps what a rebuild would hold; config-2 record shape (2 KB values)"""
    store, meta = synth.gen_store(3000, 4, 256, 2048, 30, config_id=2)
    items = dict(zip(store.keys.tolist(), store.vals.tolist()))
    a, b = Engine(0), Engine(0)
    try:
        a.load_sorted(store)
        keys = sorted(items)
        ops = [(keys[i], None) for i in range(0, len(keys), 7)] + [(keys[i], b"new" * (i % 50)) for i in range(3, len(keys), 11)]
        ops += [(keys[5][:-8] + struct.pack(">Q", 2**40 + i), b"ins%d" % i) for i in range(50)]
        a.apply_batch(ops)
        for k, v in ops:
            if v is None:
                items.pop(k, None)
            else:
                items[k] = v
        cur = PackedStore.from_items(list(items.items()))
        st = ko.OracleStore(cur)
        path = str(tmp_path / "snap.kbd")
        a.dump(path)
        b.restore(path)
        assert b.store_info()[0] == cur.n
        lo, hi = CODER.encode_object_key(b"/registry/", 0), CODER.encode_object_key(b"/registry0", 0)
        check_ranges(b, cur, st, [(lo, hi, meta.read_rev, 0), (lo, hi, 2**64 - 1, 501), (lo, hi, meta.first_rev, 0)])
        uks = [ko.decode(k)[0] for k in cur.keys.tolist()[::97]]
        check_gets(b, cur, st, [(uk, 0) for uk in uks if uk is not None])
        # a restored snapshot keeps accepting writes
        b.apply_batch([(keys[1], b"again")])
    finally:
        a.close()
        b.close()


def test_restore_refuses_bad_files(tmp_path):
    e = Engine(0)
    try:
        store = fuzz.fuzz_store(77, n_keys=80)
        e.load_sorted(store)
        good = str(tmp_path / "good.kbd")
        e.dump(good)
        blob = open(good, "rb").read()
        cases = {
            "foreign": b"not a dump at all" * 10,
            "truncated": blob[: len(blob) - 100],
            "flipped_slab_byte": blob[:-40] + bytes([blob[-40] ^ 0x55]) + blob[-39:],
            "flipped_dir_byte": blob[:120] + bytes([blob[120] ^ 0x01]) + blob[121:],
            "empty": b"",
        }
        for name, data in cases.items():
            p = str(tmp_path / (name + ".kbd"))
            with open(p, "wb") as f:
                f.write(data)
            with pytest.raises(KbError):
                e.restore(p)
            # a failed restore never leaves a half-loaded snapshot: either the old one is intact or none is loaded
            try:
                assert e.store_info()[0] == store.n
            except KbError as err:
                assert err.code == -6
            e.restore(good)
            assert e.store_info()[0] == store.n
        with pytest.raises(KbError) as ei:
            e.restore(str(tmp_path / "missing.kbd"))
        assert ei.value.code == -9
        with pytest.raises(KbError) as ei:
            e.dump(str(tmp_path / "no_such_dir" / "x.kbd"))
        assert ei.value.code == -9
    finally:
        e.close()
