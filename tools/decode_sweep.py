#!/usr/bin/env python
"""Times k_decode_lcp (and k_emit_place) alone for several launch geometries.  KB_DECODE_K / KB_DECODE_NKS /
KB_DECODE_WARPS are read once per process, so every geometry runs in its own subprocess.

    python tools/decode_sweep.py            # parent: runs the sweep, prints one line per geometry
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(which: str):
    from kubebrain_b200 import synth
    from kubebrain_b200._lib import KB_OUT_COUNT, KB_OUT_DEVICE, Engine
    from kubebrain_b200.coder import NormalCoder

    c = NormalCoder()
    lo, hi = c.encode_object_key(b"/registry/", 0), c.encode_object_key(b"/registry0", 0)
    e = Engine(0)
    if which == "long":
        store, meta = synth.gen_store(200_000, 4, 256, 2048, 1000, config_id=2)
    else:
        store, meta = synth.gen_store(1_000_000, 9, 64, 64, 10000, config_id=4, tomb_frac=0.02)
    e.load_sorted(store)
    out = {}
    if which == "long":
        reqs = Engine.pack_range_reqs([(lo, hi, meta.read_rev, 0)])
        for _ in range(5):
            e.range_batch(reqs, KB_OUT_COUNT).close()
        e.prof_reset()
        e.prof_enable(1)
        for _ in range(20):
            e.range_batch(reqs, KB_OUT_COUNT).close()
        e.prof_enable(0)
    else:
        for _ in range(3):
            e.compact_sweep(lo, hi, meta.last_rev, out_mode=KB_OUT_DEVICE).close()
        e.prof_reset()
        e.prof_enable(1)
        for _ in range(8):
            e.compact_sweep(lo, hi, meta.last_rev, out_mode=KB_OUT_DEVICE).close()
        e.prof_enable(0)
    for p in e.prof_read():
        if p["launches"] and p["name"].startswith("k_"):
            out[p["name"]] = {"us": round(1e3 * p["total_ms"] / p["launches"], 1),
                              "gbs": round(p["alg_bytes"] / p["launches"] / 1e9 / (p["total_ms"] / p["launches"] / 1e3), 0)}
    print("RESULT " + json.dumps(out), flush=True)
    e.close()


def main():
    if len(sys.argv) > 1:
        child(sys.argv[1])
        return
    geoms = {"long": [(1, 2, 0)],
             "short": [(2, 2, 0), (1, 2, 0), (1, 2, 16), (2, 2, 12), (3, 2, 0)]}
    if os.environ.get("KB_SWEEP_FULL"):
        geoms = {"long": [(1, 2, 0), (1, 3, 0), (1, 4, 0), (1, 3, 6), (1, 2, 8)],
                 "short": [(2, 2, 0), (2, 3, 0), (2, 4, 0), (3, 3, 0), (4, 3, 0), (4, 2, 0), (1, 3, 0), (1, 4, 0)]}
    for which, gl in geoms.items():
        for k, nks, w in gl:
            env = dict(os.environ, KB_DECODE_K=str(k), KB_DECODE_NKS=str(nks), KB_DECODE_WARPS=str(w))
            r = subprocess.run([sys.executable, os.path.abspath(__file__), which], env=env, capture_output=True, text=True,
                               timeout=600)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            print(which, f"K={k} NKS={nks} warps={w or 'auto'}", line[0][7:] if line else "FAILED " + r.stderr[-400:], flush=True)


if __name__ == "__main__":
    main()
