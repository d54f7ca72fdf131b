"""Timing probe for k_decode_lcp across call sequences (dev tool, GPU box only)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from kubebrain_b200 import synth
from kubebrain_b200._lib import Engine, KB_OUT_COUNT, KB_OUT_DEVICE
from kubebrain_b200.coder import NormalCoder

C = NormalCoder()
store, meta = synth.gen_store(200_000, 4, 256, 2048, 1000, config_id=2)
eng = Engine(0)
eng.load_sorted(store)
lo, hi = C.encode_object_key(b"/registry/", 0), C.encode_object_key(b"/registry0", 0)
reqs = Engine.pack_range_reqs([(lo, hi, meta.read_rev, 0)])


def prof(label, mode, n):
    for i in range(n):
        eng.prof_reset()
        eng.prof_enable(1)
        r = eng.range_batch(reqs, mode)
        r.close()
        eng.prof_enable(0)
        p = {e["name"]: e for e in eng.prof_read()}
        d = p.get("k_decode_lcp")
        g = p.get("k_gather")
        print(label, i, "decode %.1f us" % (d["total_ms"] * 1e3 / max(d["launches"], 1)),
              ("gather %.1f us" % (g["total_ms"] * 1e3 / max(g["launches"], 1))) if g and g["launches"] else "")


prof("count ", KB_OUT_COUNT, 5)
prof("device", KB_OUT_DEVICE, 4)
prof("count ", KB_OUT_COUNT, 4)
time.sleep(0.5)
prof("count-after-idle", KB_OUT_COUNT, 3)
