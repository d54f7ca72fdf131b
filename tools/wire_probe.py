"""Timing probe for the wire-mode copy kernel (dev tool, GPU box only)."""
import sys
sys.path.insert(0, ".")
from kubebrain_b200 import synth
from kubebrain_b200._lib import Engine, KB_OUT_DEVICE, KB_WIRE_ETCD_EVENTS, KB_WIRE_ETCD_KVS
from kubebrain_b200.coder import NormalCoder

C = NormalCoder()
store, meta = synth.gen_store(200_000, 4, 256, 2048, 1000, config_id=2)
eng = Engine(0)
eng.load_sorted(store)
lo, hi = C.encode_object_key(b"/registry/", 0), C.encode_object_key(b"/registry0", 0)
reqs = Engine.pack_range_reqs([(lo, hi, meta.read_rev, 0)])
for mode in (KB_WIRE_ETCD_EVENTS, KB_WIRE_ETCD_KVS, 0):
    for i in range(3):
        eng.prof_reset()
        eng.prof_enable(1)
        r = eng.range_batch(reqs, KB_OUT_DEVICE | mode)
        nb = r.n_bytes
        r.close()
        eng.prof_enable(0)
        p = {e["name"]: e for e in eng.prof_read()}
        k = p.get("k_wire_copy") or p.get("k_gather")
        print(mode, i, nb, "%s %.1f us" % (k["name"], k["total_ms"] * 1e3))
