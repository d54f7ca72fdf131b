"""Generates tests/golden/wire_golden.json: etcd v3.5.2 messages serialized by the protobuf runtime from the restated
schema (etcd_schema.py).  Run from the repo root:  python tests/golden/make_wire_golden.py"""
import json
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.golden import etcd_schema as es  # noqa: E402


def cases():
    rng = random.Random(20260922)
    out = []
    edge = [(b"", b"", 0), (b"/a", b"", 1), (b"", b"v", 127), (b"/registry/pods/ns/x", b"\x00" * 127, 128),
            (b"k" * 128, b"v" * 16383, 16383), (b"k" * 300, b"v" * 16384, 16384), (b"\xff\x00$", b"tombstone", 2**32),
            (b"/z", b"z", 2**63 - 1), (b"/z", b"z", 2**63), (b"/z", b"z", 2**64 - 1)]
    out.append(dict(name="edge", header_rev=1700000010, more=True, kvs=edge))
    out.append(dict(name="empty", header_rev=0, more=False, kvs=[]))
    for i in range(6):
        kvs, seen = [], set()
        while len(kvs) < [1, 2, 7, 40, 301, 299][i]:
            kv = (bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 20, 127, 128, 269]))),
                  bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 9, 127, 128, 200]))),
                  rng.choice([1, 5, 300, 2**21, 2**35, 2**56, 2**63 + 5]))
            if (kv[0], kv[2]) not in seen:  # (user key, revision) names one store record
                seen.add((kv[0], kv[2]))
                kvs.append(kv)
        out.append(dict(name=f"rand{i}", header_rev=rng.choice([0, 7, 2**40]), more=bool(i & 1), kvs=kvs))
    return out


def main():
    M = es.build()
    gold = []
    for c in cases():
        kvs = c["kvs"]
        n = len(kvs)
        g = dict(name=c["name"], header_rev=c["header_rev"], more=c["more"],
                 kvs=[[k.hex(), v.hex(), rev] for k, v, rev in kvs],
                 range_response=es.range_response(M, c["header_rev"], kvs, c["more"], n + (1 if c["more"] else 0)).hex(),
                 watch_batches=[es.watch_batch(M, 0, kvs[i:i + 300]).hex() for i in range(0, n, 300)],
                 watch_end=es.watch_cancel(M, c["header_rev"], "").hex(),
                 watch_end_err=es.watch_cancel(M, c["header_rev"], "context deadline exceeded").hex())
        gold.append(g)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "wire_golden.json")
    with open(path, "w") as f:
        json.dump(gold, f)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
