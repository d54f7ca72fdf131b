"""worker of tests/test_multirank_cpu.py: one rank of a world_size-2 gloo group (CPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from kubebrain_b200 import sharded, synth  # noqa: E402
from kubebrain_b200.coder import NormalCoder, prefix_end  # noqa: E402
from kubebrain_b200.packed import PackedEvents, PackedWatchers, Slab  # noqa: E402
from oracle import binding as ko  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    coder = NormalCoder()
    # every rank owns the namespaces fnv1a64("ns-%05d") mod world maps to it
    store, meta = synth.gen_store(3000, 4, 64, 32, 40, shard=(rank, world))
    st = ko.OracleStore(store)
    lo, hi = coder.encode_object_key(b"/registry/", 0), coder.encode_object_key(b"/registry0", 0)
    local_cursor = meta.last_rev - 17 * rank  # pretend rank r has committed a little less
    # the one collective of the path: all-gather of the per-shard revision cursor, min = readable revision
    mine = torch.tensor([local_cursor], dtype=torch.int64)
    gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, mine)
    readable = min(int(t[0]) for t in gathered)
    r = ko.range_(st, lo, hi, readable)
    counts = torch.tensor([len(r.emit), store.n], dtype=torch.int64)
    dist.all_reduce(counts)
    if rank == 0:
        gstore, gmeta = synth.gen_store(3000, 4, 64, 32, 40)
        gst = ko.OracleStore(gstore)
        exp = ko.range_(gst, lo, hi, readable)
        assert readable == gmeta.last_rev - 17 * (world - 1)
        assert int(counts[1]) == gstore.n, "shards do not partition the store"
        assert int(counts[0]) == len(exp.emit), "sharded scan differs from the global scan"
    # ---- routing: every key of this shard hashes to this rank
    uks = [ko.decode(k)[0] for k in store.keys.tolist()]
    assert all(sharded.shard_of_key(uk, world) == rank for uk in uks if uk is not None)

    # ---- a broad List (prefix above namespace level) goes to every shard; rank 0 merges the runs and applies limit
    def kvs_of(res, stx):  # (key, value length, revision): the synthetic value BYTES depend on the record's position
        return [(k, len(v), rev) for k, v, rev in res.kvs(stx)]

    broad = b"/registry/pods/"
    assert sharded.owner_of_prefix(broad, world) is None
    blo, bhi = coder.encode_object_key(broad, 0), coder.encode_object_key(prefix_end(broad), 0)
    mine_run = kvs_of(ko.range_(st, blo, bhi, readable), store)
    runs = [None] * world
    dist.all_gather_object(runs, mine_run)
    # ---- a namespace List touches exactly one shard
    ns_prefix = b"/registry/pods/ns-00003/"
    owner = sharded.owner_of_prefix(ns_prefix, world)
    nlo, nhi = coder.encode_object_key(ns_prefix, 0), coder.encode_object_key(prefix_end(ns_prefix), 0)
    ns_run = kvs_of(ko.range_(st, nlo, nhi, readable), store)
    ns_runs = [None] * world
    dist.all_gather_object(ns_runs, ns_run)
    # ---- watch: events routed by the same hash; a cluster-wide watcher is registered on every shard
    gstore, gmeta = synth.gen_store(3000, 4, 64, 32, 40)
    evs = sorted((ko.decode(k)[1], ko.decode(k)[0]) for k in gstore.keys.tolist() if ko.decode(k)[1] > 0)
    wprefixes = [b"/registry/", broad, ns_prefix]
    watchers = PackedWatchers(Slab.from_list(wprefixes), np.array([0, evs[len(evs) // 3][0], 0], dtype=np.uint64))

    def deliveries(ev_list):
        n = len(ev_list)
        ev = PackedEvents(Slab.from_list([k for _, k in ev_list]), np.array([r for r, _ in ev_list], dtype=np.uint64),
                          np.array(list(range(0, n, 300)) + [n], dtype=np.uint64))
        start, idx, _ = ko.fanout(ev, watchers)
        return [[ev_list[int(i)] for i in idx[int(start[w]) : int(start[w + 1])]] for w in range(len(wprefixes))]

    my_events = [e for e in evs if sharded.shard_of_key(e[1], world) == rank]
    my_del = deliveries(my_events)
    all_del = [None] * world
    dist.all_gather_object(all_del, my_del)
    if rank == 0:
        gst2 = ko.OracleStore(gstore)
        for limit in (0, 1, 7, 10_000):
            exp = ko.range_(gst2, blo, bhi, readable, limit + 1 if limit else 0)  # backend.List asks for limit+1
            exp_kvs = kvs_of(exp, gstore)
            got, more = sharded.merge_list_runs(runs, limit)
            assert got == (exp_kvs[:limit] if limit else exp_kvs), ("merged List", limit)
            assert more == (limit > 0 and len(exp_kvs) > limit)
        exp_ns = kvs_of(ko.range_(gst2, nlo, nhi, readable), gstore)
        assert ns_runs[owner] == exp_ns and exp_ns, "namespace List answered by its owner"
        assert all(not r for i, r in enumerate(ns_runs) if i != owner), "other shards hold nothing of that namespace"
        glob = deliveries(evs)
        for w in range(len(wprefixes)):
            merged = sharded.merge_watch_streams([all_del[r][w] for r in range(world)])
            assert merged == glob[w], ("watch stream", w)
        assert sharded.owner_of_prefix(ns_prefix, world) == sharded.shard_of_key(ns_prefix + b"x", world)
        print("OK", readable, int(counts[0]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
