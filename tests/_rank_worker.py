"""worker of tests/test_multirank_cpu.py: one rank of a world_size-2 gloo group (CPU)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from kubebrain_b200 import synth  # noqa: E402
from kubebrain_b200.coder import NormalCoder  # noqa: E402
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
        print("OK", readable, int(counts[0]))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
