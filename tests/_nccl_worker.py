"""worker of test_two_gpu_cursor_allgather: one rank of a 2-GPU NCCL group (run under torch.distributed.run)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kubebrain_b200._lib import Engine  # noqa: E402

r = int(os.environ["RANK"])
torch.cuda.set_device(r)
dist.init_process_group("nccl", device_id=torch.device("cuda", r))
uid = Engine.nccl_unique_id() if r == 0 else bytes(128)
t = torch.tensor(list(uid), dtype=torch.uint8, device="cuda")
dist.broadcast(t, 0)
e = Engine(r)
e.nccl_init(bytes(t.cpu().tolist()), r, 2)
import time  # noqa: E402

e.prof_enable(1)
for k in range(200):  # many epochs: the peer-memory exchange alternates between its two slot sets
    a, m = e.cursor_allgather(1000 + 7 * r + k)
    assert a.tolist() == [1000 + k, 1007 + k] and m == 1000 + k, (a, m)
    if k % 50 == r:  # skew the ranks against each other
        time.sleep(0.002)
e.prof_enable(0)
names = {p["name"] for p in e.prof_read() if p["launches"]}
dist.barrier()
t0 = time.perf_counter()
for k in range(200):
    e.cursor_allgather(5000 + k)
dt = (time.perf_counter() - t0) / 200
dist.barrier()
if r == 0:
    print("OK mode=%s %.1f us per exchange" % ("p2p" if "k_cursor_p2p" in names else "nccl", dt * 1e6))
e.close()
dist.destroy_process_group()
