#!/usr/bin/env python
"""Host <-> device copy bandwidth per rank, alone and with every rank copying at once (what bounds bench.py's e2e leg).

    python tools/pcie_probe.py                       # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py
"""
import json
import os

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
nb = 1 << 30
h_in = torch.empty(nb, dtype=torch.uint8).pin_memory(); h_out = torch.empty(2 * nb, dtype=torch.uint8).pin_memory()
d_in = torch.empty(nb, dtype=torch.uint8, device="cuda"); d_out = torch.empty(2 * nb, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    s1.synchronize(); s2.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps / 1e3


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d(); d2h()


res = {}
for name, fn, nbytes in (("h2d", h2d, nb), ("d2h", d2h, 2 * nb), ("both", both, 3 * nb)):
    # event pair on the default stream does not see s1 / s2: use host wall time around stream syncs instead
    import time
    fn(); s1.synchronize(); s2.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        fn()
    s1.synchronize(); s2.synchronize()
    dt = (time.perf_counter() - t0) / 5
    res[name] = round(nbytes / dt / 1e9, 1)
out = [None] * world
if world > 1:
    dist.all_gather_object(out, res)
else:
    out = [res]
if rank == 0:
    print(json.dumps({"world": world, "GBps_per_rank": out, "note": "every rank copies at the same time; both = 1 GiB up + 2 GiB down concurrently, total bytes / time"}))
if world > 1:
    dist.destroy_process_group()
