"""N > 1 host logic on CPU: two gloo ranks shard a frame series with no data-path collective.

Each rank takes its contiguous block of frames (j2kb200.shard.unit_range — the rule run_host_batch()
uses across devices), transforms it with the CPU oracle standing in for its GPU (this is a test of
the sharding / aggregation logic, not of the kernels), and the job-level result must equal the
single-rank result; the only collectives are control: max-over-ranks of the step time and a gather
of per-rank checksums."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200"), HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    import oracle_lib
    from j2kb200 import abi, shard
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = h = 64
        fp = abi.fwd_params(w, h, 1, 16, False, num_levels=3, reversible=True)
        orc = oracle_lib.Oracle()
        b, e = shard.unit_range(n_frames, rank, world)
        sums = []
        for f in range(b, e):  # frame f uses seed 1000 + f
            frame = np.random.default_rng(1000 + f).integers(0, 65536, h * w, dtype=np.uint16).view(np.uint8)
            sums.append(int(orc.forward(fp, frame).astype(np.int64).sum()))
        ms_local = 1.0 + rank  # pretend device times: the job time is the slowest rank's
        ms = shard.max_over_ranks(ms_local, dist)
        gathered = [None] * world
        dist.all_gather_object(gathered, (b, e, sums))
        if rank == 0:
            q.put((ms, gathered))
    finally:
        dist.destroy_process_group()


def test_unit_range_partitions_exactly():
    from j2kb200 import shard
    for n in (0, 1, 7, 8, 2000, 1024):
        for world in (1, 2, 3, 4, 8):
            rs = [shard.unit_range(n, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(rs, rs[1:]))
            assert sum(shard.shard_sizes(n, world)) == n
    assert shard.unit_range(2000, 7, 8) == (1750, 2000)  # C4: 250 frames per GPU
    assert shard.unit_range(1024, 3, 8) == (384, 512)    # C5: 128 tiles per GPU
    with pytest.raises(ValueError):
        shard.unit_range(8, 8, 8)
    assert shard.aggregate_throughput([250] * 8, 10.0) == 2000 / 0.01


def test_two_gloo_ranks_shard_frames_without_data_collective():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_frames, world, port = 7, 2, _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    ms, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ms == 2.0  # max over ranks
    assert [(g[0], g[1]) for g in gathered] == [(0, 4), (4, 7)]
    # the sharded job equals the single-rank job, frame by frame
    sys.path.insert(0, HERE)
    import oracle_lib
    from j2kb200 import abi
    orc = oracle_lib.Oracle()
    fp = abi.fwd_params(64, 64, 1, 16, False, num_levels=3, reversible=True)
    want = []
    for f in range(n_frames):
        frame = np.random.default_rng(1000 + f).integers(0, 65536, 64 * 64, dtype=np.uint16).view(np.uint8)
        want.append(int(orc.forward(fp, frame).astype(np.int64).sum()))
    assert [s for g in gathered for s in g[2]] == want
