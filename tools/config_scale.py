#!/usr/bin/env python
"""BASELINE configs C4 and C5 sharded over N GPUs (one rank per GPU, no collective on the data path):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/config_scale.py

C4: the 2000-frame 512x512 16-bit CT series, 5/3 lossless, contiguous blocks of ceil(2000/N) frames per rank.
C5: the 32768x32768 RGB slide as 1024 tiles of 1024x1024, 9/7 7 levels, contiguous blocks of 1024/N tiles per rank (a rank's
block is one image of 8 tile columns).  Device-resident forward and inverse passes, CUDA events, max over ranks;
rank 0 prints one JSON line per config with the aggregate Mpixel/s."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import j2kb200  # noqa: E402
from j2kb200 import abi, shard  # noqa: E402


def timed(fn, streams, steps, barrier):
    for i in range(4):
        fn(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(streams[0]); streams[1].wait_event(e0)
    for i in range(steps):
        fn(i)
    ev = torch.cuda.Event(); ev.record(streams[1]); streams[0].wait_event(ev)
    e1.record(streams[0])
    barrier()
    return e0.elapsed_time(e1) / steps


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_ranks(ms):
        t = torch.tensor([ms], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    steps = 10
    ctx = j2kb200.Context(devices=[local])
    streams = [torch.cuda.Stream() for _ in range(2)]
    out = []
    # ---- C4
    lo, hi = shard.unit_range(2000, rank, world)
    n = hi - lo
    w = h = 512
    fb = w * h * 2
    fp = abi.fwd_params(w, h, 1, 16, False, num_levels=5, reversible=True)
    ip = abi.inv_params(w, h, 1, 16, False, num_levels=5, reversible=True)
    g = torch.Generator(device="cuda").manual_seed(1000 + rank)
    d_in = torch.randint(0, 256, (n, fb), dtype=torch.uint8, device="cuda", generator=g)
    d_co = [torch.empty((n, w * h), dtype=torch.int32, device="cuda") for _ in range(2)]
    d_px = [torch.empty((n, fb), dtype=torch.uint8, device="cuda") for _ in range(2)]
    f_ms = max_ranks(timed(lambda i: ctx.forward_device(fp, n, d_in.data_ptr(), fb, d_co[i % 2].data_ptr(), stream=streams[i % 2].cuda_stream), streams, steps, barrier))
    i_ms = max_ranks(timed(lambda i: ctx.inverse_device(ip, n, d_co[0].data_ptr(), d_px[i % 2].data_ptr(), fb, stream=streams[i % 2].cuda_stream), streams, steps, barrier))
    ok = bool(torch.equal(d_px[0], d_in))
    out.append({"config": "C4: 2000 frames 512x512 16-bit, 5/3 lossless, frame-sharded", "n_gpus": world, "units_per_rank": n,
                "fwd_Mpixel_s": 2000 * w * h / f_ms / 1e3, "inv_Mpixel_s": 2000 * w * h / i_ms / 1e3, "fwd_ms": f_ms, "inv_ms": i_ms,
                "lossless_roundtrip_identical_rank0": ok})
    del d_in, d_co, d_px
    # ---- C5
    lo, hi = shard.unit_range(1024, rank, world)
    nt = hi - lo                      # tiles of this rank: an image of 8 tile columns x nt/8 tile rows
    W5, H5 = 8 * 1024, (nt // 8) * 1024
    enc, _ = j2kb200.openjpeg_quant_params(7, 8)
    es, ds = j2kb200.runtime_quant_steps(enc, 7, 8), j2kb200.decode_quant_steps(enc, 7, 8, False)
    fp = abi.fwd_params(W5, H5, 3, 8, False, 1024, 1024, 7, False, False, abi.MCT_ICT, es)
    ip = abi.inv_params(W5, H5, 3, 8, False, 1024, 1024, 7, False, False, abi.MCT_ICT, ds)
    fb = W5 * H5 * 3
    d_in = torch.randint(0, 256, (1, fb), dtype=torch.uint8, device="cuda", generator=g)
    d_co = [torch.empty((1, fb), dtype=torch.int32, device="cuda") for _ in range(2)]
    d_px = [torch.empty((1, fb), dtype=torch.uint8, device="cuda") for _ in range(2)]
    f_ms = max_ranks(timed(lambda i: ctx.forward_device(fp, 1, d_in.data_ptr(), fb, d_co[i % 2].data_ptr(), stream=streams[i % 2].cuda_stream), streams, steps, barrier))
    i_ms = max_ranks(timed(lambda i: ctx.inverse_device(ip, 1, d_co[0].data_ptr(), d_px[i % 2].data_ptr(), fb, stream=streams[i % 2].cuda_stream), streams, steps, barrier))
    pix = 32768.0 * 32768.0 * (nt * world / 1024.0)
    out.append({"config": "C5: 32768x32768 RGB 8-bit, 1024 tiles of 1024x1024, ICT + 9/7 L7, tile-sharded", "n_gpus": world, "units_per_rank": nt,
                "fwd_Mpixel_s": pix / f_ms / 1e3, "inv_Mpixel_s": pix / i_ms / 1e3, "fwd_ms": f_ms, "inv_ms": i_ms})
    if rank == 0:
        for o in out:
            print(json.dumps(o), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
