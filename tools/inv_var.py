#!/usr/bin/env python
"""Run-to-run spread of the inverse ring kernel on C2 (32 frames): repeated timing rounds inside one process, for
several placements of the coefficient / pixel buffers (byte offsets into one big allocation) and one or two streams.
    python tools/inv_var.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch  # noqa: E402

import j2kb200  # noqa: E402
from j2kb200 import abi  # noqa: E402

W = H = 4096; L = 6; bits = 12; B = 32
enc, _ = j2kb200.openjpeg_quant_params(L, bits)
es, ds = j2kb200.runtime_quant_steps(enc, L, bits), j2kb200.decode_quant_steps(enc, L, bits, False)
fp = abi.fwd_params(W, H, 1, bits, False, 0, 0, L, False, False, abi.MCT_NONE, es)
ip = abi.inv_params(W, H, 1, bits, False, 0, 0, L, False, False, abi.MCT_NONE, ds)
fb = W * H * 2
ctx = j2kb200.Context(devices=[0])
g = torch.Generator(device="cuda").manual_seed(7)
d_in = torch.randint(0, 256, (B, fb), dtype=torch.uint8, device="cuda", generator=g)
d_in.view(B, -1, 2)[:, :, 1] &= 15
big_co = torch.empty(B * W * H * 4 + (64 << 20), dtype=torch.uint8, device="cuda")
big_px = torch.empty(2 * (B * fb + (64 << 20)), dtype=torch.uint8, device="cuda")
st = [torch.cuda.Stream() for _ in range(2)]


def rounds(co_off, px_off, nstreams, n_rounds=5, steps=20):
    co = big_co.data_ptr() + co_off
    px = [big_px.data_ptr() + px_off, big_px.data_ptr() + B * fb + (64 << 20) + px_off]
    ctx.forward_device(fp, B, d_in.data_ptr(), fb, co, stream=st[0].cuda_stream)
    torch.cuda.synchronize()
    out = []
    for _ in range(n_rounds):
        for i in range(3):
            ctx.inverse_device(ip, B, co, px[i % 2], fb, stream=st[i % nstreams].cuda_stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st[0]); st[1].wait_event(e0)
        for i in range(steps):
            ctx.inverse_device(ip, B, co, px[i % 2], fb, stream=st[i % nstreams].cuda_stream)
        ev = torch.cuda.Event(); ev.record(st[1]); st[0].wait_event(ev)
        e1.record(st[0]); torch.cuda.synchronize()
        out.append(round(e0.elapsed_time(e1) / steps, 4))
    return out


for ns in (2, 1):
    for co_off, px_off in ((0, 0), (256, 0), (4096, 0), (1 << 20, 0), (0, 4096), (0, 1 << 20), (3 << 20, 5 << 20), (0, 0)):
        print(json.dumps({"streams": ns, "co_off": co_off, "px_off": px_off, "ms_per_step": rounds(co_off, px_off, ns)}), flush=True)
ctx.close()
