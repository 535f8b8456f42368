#!/usr/bin/env python
"""Resident throughput of the BASELINE.json configs other than the bench workload (parity cases, not bench lines):
prints one JSON line per config with forward / inverse Mpixel/s and the achieved fraction of the HBM roofline
(SURVEY 8d algorithmic bytes).  Run on a B200:  python tools/config_bench.py  [--steps 10]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import j2kb200  # noqa: E402
from j2kb200 import abi  # noqa: E402


def alg_bytes(S, s_io, L):
    return S * (s_io + 4) + 8 * S * sum(4.0 ** -k for k in range(1, L))


def run_config(ctx, name, w, h, c, bits, signed, L, rev, frames, tile=(0, 0), steps=10, peak=6547.2):
    mct = (abi.MCT_RCT if rev else abi.MCT_ICT) if c == 3 else abi.MCT_NONE
    es = ds = None
    if not rev:
        enc, _ = j2kb200.openjpeg_quant_params(L, bits)
        es, ds = j2kb200.runtime_quant_steps(enc, L, bits), j2kb200.decode_quant_steps(enc, L, bits, False)
    fp = abi.fwd_params(w, h, c, bits, signed, tile[0], tile[1], L, rev, False, mct, es)
    ip = abi.inv_params(w, h, c, bits, signed, tile[0], tile[1], L, rev, False, mct, ds)
    bps = 1 if bits <= 8 else 2
    fb = w * h * c * bps
    g = torch.Generator(device="cuda").manual_seed(7)
    d_in = torch.randint(0, 256, (frames, fb), dtype=torch.uint8, device="cuda", generator=g)
    if bits > 8 and bits < 16:  # keep the high byte inside the bit depth
        d_in.view(frames, -1, 2)[:, :, 1] &= (1 << (bits - 8)) - 1
    d_co = [torch.empty((frames, w * h * c), dtype=torch.int32, device="cuda") for _ in range(2)]
    d_px = [torch.empty((frames, fb), dtype=torch.uint8, device="cuda") for _ in range(2)]
    st = [torch.cuda.Stream() for _ in range(2)]
    torch.cuda.synchronize()

    def timed(fn):
        for i in range(6):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st[0]); st[1].wait_event(e0)
        for i in range(steps):
            fn(i)
        ev = torch.cuda.Event(); ev.record(st[1]); st[0].wait_event(ev)
        e1.record(st[0])
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    l0 = ctx.launch_count
    fms = timed(lambda i: ctx.forward_device(fp, frames, d_in.data_ptr(), fb, d_co[i % 2].data_ptr(), stream=st[i % 2].cuda_stream))
    launches_per_call = (ctx.launch_count - l0) / (6 + steps)
    ims = timed(lambda i: ctx.inverse_device(ip, frames, d_co[0].data_ptr(), d_px[i % 2].data_ptr(), fb, stream=st[i % 2].cuda_stream))
    lossless_ok = bool(torch.equal(d_px[0], d_in)) if rev else None
    S = frames * w * h * c
    ab = alg_bytes(S, bps, L)
    pix = frames * w * h
    print(json.dumps({"config": name, "frames": frames, "fwd_Mpixel_s": pix / fms / 1e3, "inv_Mpixel_s": pix / ims / 1e3,
                      "fwd_ms": fms, "inv_ms": ims, "fwd_frac_hbm": ab / (fms * 1e-3) / 1e9 / peak, "inv_frac_hbm": ab / (ims * 1e-3) / 1e9 / peak,
                      "launches_per_forward_call": launches_per_call, "lossless_roundtrip_identical": lossless_ok}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--only", default="", help="substring of the config name to run")
    a = ap.parse_args()
    peak = 6547.2
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    with j2kb200.Context(devices=[0]) as ctx:
        def run(ctx, name, *aa, **kw):
            if a.only in name:
                run_config(ctx, name, *aa, **kw)
        run(ctx, "C1x256: 512x512 16-bit signed mono, 5/3 L5", 512, 512, 1, 16, True, 5, True, 256, steps=a.steps, peak=peak)
        run(ctx, "C2x16: 4096x4096 12-bit mono, 9/7 L6", 4096, 4096, 1, 12, False, 6, False, 16, steps=a.steps, peak=peak)
        run(ctx, "C3i x8: 2048x2048 RGB 8-bit, ICT + 9/7 L5", 2048, 2048, 3, 8, False, 5, False, 8, steps=a.steps, peak=peak)
        run(ctx, "C3ii x8: 2048x2048 RGB 8-bit, RCT + 5/3 L5", 2048, 2048, 3, 8, False, 5, True, 8, steps=a.steps, peak=peak)
        run(ctx, "C4 block: 250 frames 512x512 16-bit, 5/3 L5", 512, 512, 1, 16, False, 5, True, 250, steps=a.steps, peak=peak)
        run(ctx, "C5 block: 128 tiles 1024x1024 RGB 8-bit (8192x16384 image), 9/7 L7", 8192, 16384, 3, 8, False, 7, False, 1, tile=(1024, 1024),
            steps=a.steps, peak=peak)


if __name__ == "__main__":
    main()
