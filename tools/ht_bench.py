#!/usr/bin/env python
"""HT block decode on the device: resident throughput of ht_decode_kernel (+ the inverse plan behind it) and the end-to-end
decode through j2k_inverse_ht against j2k_inverse_batch on the same frames (compressed bytes instead of int32 planes over
PCIe).  One JSON line.  Streams come from the oracle-side generator (test infrastructure; untimed)."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--bits", type=int, default=12)
    ap.add_argument("--levels", type=int, default=5)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    import torch

    import ht_oracle_lib
    import ht_parity as HP
    import j2kb200
    import oracle_lib
    from j2kb200 import abi
    ht, orc = ht_oracle_lib.HtOracle(), oracle_lib.Oracle()
    ctx = j2kb200.Context()
    W = H = a.size
    rng = np.random.default_rng(0)
    # one frame of wavelet-like coefficients: forward 5/3 of a smooth + noisy 12-bit image
    yy, xx = np.mgrid[0:H, 0:W]
    img = ((np.sin(xx / 37.0) + np.cos(yy / 23.0)) * (1 << (a.bits - 3)) + (1 << (a.bits - 1)) + rng.normal(0, 6, (H, W))).clip(0, (1 << a.bits) - 1)
    raw = img.astype("<u2").view(np.uint8).reshape(-1)
    fp = abi.fwd_params(W, H, 1, 16, False, num_levels=a.levels, reversible=True, htj2k=True)
    ip = abi.inv_params(W, H, 1, 16, False, num_levels=a.levels, reversible=True, htj2k=True)
    co = ctx.forward(fp, raw).reshape(H, W)
    t0 = time.time()
    st, off, ln, km, mm, lay = HP.generated_stream(ht, orc, co, a.levels, 64, 64, rng, slack=0)
    gen_s = time.time() - t0
    nb = len(lay)
    F = a.frames
    stream = np.concatenate([st[:-16]] * F + [np.zeros(16, np.uint8)])
    per = st.size - 16
    rec = j2kb200.Context.ht_records(np.concatenate([off + f * per for f in range(F)]), np.tile(ln, F), np.tile(km, F), np.tile(mm, F))
    px, status = ctx.inverse_ht(ip, F, stream, rec)
    assert not status.any() and all(np.array_equal(px[f], raw) for f in range(F)), "HT decode + inverse != the image"
    # resident kernel time
    dev = torch.device("cuda:0")
    d_bytes = torch.from_numpy(stream).to(dev)
    d_rec = torch.from_numpy(rec.view(np.uint8)).to(dev)
    d_co = torch.empty(F * H * W, dtype=torch.int32, device=dev)
    d_px = torch.empty(F * H * W * 2, dtype=torch.uint8, device=dev)
    ts = torch.cuda.Stream()   # the library's launches and the timing events share this stream
    s = ts.cuda_stream
    def step():
        ctx.ht_decode_device(ip, F, C.c_void_p(d_bytes.data_ptr()), C.c_void_p(d_rec.data_ptr()), C.c_void_p(d_co.data_ptr()), True, stream=C.c_void_p(s))
    def step_full():
        step()
        ctx.inverse_device(ip, F, d_co.data_ptr(), d_px.data_ptr(), H * W * 2, stream=s)
    res = {}
    for name, fn in (("ht_decode_ms", step), ("ht_decode_plus_inverse_ms", step_full)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ts)
        for _ in range(a.steps):
            fn()
        e1.record(ts)
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / a.steps
    assert np.array_equal(d_px.cpu().numpy().reshape(F, -1)[F - 1], raw)
    # end to end, host buffers
    def timed(fn, n=5):
        fn()
        t = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t) / n * 1e3
    coeffs = np.tile(co.reshape(1, -1), (F, 1))
    e2e_ht = timed(lambda: ctx.inverse_ht(ip, F, stream, rec))
    e2e_planes = timed(lambda: ctx.inverse_batch(ip, coeffs))
    mpix = F * H * W / 1e6
    out = {"what": "HT cleanup decode on the device", "frame": [W, H], "frames": F, "bits": a.bits, "blocks_per_frame": nb,
           "compressed_bytes_per_frame": int(per), "bits_per_sample": per * 8 / (H * W), "generator_s_per_frame": gen_s,
           **res, "ht_decode_Mpixel_s": mpix / res["ht_decode_ms"] * 1e3,
           "e2e_inverse_ht_ms": e2e_ht, "e2e_inverse_ht_Mpixel_s": mpix / e2e_ht * 1e3,
           "e2e_inverse_batch_ms": e2e_planes, "e2e_inverse_batch_Mpixel_s": mpix / e2e_planes * 1e3,
           "h2d_bytes_ht": int(stream.size + rec.nbytes), "h2d_bytes_planes": int(coeffs.nbytes)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
