#!/usr/bin/env python
"""j2k_forward_ht / j2k_inverse_ht end to end (pinned host buffers) on one device against every visible device of ONE process:
the in-process multi-GPU form of the codec adapters (j2k_init over all devices).  python tools/ht_multi_probe.py [frames]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import torch
import j2kb200
from j2kb200 import abi
import ht_parity as HP

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W = H = 4096
L, bits = 6, 12
enc, _ = j2kb200.openjpeg_quant_params(L, bits)
es = j2kb200.runtime_quant_steps(enc, L, bits)
ds = j2kb200.decode_quant_steps(enc, L, bits, True)
fp = abi.fwd_params(W, H, 1, bits, False, num_levels=L, reversible=False, htj2k=True, steps=es)
ip = abi.inv_params(W, H, 1, bits, False, num_levels=L, reversible=False, htj2k=True, steps=ds)
kmax = HP.band_kmax_table(1, L, bits + 2)
yy, xx = np.mgrid[0:H, 0:W]
rng = np.random.default_rng(2)
base = ((np.sin(xx / 17.0) * np.cos(yy / 23.0) * 0.25 + 0.5) * 4095)
out = {"frames": B, "visible_devices": torch.cuda.device_count()}
for devs in ([0], list(range(torch.cuda.device_count()))):
    with j2kb200.Context(devices=devs) as ctx:
        h_in = ctx.pinned(B * W * H * 2).reshape(B, -1)
        for f in range(B):
            h_in[f] = (base + rng.normal(0, 16, (H, W))).clip(0, 4095).astype("<u2").reshape(-1).view(np.uint8)
        cap = int(B * W * H * 1.5)
        h_bytes = ctx.pinned(cap)
        stream, rec = ctx.forward_ht(fp, h_in, kmax, 64, 64, out=h_bytes)
        t0 = time.perf_counter(); K = 3
        for _ in range(K):
            stream, rec = ctx.forward_ht(fp, h_in, kmax, 64, 64, out=h_bytes)
        dt = (time.perf_counter() - t0) / K
        h_px = ctx.pinned(B * W * H * 2).reshape(B, -1)
        h_st = ctx.pinned(rec.size * 4, np.int32)
        ctx.wait(ctx.submit_inverse_ht(ip, B, h_bytes[:stream.size + 16], rec, h_px, h_st))
        t0 = time.perf_counter()
        for _ in range(K):
            ctx.wait(ctx.submit_inverse_ht(ip, B, h_bytes[:stream.size + 16], rec, h_px, h_st))
        dti = (time.perf_counter() - t0) / K
        st = h_st
        out[f"devices_{len(devs)}"] = {"encode_Mpixel_s": B * W * H / dt / 1e6, "decode_Mpixel_s": B * W * H / dti / 1e6,
                                       "stream_bytes": int(stream.size), "bits_per_sample": 8.0 * stream.size / (B * W * H),
                                       "decode_status_clean": bool(not st.any())}
    if len(devs) == torch.cuda.device_count() == 1:
        break
print(json.dumps(out))
