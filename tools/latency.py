#!/usr/bin/env python
"""Latency of the blocking single-frame calls (what a frame-by-frame codec adapter sees): j2k_forward / j2k_inverse on one
C1 frame (512x512 16-bit, 5/3, 5 levels) and one C2 frame (4096x4096 12-bit, 9/7, 6 levels), host buffers, wall clock."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import j2kb200
from j2kb200 import abi

def med(fn, n=200):
    for _ in range(10):
        fn()
    t = []
    for _ in range(n):
        t0 = time.perf_counter(); fn(); t.append(time.perf_counter() - t0)
    t.sort()
    return t[len(t) // 2] * 1e6, t[int(len(t) * 0.99)] * 1e6

with j2kb200.Context(devices=[0]) as ctx:
    rng = np.random.default_rng(1)
    out = {}
    for name, w, h, bits, L, rev in (("C1", 512, 512, 16, 5, True), ("C2", 4096, 4096, 12, 6, False)):
        es = ds = None
        if not rev:
            enc, _ = j2kb200.openjpeg_quant_params(L, bits)
            es, ds = j2kb200.runtime_quant_steps(enc, L, bits), j2kb200.decode_quant_steps(enc, L, bits, False)
        fp = abi.fwd_params(w, h, 1, bits, False, num_levels=L, reversible=rev, steps=es)
        ip = abi.inv_params(w, h, 1, bits, False, num_levels=L, reversible=rev, steps=ds)
        px = ctx.pinned(w * h * 2)
        px[:] = rng.integers(0, 1 << (bits - 8), w * h * 2, dtype=np.uint8)
        co = ctx.pinned(w * h * 4, np.int32)
        f = med(lambda: ctx.lib.j2k_forward(ctx.h, fp, px.ctypes.data, px.size, co.ctypes.data, co.size), 200 if w < 1000 else 40)
        i = med(lambda: ctx.lib.j2k_inverse(ctx.h, ip, co.ctypes.data, co.size, px.ctypes.data, px.size, None), 200 if w < 1000 else 40)
        out[name] = {"forward_us_median_p99": [round(x, 1) for x in f], "inverse_us_median_p99": [round(x, 1) for x in i],
                     "forward_Mpixel_s": round(w * h / f[0], 1), "inverse_Mpixel_s": round(w * h / i[0], 1)}
    print(json.dumps(out))
