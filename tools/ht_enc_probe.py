#!/usr/bin/env python
"""One j2k_forward_ht call on C2-shaped frames (for `ncu --metrics gpu__time_duration.sum` launch lists and quick timings)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200")):
    sys.path.insert(0, p)
import bench, j2kb200
from j2kb200 import abi
F = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = j2kb200.Context()
W, H, L, BITS = bench.W, bench.H, bench.LEVELS, bench.BITS
enc, _ = j2kb200.openjpeg_quant_params(L, BITS)
fp = abi.fwd_params(W, H, 1, BITS, False, num_levels=L, reversible=False, htj2k=True, steps=j2kb200.runtime_quant_steps(enc, L, BITS))
kmax = np.array([[(int(e) >> 11) + 1 for e in enc]], np.uint8)
host = bench.synth_frames(2, 2)
pix = ctx.pinned(F * host.shape[1]).reshape(F, -1)
for f in range(F):
    pix[f] = host[f % 2]
cap = int(ctx.lib.j2k_ht_encode_bound(abi.C.byref(fp) if hasattr(abi, "C") else __import__("ctypes").byref(fp), 64, 64, int(kmax.max()), F))
out = ctx.pinned(cap)
for i in range(reps):
    t = time.perf_counter()
    s, r = ctx.forward_ht(fp, pix, kmax, out=out)
    dt = time.perf_counter() - t
    print("call %d: %.2f ms, %.0f Mpixel/s, %d bytes" % (i, dt * 1e3, F * W * H / dt / 1e6, s.size))
