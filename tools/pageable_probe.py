#!/usr/bin/env python
"""Blocking j2k_forward_batch on PAGEABLE numpy buffers (what a Go []byte is) for the bench workload, one setting of the
staging ring per process:  J2K_STAGE_THREADS=8 J2K_STAGE_CHUNK_KB=4096 python tools/pageable_probe.py [frames]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import j2kb200
from j2kb200 import abi

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
W = H = 4096
enc, _ = j2kb200.openjpeg_quant_params(6, 12)
es = j2kb200.runtime_quant_steps(enc, 6, 12)
fp = abi.fwd_params(W, H, 1, 12, False, num_levels=6, reversible=False, steps=es)
rng = np.random.default_rng(1)
n_in = rng.integers(0, 4096, (B, W * H), dtype=np.uint16).view(np.uint8).reshape(B, -1)
n_out = np.empty((B, W * H), np.int32)
with j2kb200.Context(devices=[0]) as ctx:
    ctx.forward_batch(fp, n_in, n_out)
    t0 = time.perf_counter()
    K = 3
    for _ in range(K):
        ctx.forward_batch(fp, n_in, n_out)
    dt = (time.perf_counter() - t0) / K
    # host memcpy rate of this box, one thread, same volume (for scale)
    a = np.empty(B * W * H * 4, np.uint8); b = np.empty_like(a); b[:] = 1
    t1 = time.perf_counter(); a[:] = b; t2 = time.perf_counter()
print(json.dumps({"threads": os.environ.get("J2K_STAGE_THREADS"), "chunk_kb": os.environ.get("J2K_STAGE_CHUNK_KB"), "frames": B,
                  "Mpixel_s": B * W * H / dt / 1e6, "ms": dt * 1e3, "host_GBps_moved": B * W * H * 6 / dt / 1e9,
                  "numpy_copy_GBps_1thread": a.nbytes / (t2 - t1) / 1e9, "cores": os.cpu_count()}))
