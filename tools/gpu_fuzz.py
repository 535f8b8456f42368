#!/usr/bin/env python
"""Extra seeded parity sweeps on the GPU (beyond the fixed seeds of tests/test_gpu_parity.py): random geometries through the
C ABI against the oracle, and random batches through the persistent launch with both job orders.
    python tools/gpu_fuzz.py [seed] [cases]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402

import j2kb200  # noqa: E402
import oracle_lib  # noqa: E402
import parity_cases as PC  # noqa: E402

seed = int(sys.argv[1]) if len(sys.argv) > 1 else 777
n = int(sys.argv[2]) if len(sys.argv) > 2 else 600
orc = oracle_lib.Oracle()
t0 = time.time()
bad = 0
with j2kb200.Context(devices=[0]) as ctx:
    for k, case in enumerate(PC.random_geometry_cases(n, seed, 1500, 400)):
        try:
            PC.check_random_case(ctx, orc, case, seed + k)
        except AssertionError as e:
            bad += 1
            print("FAIL geometry", case, e, flush=True)
    rng = np.random.default_rng(seed)
    for k in range(24):
        w = int(rng.integers(1, 12)) * 64; h = int(rng.integers(8, 200)); c = int(rng.choice([1, 1, 3])); rev = bool(rng.integers(0, 2))
        bits = 8 if c == 3 else int(rng.choice([8, 12, 16])); L = int(rng.integers(1, 5)); nf = int(rng.integers(2, 40))
        lag = int(rng.integers(0, 4)); gks = int(rng.choice([16, 64, 256, 1024]))
        try:
            PC.check_pipelined_order(ctx, orc, w, h, c, bits, L, rev, nf, gks, lag)
        except AssertionError as e:
            bad += 1
            print("FAIL batch", (w, h, c, bits, L, rev, nf, gks, lag), e, flush=True)
print("gpu_fuzz seed %d: %d geometry cases + 24 batches, %d failures, %.0f s" % (seed, n, bad, time.time() - t0))
sys.exit(1 if bad else 0)
