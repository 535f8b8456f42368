#!/usr/bin/env python
"""A small parity set that touches every kernel family once, sized for compute-sanitizer (memcheck / synccheck slow kernels
by one to two orders of magnitude):  compute-sanitizer --tool memcheck python tools/gpu_sanitize.py
The GPU pool of this project refuses compute-sanitizer (round 1: "closed on this pool"), so the bounds check that actually
ran is tools/emu_asan.sh: the same kernel sources under AddressSanitizer on the CPU emulator."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import j2kb200  # noqa: E402
import oracle_lib  # noqa: E402
import parity_cases as PC  # noqa: E402

orc = oracle_lib.Oracle()
with j2kb200.Context(devices=[0]) as ctx:
    PC.check_pipeline(ctx, orc, 512, 64, 1, 16, True, 4, True, seed=1)            # 5/3 ring, mono
    PC.check_pipeline(ctx, orc, 512, 64, 1, 12, False, 4, False, seed=2)          # 9/7 ring, mono + quantizer
    PC.check_pipeline(ctx, orc, 256, 48, 3, 8, False, 3, False, seed=3)           # ICT + 9/7 (component-split forward, NP = 2 inverse)
    PC.check_pipeline(ctx, orc, 256, 48, 3, 8, False, 3, True, seed=4)            # RCT + 5/3
    PC.check_pipeline(ctx, orc, 131, 77, 1, 12, False, 3, False, seed=5)          # per-level kernels (odd geometry)
    PC.check_pipeline(ctx, orc, 520, 40, 1, 16, False, 4, False, seed=6)          # hybrid plan
    PC.check_pipeline(ctx, orc, 200, 120, 3, 8, False, 3, False, tile=(64, 64), seed=7)  # tiles, several classes
    PC.check_blocks(ctx, orc, 200, 120, 1, 12, 3, False, cb=(32, 32))
    PC.check_blocks_roi(ctx, orc, 128, 96, 3, 8, 2, True, [4, 0, 7], cb=(16, 16))
    PC.check_pipelined_order(ctx, orc, 264, 40, 1, 12, 3, False, 12, 16, 2)
    PC.check_custom_mct(ctx, orc, 80, 48, 8, 2, True, "bindings")
    PC.check_wavelet_api(ctx, orc, 130, 70, 5, 1, 0)
    PC.check_package_api(ctx, orc, n=5003)
print("sanitize set done")
