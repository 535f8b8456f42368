#!/usr/bin/env python
"""Input file for tools/ht_micro.cu (kernel-only timing harness of ht_decode_kernel): one frame of 5/3 coefficients of a
synthetic image, HT-coded block by block with the oracle-side generator, written with its block table and records.
Runs in the build container (no GPU): the forward transform here is the ORACLE's (test infrastructure, untimed)."""
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import ht_oracle_lib  # noqa: E402
import ht_parity as HP  # noqa: E402
import oracle_lib  # noqa: E402
from j2kb200 import abi  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
bits = int(sys.argv[2]) if len(sys.argv) > 2 else 12
levels = 5
ht, orc = ht_oracle_lib.HtOracle(), oracle_lib.Oracle()
rng = np.random.default_rng(0)
H = W = size
yy, xx = np.mgrid[0:H, 0:W]
img = ((np.sin(xx / 37.0) + np.cos(yy / 23.0)) * (1 << (bits - 3)) + (1 << (bits - 1)) + rng.normal(0, 6, (H, W))).clip(0, (1 << bits) - 1)
raw = img.astype("<u2").view(np.uint8).reshape(-1)
fp = abi.fwd_params(W, H, 1, 16, False, num_levels=levels, reversible=True, htj2k=True)
co = orc.forward(fp, raw).reshape(H, W)
st, off, ln, km, mm, lay = HP.generated_stream(ht, orc, co, levels, 64, 64, rng, slack=0)
out = os.path.join(ROOT, "tools", "_ht", f"ht_{size}_{bits}.bin")
with open(out, "wb") as f:
    f.write(struct.pack("<qqqqqq", len(lay), W, H, 64, 64, st.size))
    for b in lay:
        f.write(struct.pack("<qqiiiiii", b.y0 * W + b.x0, b.offset, W, b.width, b.height, 0, 0, 0))
    for o, n, k, m in zip(off, ln, km, mm):
        f.write(struct.pack("<QIBBH", int(o), int(n), int(k), int(m), 0))
    f.write(st.tobytes())
    f.write(co.astype("<i4").tobytes())
print(out, len(lay), "blocks", st.size, "bytes", st.size * 8 / (H * W), "bpp")
