#!/usr/bin/env python
"""tests/golden/ht_blocks_64.npz: 256 HT-coded 64 x 64 code-blocks of a C2-shaped frame (bench.synth_frames -> 9/7 + OpenJPEG
default quantization through the ORACLE), coded with the oracle-side generator, each round-tripped through the pinned
decoder oracle, with the coefficients they decode to.  bench.py's `ht_decode` leg assembles C2 frames from these blocks and
checks the device decoder against the stored coefficients, so the bench itself never touches oracle/.
Run in the build container:  python tools/make_ht_fixture.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "go-dicom-codec_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import bench  # noqa: E402
import ht_oracle_lib  # noqa: E402
import oracle_lib  # noqa: E402
from j2kb200 import abi  # noqa: E402

ht, orc = ht_oracle_lib.HtOracle(), oracle_lib.Oracle()
W, H, L, BITS = bench.W, bench.H, bench.LEVELS, bench.BITS
raw = bench.synth_frames(1, 1234)[0]
enc, _ = orc.openjpeg_quant_params(L, BITS)
fp = abi.fwd_params(W, H, 1, BITS, False, num_levels=L, reversible=False, htj2k=True, steps=orc.runtime_quant_steps(enc, L, BITS))
co = orc.forward(fp, raw).reshape(H, W)
lay = orc.codeblock_layout(W, H, L, 64, 64)
assert all(b.width == 64 and b.height == 64 for b in lay) and len(lay) == 4096
pick = list(range(0, len(lay), 16))   # every 16th block: all resolutions, all three orientations
chunks, offsets, lengths, kmax, mmsb, coeffs = [], [], [], [], [], []
pos = 0
for i in pick:
    b = lay[i]
    blk = np.ascontiguousarray(co[b.y0:b.y0 + 64, b.x0:b.x0 + 64])
    mm = max(int(np.abs(blk).max()).bit_length() - 1, 0)
    data = ht.encode_block(blk, mm)
    rc, back = ht.decode_block(data, 64, 64, mm + 1, mm)
    assert rc == 0 and np.array_equal(back, blk)
    offsets.append(pos); lengths.append(len(data)); kmax.append(mm + 1); mmsb.append(mm); coeffs.append(blk.astype(np.int32))
    chunks.append(data); pos += len(data)
out = os.path.join(ROOT, "tests", "golden", "ht_blocks_64.npz")
np.savez_compressed(out, stream=np.frombuffer(b"".join(chunks), np.uint8), offsets=np.array(offsets, np.uint64), lengths=np.array(lengths, np.uint32),
                    kmax=np.array(kmax, np.uint8), mmsb=np.array(mmsb, np.uint8), coeffs=np.stack(coeffs))
print(out, os.path.getsize(out), "bytes;", pos, "stream bytes;", pos * 8 / (len(pick) * 4096), "bits per sample")
