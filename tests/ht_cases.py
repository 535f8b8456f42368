"""Shared by the HTJ2K tests: the OpenJPH interop fixtures as code-block descriptor tables (TEST INFRASTRUCTURE)."""
from __future__ import annotations

import json
import os

import numpy as np

import j2c_parse

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def fixtures():
    man = json.load(open(os.path.join(GOLD, "interop", "manifest.json")))
    out = []
    for fx in man["fixtures"]:
        for kind in ("htj2k_lossless", "htj2k_lossless_rpcl"):
            out.append((fx["name"], "fo_" + kind))
    return out


def load(name, kind, layout):
    """-> dict(header, raw, stream, offsets, lengths, kmax, mmsb, widths, heights, out_offsets, plane_samples, blocks)
    Blocks in the order of the code-block interface: component-major, then j2k_codeblock_layout order."""
    d = open(os.path.join(GOLD, "htj2k_interop", f"{name}__{kind}.j2c"), "rb").read()
    raw = np.fromfile(os.path.join(GOLD, "interop", name + ".raw"), np.uint8)
    h, blocks = j2c_parse.parse(d, layout)
    lay = layout(h.width, h.height, h.num_levels, h.cbw, h.cbh)
    plane = h.width * h.height
    stream = bytearray()
    offsets, lengths, kmax, mmsb, widths, heights, out_offsets = [], [], [], [], [], [], []
    for c in range(h.components):
        for b, blk in zip(lay, blocks[c]):
            band_index = 0 if b.res == 0 else 1 + 3 * (b.res - 1) + (b.band - 1)
            offsets.append(len(stream))
            lengths.append(len(blk.data))
            stream += blk.data
            kmax.append(j2c_parse.band_kmax(h, band_index))
            mmsb.append(blk.zero_bitplanes)
            widths.append(b.width)
            heights.append(b.height)
            out_offsets.append(c * plane + b.offset)
    stream += b"\0" * 16
    return dict(header=h, raw=raw, stream=np.frombuffer(bytes(stream), np.uint8), offsets=np.array(offsets, np.uint64),
                lengths=np.array(lengths, np.uint32), kmax=np.array(kmax, np.uint8), mmsb=np.array(mmsb, np.uint8),
                widths=np.array(widths, np.int32), heights=np.array(heights, np.int32), out_offsets=np.array(out_offsets, np.int64),
                plane_samples=plane, blocks=blocks, layout=lay)
