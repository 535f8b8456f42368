"""Minimal JPEG 2000 Part-1/Part-15 codestream reader -- TEST INFRASTRUCTURE.

It exists to pull the HT code-block byte segments (and the packet-header facts the block decoder needs: zero bit-planes,
pass count) out of the reference's OpenJPH interop fixtures (test-data/htj2k/interop, copied to tests/golden/htj2k_interop), so
the HT block-decoder oracle and the CUDA decoder can be pinned against third-party codestreams the way the reference's own
test does (jpeg2000/htj2k/interop_manifest_test.go:43-74: decode == input.raw).  Scope: one tile, no SOP/EPH requirement,
default (maximal) precincts, any progression with a single layer -- which is every fixture.  T2 packet parsing stays on the
host in the product as well (SURVEY 8 "stays in Go"); this file is its stand-in for the tests.
Written from ISO/IEC 15444-1 Annex A / B.10; the reference's parser is jpeg2000/codestream + jpeg2000/t2/packet_decoder.go.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field


@dataclass
class Header:
    width: int = 0
    height: int = 0
    components: int = 0
    depth: list = field(default_factory=list)
    signed: list = field(default_factory=list)
    num_levels: int = 0
    cbw: int = 64
    cbh: int = 64
    cb_style: int = 0
    reversible: bool = True
    mct: int = 0
    progression: int = 0
    layers: int = 1
    scod: int = 0
    guard_bits: int = 0
    qstyle: int = 0
    spqcd: bytes = b""


@dataclass
class Block:
    data: bytes = b""
    zero_bitplanes: int = 0
    passes: int = 0
    included: bool = False
    lblock: int = 3


class _Bits:
    """packet-header bit reader: MSB first, a byte after 0xFF carries 7 bits (B.10.1)"""

    def __init__(self, buf, pos):
        self.buf, self.pos, self.cur, self.n = buf, pos, 0, 0

    def bit(self):
        if self.n == 0:
            prev = self.cur
            self.cur = self.buf[self.pos]
            self.pos += 1
            self.n = 7 if prev == 0xFF else 8
        self.n -= 1
        return (self.cur >> self.n) & 1

    def bits(self, k):
        v = 0
        for _ in range(k):
            v = (v << 1) | self.bit()
        return v

    def align(self):
        # B.10.1: the header ends on a byte boundary; a trailing 0xFF is followed by a stuffed 0 byte
        if self.cur == 0xFF:
            self.pos += 1
        self.n = 0
        self.cur = 0
        return self.pos


class _TagTree:
    def __init__(self, w, h):
        self.dims = [(w, h)]
        while self.dims[-1] != (1, 1):
            pw, ph = self.dims[-1]
            self.dims.append(((pw + 1) // 2, (ph + 1) // 2))
        self.val = [[None] * (a * b) for a, b in self.dims]
        self.low = [[0] * (a * b) for a, b in self.dims]

    def decode(self, x, y, threshold, br):
        """True when the leaf's value is known to be < threshold (B.10.2)"""
        path = []
        for lvl, (w, _) in enumerate(self.dims):
            path.append((lvl, (y >> lvl) * w + (x >> lvl)))
        low = 0
        for lvl, idx in reversed(path):
            if low > self.low[lvl][idx]:
                self.low[lvl][idx] = low
            else:
                low = self.low[lvl][idx]
            while low < threshold and self.val[lvl][idx] is None:
                if br.bit():
                    self.val[lvl][idx] = low
                else:
                    low += 1
            self.low[lvl][idx] = low
        lvl, idx = path[0]
        return self.val[lvl][idx] is not None and self.val[lvl][idx] < threshold


def _num_passes(br):
    if not br.bit():
        return 1
    if not br.bit():
        return 2
    n = br.bits(2)
    if n != 3:
        return 3 + n
    n = br.bits(5)
    if n != 31:
        return 6 + n
    return 37 + br.bits(7)


def parse_header(d):
    assert d[:2] == b"\xff\x4f", "no SOC"
    h = Header()
    i = 2
    while True:
        m, L = struct.unpack(">HH", d[i:i + 4])
        seg = d[i + 4:i + 2 + L]
        if m == 0xFF90:  # SOT: first tile part
            return h, i
        if m == 0xFF51:
            _, xs, ys, xo, yo, xt, yt, xto, yto, c = struct.unpack(">HIIIIIIIIH", seg[:36])
            assert xo == 0 and yo == 0 and xto == 0 and yto == 0 and xt >= xs and yt >= ys, "single tile at the origin only"
            h.width, h.height, h.components = xs, ys, c
            for k in range(c):
                ssiz, xr, yr = seg[36 + 3 * k:39 + 3 * k]
                assert xr == 1 and yr == 1
                h.depth.append((ssiz & 0x7F) + 1)
                h.signed.append(bool(ssiz & 0x80))
        elif m == 0xFF52:
            h.scod, h.progression, h.layers, h.mct, h.num_levels, cw, ch, h.cb_style, tr = struct.unpack(">BBHBBBBBB", seg[:10])
            assert (h.scod & 1) == 0, "default precincts only"
            h.cbw, h.cbh, h.reversible = 1 << (cw + 2), 1 << (ch + 2), tr == 1
        elif m == 0xFF5C:
            h.qstyle, h.guard_bits, h.spqcd = seg[0] & 0x1F, seg[0] >> 5, bytes(seg[1:])
        i += 2 + L


def band_kmax(h: Header, band_index: int) -> int:
    """bandNumbpsFromQCD, jpeg2000/t2/bitplane.go:22-61 (band_index: 0 = LL, then HL, LH, HH from the coarsest level)"""
    if h.qstyle == 0:
        return (h.spqcd[band_index] >> 3) + h.guard_bits - 1
    if h.qstyle == 1:
        e = (struct.unpack(">H", h.spqcd[:2])[0] >> 11) & 0x1F
        if band_index > 0:
            e = max(0, e - (band_index - 1) // 3)
        return e + h.guard_bits - 1
    e = (struct.unpack(">H", h.spqcd[2 * band_index:2 * band_index + 2])[0] >> 11) & 0x1F
    return e + h.guard_bits - 1


def tile_body(d, first_sot):
    """concatenated packet bytes of every tile part of tile 0"""
    out = bytearray()
    i = first_sot
    while i + 2 <= len(d):
        m = struct.unpack(">H", d[i:i + 2])[0]
        if m == 0xFFD9:  # EOC
            break
        assert m == 0xFF90, hex(m)
        lsot, isot, psot, tp, ntp = struct.unpack(">HHIBB", d[i + 2:i + 12])
        assert isot == 0
        j = i + 2 + lsot
        while struct.unpack(">H", d[j:j + 2])[0] != 0xFF93:  # tile-part header segments up to SOD
            j += 2 + struct.unpack(">H", d[j + 2:j + 4])[0]
        end = i + psot if psot else len(d) - 2
        out += d[j + 2:end]
        i = end
    return bytes(out)


def parse(d: bytes, layout):
    """layout(width, height, levels, cbw, cbh) -> list of block records with .res, .band, .cbx, .cby (the reference's order:
    resolution 0, then HL, LH, HH per resolution, row-major inside a band).
    Returns (Header, blocks[component][block index in layout order])."""
    h, sot = parse_header(d)
    body = tile_body(d, sot)
    lay = layout(h.width, h.height, h.num_levels, h.cbw, h.cbh)
    # bands of a resolution in packet order, each with its code-block grid
    grids = {}
    for bi, b in enumerate(lay):
        g = grids.setdefault((b.res, b.band), {"nx": 0, "ny": 0, "idx": {}})
        g["nx"], g["ny"] = max(g["nx"], b.cbx + 1), max(g["ny"], b.cby + 1)
        g["idx"][(b.cbx, b.cby)] = bi
    blocks = [[Block() for _ in lay] for _ in range(h.components)]
    trees = {}
    pos = 0
    assert h.layers == 1
    # one precinct per resolution and a single layer: LRCP, RLCP, RPCL, PCRL with default precincts all visit (r, c) like this
    # except CPRL, which is component-major
    order = [(r, c) for r in range(h.num_levels + 1) for c in range(h.components)]
    if h.progression == 4:
        order = [(r, c) for c in range(h.components) for r in range(h.num_levels + 1)]
    for r, c in order:
        if body[pos:pos + 2] == b"\xff\x91":  # SOP
            pos += 6
        br = _Bits(body, pos)
        included = []
        if br.bit():
            for band in ([0] if r == 0 else [1, 2, 3]):
                g = grids.get((r, band))
                if not g:
                    continue
                key = (c, r, band)
                if key not in trees:
                    trees[key] = (_TagTree(g["nx"], g["ny"]), _TagTree(g["nx"], g["ny"]))
                incl, zbp = trees[key]
                for cby in range(g["ny"]):
                    for cbx in range(g["nx"]):
                        blk = blocks[c][g["idx"][(cbx, cby)]]
                        if blk.included:
                            inc = br.bit()
                        else:
                            inc = incl.decode(cbx, cby, 1, br)  # layer 0: threshold 1
                        if not inc:
                            continue
                        if not blk.included:
                            k = 1
                            while not zbp.decode(cbx, cby, k, br):
                                k += 1
                            blk.zero_bitplanes = k - 1
                            blk.included = True
                        blk.passes = _num_passes(br)
                        while br.bit():
                            blk.lblock += 1
                        # Part 15 7.? : the cleanup pass is a segment of its own, SigProp/MagRef share the next one
                        assert blk.passes <= 3, "placeholder passes are not produced by the fixtures"
                        seg1 = br.bits(blk.lblock)
                        seg2 = br.bits(blk.lblock + (1 if blk.passes == 3 else 0)) if blk.passes > 1 else 0
                        included.append((blk, seg1, seg2))
        pos = br.align()
        if body[pos:pos + 2] == b"\xff\x92":  # EPH
            pos += 2
        for blk, seg1, seg2 in included:
            blk.data = body[pos:pos + seg1]  # the cleanup segment: what HTDecoder.Decode consumes
            pos += seg1 + seg2
    return h, blocks
