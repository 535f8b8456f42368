"""HTJ2K block-decoder parity checks shared by the emulator (CPU) and the GPU suites: the product's j2k_ht_decode_blocks /
j2k_inverse_ht through the C ABI against oracle/ht_oracle.c, bit for bit, status codes included."""
from __future__ import annotations

import numpy as np

import ht_cases
from j2kb200 import abi
from j2kb200.codec import Context


def _inv_params(h):
    return abi.inv_params(h.width, h.height, h.components, h.depth[0], h.signed[0], num_levels=h.num_levels, reversible=True, htj2k=True,
                          mct_mode=abi.MCT_RCT if h.mct else abi.MCT_NONE)


def oracle_blocks(ht, fx, stream=None, kmax=None, mmsb=None, lengths=None):
    h = fx["header"]
    return ht.decode_blocks(fx["stream"] if stream is None else stream, fx["offsets"], fx["lengths"] if lengths is None else lengths,
                            fx["kmax"] if kmax is None else kmax, fx["mmsb"] if mmsb is None else mmsb, fx["widths"], fx["heights"],
                            fx["out_offsets"], h.components * fx["plane_samples"])


def check_fixture(ctx, ht, oracle, name, kind, nframes=1):
    """decode == oracle for every block, and the whole tail on the device == input.raw (interop_manifest_test.go:43-74)"""
    fx = ht_cases.load(name, kind, oracle.codeblock_layout)
    h = fx["header"]
    ip = _inv_params(h)
    want, wst = oracle_blocks(ht, fx)
    assert not wst.any()
    nb = len(fx["offsets"])
    # nframes copies of the frame: the records of frame f point into the same stream
    rec = Context.ht_records(np.tile(fx["offsets"], nframes), np.tile(fx["lengths"], nframes), np.tile(fx["kmax"], nframes),
                             np.tile(fx["mmsb"], nframes))
    got, st = ctx.ht_decode_blocks(ip, nframes, fx["stream"], rec, h.cbw, h.cbh)
    assert st.shape == (nframes, nb) and not st.any()
    for f in range(nframes):
        assert np.array_equal(got[f], want), (name, kind, f)
    px, st = ctx.inverse_ht(ip, nframes, fx["stream"], rec, h.cbw, h.cbh)
    for f in range(nframes):
        assert np.array_equal(px[f], fx["raw"]), (name, kind, f)


def check_mutations(ctx, ht, oracle, name, kind, rounds=3, seed=0):
    """corrupted segments and coding contexts: whatever the reference decoder makes of them (values, zeros + error), bit for bit"""
    fx = ht_cases.load(name, kind, oracle.codeblock_layout)
    h = fx["header"]
    ip = _inv_params(h)
    rng = np.random.default_rng(seed)
    for r in range(rounds):
        stream = fx["stream"].copy()
        n = max(1, stream.size // (20 if r else 200))
        idx = rng.integers(0, stream.size, n)
        stream[idx] = rng.integers(0, 256, n).astype(np.uint8) if r != 1 else 0xFF
        kmax, mmsb, lengths = fx["kmax"].copy(), fx["mmsb"].copy(), fx["lengths"].copy()
        pick = rng.random(kmax.size) < 0.3
        mmsb[pick] = rng.integers(0, 32, int(pick.sum())).astype(np.uint8)
        pick = rng.random(kmax.size) < 0.2
        kmax[pick] = rng.integers(0, 34, int(pick.sum())).astype(np.uint8)
        pick = (rng.random(kmax.size) < 0.2) & (lengths > 4)
        lengths[pick] = (lengths[pick] * rng.random(int(pick.sum()))).astype(np.uint32)   # truncated segments
        want, wst = oracle_blocks(ht, fx, stream, kmax, mmsb, lengths)
        got, st = ctx.ht_decode_blocks(ip, 1, stream, Context.ht_records(fx["offsets"], lengths, kmax, mmsb), h.cbw, h.cbh)
        assert np.array_equal(st[0], wst), (name, kind, r, np.flatnonzero(st[0] != wst)[:8])
        assert np.array_equal(got[0], want), (name, kind, r)


def check_random_streams(ctx, ht, oracle, w, h, levels, cbw, cbh, seed, mmsb_lo=18, mean_len=None):
    """random bytes as cleanup segments, generous missing-MSB counts so that most blocks decode to the end: every table entry,
    the MEL state machine, both un-stuffing rules and the exponent predictor at every block shape"""
    rng = np.random.default_rng(seed)
    lay = oracle.codeblock_layout(w, h, levels, cbw, cbh)
    nb = len(lay)
    lengths = np.array([int(rng.integers(2, 3 * b.width * b.height + 8)) if mean_len is None else mean_len for b in lay], np.uint32)
    lengths[rng.random(nb) < 0.1] = 0
    offsets = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.uint64)
    stream = rng.integers(0, 256, int(lengths.sum()) + 16).astype(np.uint8)
    stream[rng.random(stream.size) < 0.05] = 0xFF
    # a plausible Scup in most blocks: the last byte and the low nibble of the one before hold the suffix length
    for o, n in zip(offsets, lengths):
        if n >= 4 and rng.random() < 0.9:
            scup = int(rng.integers(2, min(int(n), 4079) + 1))
            stream[int(o) + n - 1] = scup >> 4
            stream[int(o) + n - 2] = (stream[int(o) + n - 2] & 0xF0) | (scup & 0xF)
    kmax = rng.integers(1, 32, nb).astype(np.uint8)
    mmsb = rng.integers(mmsb_lo, 30, nb).astype(np.uint8)
    widths = np.array([b.width for b in lay], np.int32)
    heights = np.array([b.height for b in lay], np.int32)
    out_off = np.array([b.offset for b in lay], np.int64)
    want, wst = ht.decode_blocks(stream, offsets, lengths, kmax, mmsb, widths, heights, out_off, w * h)
    ip = abi.inv_params(w, h, 1, 16, False, num_levels=levels, reversible=True, htj2k=True)
    got, st = ctx.ht_decode_blocks(ip, 1, stream, Context.ht_records(offsets, lengths, kmax, mmsb), cbw, cbh)
    assert np.array_equal(st[0], wst), (np.flatnonzero(st[0] != wst)[:8], st[0][:16], wst[:16])
    assert np.array_equal(got[0], want)
    return int((wst == 0).sum()), nb


def check_error_codes(ctx, ht, oracle):
    w = h = 8
    ip = abi.inv_params(w, h, 1, 8, False, num_levels=0, reversible=True, htj2k=True)
    stream = np.zeros(64, np.uint8)
    for length, kmax, mmsb, want in ((0, 10, 3, abi.HT_OK), (4, 0, 3, abi.HT_ERR_KMAX), (4, 10, 30, abi.HT_ERR_SEGMENT), (1, 10, 3, abi.HT_ERR_SEGMENT),
                                     (4, 10, 3, abi.HT_ERR_SEGMENT)):
        got, st = ctx.ht_decode_blocks(ip, 1, stream, Context.ht_records([0], [length], [kmax], [mmsb]), 8, 8)
        assert st[0, 0] == want and not got.any(), (length, kmax, mmsb, st)
        if length != 1:   # the Go decoder indexes codeblock[lcup-2]: a one-byte segment panics there, the oracle refuses it the same way
            rc, blk = ht.decode_block(bytes(stream[:length]), w, h, kmax, mmsb)
            assert rc == want and not blk.any()


def generated_stream(ht, oracle, plane, levels, cbw, cbh, rng, slack=2):
    """HT-code every code-block of a Mallat coefficient plane with the oracle-side generator (round trip checked through the
    pinned decoder) -> (stream, offsets, lengths, kmax, mmsb, layout)"""
    h, w = plane.shape
    lay = oracle.codeblock_layout(w, h, levels, cbw, cbh)
    chunks, offsets, lengths, kmax, mmsb = [], [], [], [], []
    pos = 0
    for b in lay:
        blk = np.ascontiguousarray(plane[b.y0:b.y0 + b.height, b.x0:b.x0 + b.width])
        need = max(int(np.abs(blk).max()).bit_length() - 1, 0)        # |x| <= 2^(mmsb + 1)
        mm = min(29, need + int(rng.integers(0, slack + 1)))
        data = ht.encode_block(blk, mm)
        if data:
            rc, back = ht.decode_block(data, b.width, b.height, mm + 1, mm)
            assert rc == 0 and np.array_equal(back, blk), "generator round trip"
        offsets.append(pos); lengths.append(len(data)); kmax.append(mm + 1); mmsb.append(mm)
        chunks.append(data); pos += len(data)
    stream = np.frombuffer(b"".join(chunks) + b"\0" * 16, np.uint8)
    return (stream, np.array(offsets, np.uint64), np.array(lengths, np.uint32), np.array(kmax, np.uint8), np.array(mmsb, np.uint8), lay)


def check_generated(ctx, ht, oracle, w, h, levels, cbw, cbh, bits, density, seed, components=1, nframes=1, reversible=True):
    """a random coefficient image, HT-coded block by block: decode == the coefficients (block-major), and the whole tail
    (decode + assembleSubbands + inverse transform) == the oracle's inverse of the same planes"""
    rng = np.random.default_rng(seed)
    C_ = components
    frames, streams, recs = [], [], []
    base = 0
    for f in range(nframes):
        planes = []
        for c in range(C_):
            # sparse / heavy-tailed magnitudes, denser and larger towards the LL corner like real sub-bands
            mag = np.floor(np.abs(rng.laplace(0, (1 << bits) / 6.0, (h, w)))).astype(np.int64)
            mag = np.minimum(mag, (1 << bits) - 1) * (rng.random((h, w)) < density)
            planes.append((mag * rng.choice([-1, 1], (h, w))).astype(np.int32))
        frames.append(planes)
        for c in range(C_):
            st, off, ln, km, mm, lay = generated_stream(ht, oracle, planes[c], levels, cbw, cbh, rng)
            streams.append(st[:-16]); recs.append((off + base, ln, km, mm)); base += st.size - 16
    stream = np.concatenate(streams + [np.zeros(16, np.uint8)])
    rec = Context.ht_records(*[np.concatenate([r[i] for r in recs]) for i in range(4)])
    depth = 8 if bits <= 8 else 16
    steps = None
    if reversible:
        ip = abi.inv_params(w, h, C_, depth, False, num_levels=levels, reversible=True, htj2k=True, mct_mode=abi.MCT_RCT if C_ == 3 else abi.MCT_NONE)
    else:
        enc, _ = oracle.openjpeg_quant_params(levels, depth)
        ip = abi.inv_params(w, h, C_, depth, False, num_levels=levels, reversible=False, htj2k=True, mct_mode=abi.MCT_ICT if C_ == 3 else abi.MCT_NONE,
                            steps=oracle.decode_quant_steps(enc, levels, depth))
    got, st = ctx.ht_decode_blocks(ip, nframes, stream, rec, cbw, cbh)
    assert not st.any()
    for f in range(nframes):
        want_blocks = np.concatenate([oracle.gather_blocks(frames[f][c], levels, cbw, cbh, htj2k=True)[0] for c in range(C_)])
        assert np.array_equal(got[f], want_blocks), f
    px, st = ctx.inverse_ht(ip, nframes, stream, rec, cbw, cbh)
    for f in range(nframes):
        co = np.concatenate([p.reshape(-1) for p in frames[f]])
        assert np.array_equal(px[f], oracle.inverse(ip, co)), f


def band_kmax_table(components, levels, base):
    """a plausible bandNumbps table: `base` bits + the usual sub-band gains"""
    t = np.zeros((components, 3 * levels + 1), np.uint8)
    for c in range(components):
        t[c, 0] = base
        for r in range(1, levels + 1):
            t[c, 1 + 3 * (r - 1):4 + 3 * (r - 1)] = (base + 1, base + 1, base + 2)
    return np.minimum(t, 30)


def check_encode(ctx, ht, oracle, w, h, comps, bits, levels, cbw, cbh, seed, nframes=1, reversible=True, base=None, tile=None, ordered=True):
    """forward transform + HT block encoding on the device: every block's bytes == the oracle's restatement of HTEncoder.Encode
    applied to the oracle's coefficients of the same frame; records consistent; the stream decodes back (device decoder).
    ordered=False: a context with several devices appends the sub-batches as it collects them, so only the records locate the
    segments; they must still tile the stream exactly (disjoint, no gaps)."""
    rng = np.random.default_rng(seed)
    depth = 8 if bits <= 8 else 16
    bpp = 1 if depth == 8 else 2
    yy, xx = np.mgrid[0:h, 0:w]
    frames = []
    for f in range(nframes):
        planes = [((np.sin((xx + 31 * f) / (9.0 + c)) * np.cos(yy / (7.0 + c)) * 0.35 + 0.5) * ((1 << bits) - 1) + rng.normal(0, (1 << bits) / 64.0, (h, w))).clip(0, (1 << bits) - 1)
                  for c in range(comps)]
        if f % 2 == 1:
            planes[0][: h // 2, : w // 2] = 0   # flat regions: empty blocks, MEL runs
        px = np.stack(planes, axis=-1).astype("<u2" if bpp == 2 else np.uint8)
        frames.append(px.reshape(-1).view(np.uint8))
    frames = np.ascontiguousarray(np.stack(frames))
    kw = dict(num_levels=levels, reversible=reversible, htj2k=True)
    if tile:
        kw.update(tile_width=tile[0], tile_height=tile[1])
    if reversible:
        kw["mct_mode"] = abi.MCT_RCT if comps == 3 else abi.MCT_NONE
    else:
        enc, _ = oracle.openjpeg_quant_params(levels, bits)
        kw["steps"] = oracle.runtime_quant_steps(enc, levels, bits)
        kw["mct_mode"] = abi.MCT_ICT if comps == 3 else abi.MCT_NONE
    fp = abi.fwd_params(w, h, comps, bits, False, **kw)
    kmax = band_kmax_table(comps, levels, base if base is not None else bits + 2)
    stream, rec = ctx.forward_ht(fp, frames, kmax, cbw, cbh)
    nblk = rec.size // nframes
    # the oracle's coefficients and block layout (per tile, per component)
    pos = 0
    spans = []
    for f in range(nframes):
        co = oracle.forward(fp, frames[f])
        k = f * nblk
        off = 0
        ntiles = 1
        tiles = [(w, h)]
        if tile:
            tiles = []
            for ty in range(0, h, tile[1]):
                for tx in range(0, w, tile[0]):
                    tiles.append((min(tile[0], w - tx), min(tile[1], h - ty)))
        for (tw, th) in tiles:
            lay = oracle.codeblock_layout(tw, th, levels, cbw, cbh)
            for c in range(comps):
                plane = co[off:off + tw * th].reshape(th, tw)
                off += tw * th
                for b in lay:
                    idx = 0 if b.res == 0 else 1 + 3 * (b.res - 1) + (b.band - 1)
                    km = int(kmax[c, idx])
                    want = ht.encode_ref(plane[b.y0:b.y0 + b.height, b.x0:b.x0 + b.width], km)
                    r = rec[k]
                    assert int(r["kmax"]) == km and int(r["missing_msbs"]) == km - 1, (f, k)
                    assert int(r["length"]) == len(want), (f, k, b.width, b.height, int(r["length"]), len(want))
                    if want:
                        ro = int(r["offset"])
                        if ordered:
                            assert ro == pos, (f, k)
                        assert ro + len(want) <= stream.size, (f, k)
                        got = stream[ro:ro + len(want)].tobytes()
                        assert got == want, (f, k, b.width, b.height, [i for i in range(len(want)) if got[i] != want[i]][:6])
                        pos += len(want)
                        spans.append((ro, len(want)))
                    k += 1
        assert k == (f + 1) * nblk
    assert pos == stream.size
    spans.sort()
    assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1)) and (not spans or spans[0][0] == 0)
    return stream, rec, fp, frames


def check_fixture_encode(ctx, oracle, name, kind):
    """input.raw -> forward 5/3 (+ RCT) + HT block encoding on the device == the code-block bytes OpenJPH wrote into the
    fixture codestream (the reference's byte-parity test, htj2k/go_byte_parity_test.go:11-44, below the packet layer)"""
    import j2c_parse
    fx = ht_cases.load(name, kind, oracle.codeblock_layout)
    h = fx["header"]
    L = h.num_levels
    fp = abi.fwd_params(h.width, h.height, h.components, h.depth[0], h.signed[0], num_levels=L, reversible=True, htj2k=True,
                        mct_mode=abi.MCT_RCT if h.mct else abi.MCT_NONE)
    kmax = np.array([[j2c_parse.band_kmax(h, i) for i in range(3 * L + 1)] for _ in range(h.components)], np.uint8)
    stream, rec = ctx.forward_ht(fp, fx["raw"].reshape(1, -1), kmax, h.cbw, h.cbh)
    assert rec.size == len(fx["offsets"])
    for k in range(rec.size):
        o, n = int(fx["offsets"][k]), int(fx["lengths"][k])
        assert int(rec[k]["length"]) == n, (k, int(rec[k]["length"]), n)
        if n:
            assert int(rec[k]["kmax"]) == int(fx["kmax"][k]) and int(rec[k]["missing_msbs"]) == int(fx["mmsb"][k])
            ro = int(rec[k]["offset"])
            assert stream[ro:ro + n].tobytes() == fx["stream"][o:o + n].tobytes(), k
