"""Parity checks shared by the CPU-emulator suite (tests/test_emu_parity.py, small sizes) and the
GPU suite (tests/test_gpu_parity.py, through the real libj2kb200.so).  Every check compares the
product path with the oracle on the same seeded input: bit-exact for all integer outputs
(5/3, RCT, DC shift, quantized 9/7 coefficients, packed pixels) and for the float32 9/7 planes."""
import os
import sys

import numpy as np

from j2kb200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import np_mirror as M  # noqa: E402


def synth(rng, h, w, c, bits, signed=False, kind="smooth"):
    if kind == "noise":
        v = rng.integers(0, 2 ** bits, (h, w, c))
    else:
        yy, xx = np.mgrid[0:h, 0:w]
        base = 2 ** (bits - 1) + 2 ** (bits - 2) * np.sin(xx / 17.0) * np.cos(yy / 23.0)
        v = np.clip(np.rint(base[..., None] + rng.normal(0, 2 ** bits / 64, (h, w, c))), 0, 2 ** bits - 1)
    v = v.astype(np.int64)
    if signed:
        v = v - 2 ** (bits - 1)
        v = np.where(v < 0, v + 2 ** bits, v)  # un-sign-extended storage (decoder.go:807-809)
    return v.astype(np.uint8 if bits <= 8 else "<u2")


def raw_bytes(a):
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


def steps_for(oracle, L, bits, kind="openjpeg", quality=80):
    enc, _ = (oracle.openjpeg_quant_params(L, bits) if kind == "openjpeg" else oracle.quality_quant_params(quality, L, bits))
    return oracle.runtime_quant_steps(enc, L, bits), oracle.decode_quant_steps(enc, L, bits, False)


def check_wavelet_api(ctx, oracle, w, h, levels, x0, y0, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.integers(-30000, 30000, (h, w)).astype(np.int32)
    f = ctx.dwt53_forward(a, levels, x0, y0)
    assert np.array_equal(f, oracle.fwd53(a, levels, x0, y0)), "5/3 forward"
    assert np.array_equal(ctx.dwt53_inverse(f, levels, x0, y0), a), "5/3 round trip"
    y = rng.integers(-30000, 30000, (h, w)).astype(np.int32)
    assert np.array_equal(ctx.dwt53_inverse(y, levels, x0, y0), oracle.inv53(y, levels, x0, y0)), "5/3 inverse"
    b = (rng.standard_normal((h, w)) * 3000).astype(np.float32)
    g = ctx.dwt97_forward(b, levels, x0, y0)
    assert np.array_equal(g.view(np.uint32), oracle.fwd97(b, levels, x0, y0).view(np.uint32)), "9/7 forward"
    r = ctx.dwt97_inverse(b, levels, x0, y0)
    assert np.array_equal(r.view(np.uint32), oracle.inv97(b, levels, x0, y0).view(np.uint32)), "9/7 inverse"


def fwd_inv_params(w, h, c, bits, signed, L, reversible, oracle, tile=(0, 0), htj2k=False, mct=None, fuse=False,
                   steps_kind="openjpeg", **kw):
    if mct is None:
        mct = (abi.MCT_RCT if reversible else abi.MCT_ICT) if c == 3 else abi.MCT_NONE
    es = ds = None
    if not reversible:
        es, ds = steps_for(oracle, L, bits, steps_kind)
    fp = abi.fwd_params(w, h, c, bits, signed, tile[0], tile[1], L, reversible, htj2k, mct, es, fuse_t1_shift=fuse, **kw)
    imct = mct
    ip = abi.inv_params(w, h, c, bits, signed, tile[0], tile[1], L, reversible, htj2k, imct, ds, fuse_t1_halve=fuse)
    return fp, ip


def check_pipeline(ctx, oracle, w, h, c, bits, signed, L, reversible, tile=(0, 0), htj2k=False, fuse=False, kind="smooth",
                   seed=1, mct=None, steps_kind="openjpeg", want_planes=True, identity=True):
    rng = np.random.default_rng(seed)
    img = synth(rng, h, w, c, bits, signed, kind)
    raw = raw_bytes(img)
    fp, ip = fwd_inv_params(w, h, c, bits, signed, L, reversible, oracle, tile, htj2k, mct, fuse, steps_kind)
    got = ctx.forward(fp, raw)
    want = oracle.forward(fp, raw)
    nd = int(np.count_nonzero(got != want))
    assert nd == 0, f"forward: {nd} differing coefficients (max abs {np.abs(got.astype(np.int64) - want).max()})"
    # what T1 hands back
    if reversible:
        back_in = ((want >> 6) * 2) if (fuse and not htj2k) else want
        if fuse and not htj2k:
            assert np.all((want & 63) == 0)
    else:
        back_in = M.t1_emulate(want, htj2k)
    if want_planes:
        px, planes = ctx.inverse(ip, back_in, want_planes=True)
        opx, oplanes = oracle.inverse(ip, back_in, want_planes=True)
        assert np.array_equal(planes, oplanes), "inverse planes (GetImageData)"
        if c == 3 and not reversible:  # without the planes the level-1 inverse may take the float32 inverse-ICT fast path
            assert np.array_equal(ctx.inverse(ip, back_in), opx), "inverse pixels (no planes requested)"
    else:
        px = ctx.inverse(ip, back_in)
        opx = oracle.inverse(ip, back_in)
    nd = int(np.count_nonzero(px != opx))
    assert nd == 0, f"inverse: {nd} differing bytes"
    if reversible and identity:
        assert np.array_equal(px, raw), "lossless identity"
    return got, px


# ---- Part-2 custom MCT / binding lists (encoder.go:465-665, decoder.go:630-735), planar entry, package APIs

CUSTOM_CASES = {
    # name: (components, fwd mode, fwd kwargs, inv mode, inv kwargs)
    "int_matrix": (3, abi.MCT_CUSTOM_INT, dict(mct_matrix=[[1, 1, 0], [0, 1, -1], [2, 0, 1]], mct_offsets=[3, -7, 0]),
                   abi.MCT_CUSTOM_FLOAT, dict(mct_matrix=[[0.5, -0.25, 0.125], [0.0, 1.0, 0.5], [-1.0, 0.5, 0.75]], mct_offsets=[3, -7, 0])),
    "q13_matrix": (3, abi.MCT_CUSTOM_Q13, dict(mct_matrix=[[0.299, 0.587, 0.114], [-0.16875, -0.33126, 0.5], [0.5, -0.41869, -0.08131]]),
                   abi.MCT_CUSTOM_FLOAT, dict(mct_matrix=[[1.0, 0.0, 1.402], [1.0, -0.34413, -0.71414], [1.0, 1.772, 0.0]])),
    "q13_4comp": (4, abi.MCT_CUSTOM_Q13, dict(mct_matrix=[[0.25, 0.25, 0.25, 0.25], [1, -1, 0, 0], [0, 1, -1, 0], [0.5, 0, 0, -0.5]],
                                              mct_offsets=[1, 2, 3, 4]),
                  abi.MCT_CUSTOM_FLOAT, dict(mct_matrix=[[1, 0.75, 0.5, 0.5], [1, -0.25, 0.5, 0.5], [1, -0.25, -0.5, 0.5], [1, 0.75, 0.5, -1.5]],
                                             mct_offsets=[1, 2, 3, 4])),
}


def binding_cases():
    fb = [abi.make_binding((0, 1, 2), [[1, 0, 1], [0, 1, 0], [-1, 0, 2]], [5, 0, -5], element_type=0),
          abi.make_binding((2, 0), [[0.5, 0.5], [1.0, -1.0]], None, element_type=1),
          abi.make_binding((), None, [1, 1, 1], element_type=0)]
    ib = [abi.make_binding((2, 0), [[1.0, 0.5], [1.0, -0.5]], None, element_type=1),
          abi.make_binding((0, 1, 2), [[2, 0, -1], [0, 1, 0], [1, 0, 1]], [5, 0, -5], element_type=0),
          abi.make_binding((1,), None, [9], element_type=0)]
    return fb, ib


def check_custom_mct(ctx, oracle, w, h, bits, L, reversible, case, tile=(0, 0), seed=3):
    """Forward and inverse pipelines with Part-2 transforms; the inverse is checked on the forward output (not an
    identity: the reference's custom inverse is float64 + math.Round, decoder.go:696-723)."""
    rng = np.random.default_rng(seed)
    es = ds = None
    if not reversible:
        es, ds = steps_for(oracle, L, bits)
    if case == "bindings":
        C = 3
        fb, ib = binding_cases()
        fp = abi.fwd_params(w, h, C, bits, False, tile[0], tile[1], L, reversible, False, abi.MCT_BINDINGS, es, bindings=fb)
        ip = abi.inv_params(w, h, C, bits, False, tile[0], tile[1], L, reversible, False, abi.MCT_BINDINGS, ds, bindings=ib)
    else:
        C, fm, fkw, im, ikw = CUSTOM_CASES[case]
        fp = abi.fwd_params(w, h, C, bits, False, tile[0], tile[1], L, reversible, False, fm, es, **fkw)
        ip = abi.inv_params(w, h, C, bits, False, tile[0], tile[1], L, reversible, False, im, ds, **ikw)
    raw = raw_bytes(synth(rng, h, w, C, bits, False, "noise"))
    got, want = ctx.forward(fp, raw), oracle.forward(fp, raw)
    assert np.array_equal(got, want), f"custom MCT forward ({case})"
    back_in = want if reversible else M.t1_emulate(want, False)
    px, planes = ctx.inverse(ip, back_in, want_planes=True)
    opx, oplanes = oracle.inverse(ip, back_in, want_planes=True)
    assert np.array_equal(planes, oplanes), f"custom MCT inverse planes ({case})"
    assert np.array_equal(px, opx), f"custom MCT inverse pixels ({case})"


def check_planar(ctx, oracle, w, h, c, bits, signed, L, reversible, seed=4):
    """EncodeComponents entry (encoder.go:221-273): planar int32 input, no byte conversion."""
    rng = np.random.default_rng(seed)
    lo, hi = (-(2 ** (bits - 1)), 2 ** (bits - 1)) if signed else (0, 2 ** bits)
    planes = [rng.integers(lo, hi, (h, w)).astype(np.int32) for _ in range(c)]
    fp, _ = fwd_inv_params(w, h, c, bits, signed, L, reversible, oracle)
    want = oracle.forward_planar(fp, planes)
    assert np.array_equal(ctx.forward_planar(fp, planes), want), "planar forward"
    # the cgo-callable twin: one buffer, component c at c * plane_stride (tight, and with a padded stride)
    flat = np.stack([p.reshape(-1) for p in planes])
    assert np.array_equal(ctx.forward_planar_flat(fp, flat), want), "flat planar forward"
    padded = np.zeros((c, w * h + 24), np.int32)
    padded[:, :w * h] = flat
    assert np.array_equal(ctx.forward_planar_flat(fp, padded), want), "flat planar forward, padded stride"


def check_package_api(ctx, oracle, n=100_003, seed=6):
    """colorspace / quantization package functions (rct.go, ict.go, quantization.go:310-340, dwt97.go:473-503)."""
    rng = np.random.default_rng(seed)
    r, g, b = (rng.integers(-40000, 70000, n).astype(np.int32) for _ in range(3))
    for name in ("rct_forward", "rct_inverse", "ict_forward", "ict_inverse"):
        got, want = getattr(ctx, name)(r, g, b), getattr(oracle, name)(r, g, b)
        for k in range(3):
            assert np.array_equal(got[k], want[k]), f"{name} component {k}"
    y, cb, cr = oracle.rct_forward(r, g, b)
    back = ctx.rct_inverse(y, cb, cr)
    assert all(np.array_equal(u, v) for u, v in zip(back, (r, g, b))), "RCT identity"
    co = rng.integers(-2 ** 20, 2 ** 20, n).astype(np.int32)
    co[:8] = [0, 1, -1, 3, -3, 5, 7, -7]  # ties at step 2
    for step in (1.0, 2.0, 0.5, 3.7, 0.0078125, 1234.5):
        assert np.array_equal(ctx.quantize_coefficients(co, step), oracle.quantize_coefficients(co, step)), f"quantize step {step}"
        assert np.array_equal(ctx.dequantize_coefficients(co >> 6, step), oracle.dequantize_coefficients(co >> 6, step)), f"dequantize step {step}"
    f = (rng.standard_normal(n) * 1000).astype(np.float32)
    f[:10] = [0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 0.49999997, 2147483648.0, -2147483904.0, np.float32("nan")]
    assert np.array_equal(ctx.convert_f32_to_i32(f), oracle.convert_f32_to_i32(f)), "float32 -> int32 rounding"


# ---- code-block interface: block-major coefficients + numbps, and the decode-side scatter

def tile_list(oracle, fp):
    n, _ = oracle.fwd_tile_bounds(fp, 0)
    return [oracle.fwd_tile_bounds(fp, i)[1] for i in range(n)]


def check_blocks(ctx, oracle, w, h, c, bits, L, reversible, tile=(0, 0), cb=(64, 64), htj2k=False, seed=11, nframes=2):
    """forward_blocks == oracle planes pushed through the restated getSubbandsForResolution / partitionIntoCodeBlocks /
    T1 shift / codeBlockNumBps; inverse_blocks == oracle inverse of the scattered planes; layout tables identical."""
    rng = np.random.default_rng(seed)
    fp, ip = fwd_inv_params(w, h, c, bits, False, L, reversible, oracle, tile, htj2k)
    frames = np.stack([raw_bytes(synth(rng, h, w, c, bits, False, "smooth" if f else "noise")) for f in range(nframes)])
    got_blocks, got_nb = ctx.forward_blocks(fp, frames, cb[0], cb[1])
    shift6 = reversible and not htj2k
    for f in range(nframes):
        planes = oracle.forward(fp, frames[f])
        want_blocks, want_nb, off = [], [], 0
        for (x0, y0, x1, y1) in tile_list(oracle, fp):
            tw, th = x1 - x0, y1 - y0
            lay_o = oracle.codeblock_layout(tw, th, L, cb[0], cb[1])
            lay_p = ctx.codeblock_layout(tw, th, L, cb[0], cb[1])
            assert len(lay_o) == len(lay_p)
            for a, b in zip(lay_o, lay_p):
                assert all(getattr(a, k) == getattr(b, k) for k, _ in abi.Cblk._fields_), "code-block table"
            for comp in range(c):
                plane = planes[off:off + tw * th].reshape(th, tw)
                bl, nb = oracle.gather_blocks(plane, L, cb[0], cb[1], shift6, htj2k)
                want_blocks.append(bl); want_nb.append(nb)
                off += tw * th
        want_blocks, want_nb = np.concatenate(want_blocks), np.concatenate(want_nb)
        assert np.array_equal(got_blocks[f], want_blocks), f"block-major coefficients, frame {f}"
        assert np.array_equal(got_nb[f], want_nb), f"numbps, frame {f}"
    # decode side: what T1 hands back per block (classic 5/3: one half bit, halved inside when fuse_t1_halve is set)
    if reversible:
        back = (got_blocks >> 6) * 2 if shift6 else got_blocks
        ip.fuse_t1_halve = 1 if shift6 else 0
    else:
        back = np.stack([M.t1_emulate(got_blocks[f], htj2k) for f in range(nframes)])
    px = ctx.inverse_blocks(ip, np.ascontiguousarray(back, dtype=np.int32), cb[0], cb[1])
    for f in range(nframes):
        planes, off = [], 0
        for (x0, y0, x1, y1) in tile_list(oracle, fp):
            tw, th = x1 - x0, y1 - y0
            for comp in range(c):
                planes.append(oracle.scatter_blocks(back[f][off:off + tw * th], tw, th, L, cb[0], cb[1]).reshape(-1))
                off += tw * th
        want_px = oracle.inverse(ip, np.concatenate(planes))
        assert np.array_equal(px[f], want_px), f"inverse from blocks, frame {f}"
        if reversible:
            assert np.array_equal(px[f], frames[f]), "lossless identity through the block interface"


def check_package_api_x1(ctx, oracle, seed=8):
    """The rest of the exported API surface with test-only callers: float64 9/7 wrappers
    (dwt97.go:340-351,410-421), ConvertFloat64ToInt32 (:515-526), LLDimensions (layout.go:5-33), rgb.go wrappers."""
    rng = np.random.default_rng(seed)
    for (w, h, L, x0, y0) in ((64, 48, 3, 0, 0), (33, 17, 2, 1, 0), (7, 1, 2, 0, 0), (130, 70, 5, 0, 0)):
        b = rng.standard_normal((h, w)) * 3000
        assert np.array_equal(ctx.dwt97_forward_f64(b, L, x0, y0).view(np.uint64), oracle.fwd97_f64(b, L, x0, y0).view(np.uint64)), "9/7 f64 forward"
        assert np.array_equal(ctx.dwt97_inverse_f64(b, L, x0, y0).view(np.uint64), oracle.inv97_f64(b, L, x0, y0).view(np.uint64)), "9/7 f64 inverse"
        assert ctx.ll_dimensions(w, h, L, x0, y0) == oracle.ll_dimensions(w, h, L, x0, y0)
    for args in ((0, 5, 3), (5, 5, 0), (1, 1, 4), (1000, 3, 12)):
        assert ctx.ll_dimensions(*args) == oracle.ll_dimensions(*args)
    f = rng.standard_normal(5001) * 1000
    f[:12] = [0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 0.49999999999999994, -0.49999999999999994, 2147483647.4, 2147483647.5, -2147483648.4, np.nan]
    assert np.array_equal(ctx.convert_f64_to_i32(f), oracle.convert_f64_to_i32(f)), "float64 -> int32 (half away from zero)"
    n_w, n_h = 37, 21
    rgb = rng.integers(-300, 300, n_w * n_h * 3).astype(np.int32)
    r, g, b = rgb[0::3], rgb[1::3], rgb[2::3]
    y, cb, cr = ctx.convert_rgb_to_ycbcr(rgb, n_w, n_h)
    oy, ocb, ocr = oracle.ict_forward(r, g, b)
    assert np.array_equal(y, oy) and np.array_equal(cb, ocb) and np.array_equal(cr, ocr), "ConvertRGBToYCbCr"
    back = ctx.convert_ycbcr_to_rgb(y, cb, cr, n_w, n_h)
    orr, og, ob = oracle.ict_inverse(oy, ocb, ocr)
    assert np.array_equal(back[0::3], orr) and np.array_equal(back[1::3], og) and np.array_equal(back[2::3], ob), "ConvertYCbCrToRGB"
    comps = [rng.integers(-9, 9, 1000).astype(np.int32) for _ in range(4)]
    inter = ctx.interleave_components(comps)
    assert np.array_equal(inter, np.stack(comps, axis=1).reshape(-1)), "InterleaveComponents"   # rgb_test.go:259-322
    de = ctx.deinterleave_components(inter, 4)
    assert all(np.array_equal(a, b) for a, b in zip(de, comps)), "DeinterleaveComponents"
    assert ctx.interleave_components([]) is None and ctx.deinterleave_components(np.zeros(0, np.int32), 3) is None


def random_geometry_cases(n, seed, max_w, max_h):
    """Seeded sweep over everything a caller can vary: size, components, bit depth, signedness, levels, wavelet, tiling,
    HTJ2K scaling, fused T1 shift.  Widths are drawn so that aligned (ring-eligible), hybrid and odd geometries all occur."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        kind = rng.integers(0, 3)
        if kind == 0:
            w = int(rng.integers(1, max_w // 16 + 1)) * 16                  # aligned at level 1
        elif kind == 1:
            w = int(rng.integers(1, max_w // 64 + 1)) * 64                  # aligned for several levels
        else:
            w = int(rng.integers(1, max_w + 1))                             # anything
        h = int(rng.integers(1, max_h + 1))
        c = int(rng.choice([1, 1, 1, 3, 3, 2, 4]))
        bits = int(rng.choice([8, 8, 12, 16, 10, 1, 5]))
        signed = bool(rng.integers(0, 2)) and c != 3
        L = int(rng.integers(0, 7))
        rev = bool(rng.integers(0, 2))
        tile = (0, 0)
        if rng.integers(0, 4) == 0:
            tile = (int(rng.choice([16, 32, 48, 64, 33])), int(rng.choice([16, 32, 40, 64, 17])))
        htj2k = bool(rng.integers(0, 5) == 0)
        fuse = bool(rng.integers(0, 4) == 0) and rev
        out.append((w, h, c, bits, signed, L, rev, tile, htj2k, fuse))
    return out


def check_random_case(ctx, oracle, case, seed):
    w, h, c, bits, signed, L, rev, tile, htj2k, fuse = case
    # signed samples narrower than their storage word are not sign-extended by convertPixelData (encoder.go:362-376 tests the
    # raw word), so the reference itself is not an identity there: parity with the oracle is checked, the identity is not
    ident = not (signed and bits not in (8, 16))
    check_pipeline(ctx, oracle, w, h, c, bits, signed, L, rev, tile=tile, htj2k=htj2k, fuse=fuse, kind="noise", seed=seed, identity=ident)


def check_inverse_with_offsets(ctx, oracle, w, h, c, bits, L, reversible, xosiz, yosiz, tile=(0, 0), xtosiz=0, ytosiz=0, seed=21):
    """Decoder-side SIZ geometry with a shifted image / tile grid (tile_assembler.go:33-101, t2/tile_decoder.go:269-294):
    tile windows start at odd coordinates, so the inverse DWT runs with odd origin parity.  Random coefficient planes in,
    pixels and planes compared with the oracle."""
    rng = np.random.default_rng(seed)
    mct = (abi.MCT_RCT if reversible else abi.MCT_ICT) if c == 3 else abi.MCT_NONE
    ds = None
    if not reversible:
        _, ds = steps_for(oracle, L, bits)
    ip = abi.inv_params(w, h, c, bits, False, tile[0], tile[1], L, reversible, False, mct, ds,
                        xosiz=xosiz, yosiz=yosiz, xtosiz=xtosiz, ytosiz=ytosiz)
    n = w * h * c
    co = rng.integers(-(1 << (bits + 2)), 1 << (bits + 2), n).astype(np.int32)
    px, planes = ctx.inverse(ip, co, want_planes=True)
    opx, oplanes = oracle.inverse(ip, co, want_planes=True)
    assert np.array_equal(planes, oplanes), "inverse planes with image / tile offsets"
    assert np.array_equal(px, opx), "inverse pixels with image / tile offsets"


def check_pipelined_order(ctx, oracle, w, h, c, bits, L, reversible, nframes, group_ks, lag, capfd=None):
    """A batch through the persistent launch with the group-pipelined job order forced on (ring_schedule in j2k_b200.cu:
    J2K_RING_GROUP_KS / J2K_RING_LAG are read when the plan is built, so the geometry must be new to the context)."""
    import os
    import re
    old = {k: os.environ.get(k) for k in ("J2K_RING_GROUP_KS", "J2K_RING_LAG", "J2K_B200_TRACE", "J2K_FWD3W")}
    # (the one-producer RGB forward numbers its jobs in triples and keeps the level-major list: component-split jobs here)
    os.environ.update(J2K_RING_GROUP_KS=str(group_ks), J2K_RING_LAG=str(lag), J2K_B200_TRACE="1", J2K_FWD3W="0")
    try:
        rng = np.random.default_rng(w + nframes)
        frames = np.stack([raw_bytes(synth(rng, h, w, c, bits, False, "noise")) for _ in range(nframes)])
        fp, ip = fwd_inv_params(w, h, c, bits, False, L, reversible, oracle)
        co = ctx.forward_batch(fp, frames)
        for f in range(nframes):
            assert np.array_equal(co[f], oracle.forward(fp, frames[f])), f"forward, frame {f}"
        back = co if reversible else np.stack([M.t1_emulate(co[f], False) for f in range(nframes)])
        px = ctx.inverse_batch(ip, back)
        for f in range(nframes):
            assert np.array_equal(px[f], oracle.inverse(ip, back[f])), f"inverse, frame {f}"
        if reversible:
            assert np.array_equal(px, frames), "lossless identity"
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if capfd is not None:
        err = capfd.readouterr().err
        if "[j2k]" in err:  # the trace is latched at first use
            n = [int(m) for m in re.findall(r" ring .* slices=(\d+)", err)]
            assert n and (all(v > 0 for v in n) if nframes > lag else all(v == 0 for v in n)), err


def check_blocks_roi(ctx, oracle, w, h, c, bits, L, reversible, shifts, tile=(0, 0), cb=(32, 32), seed=31, nframes=2):
    """Decode-side MaxShift ROI (decodeCodeBlock, t2/tile_decoder.go:726-730 + applyInverseMaxShift :1113-1138) fused into
    the block scatter: the blocks arrive as T1 leaves them, i.e. with the ROI samples of component k still shifted up by
    shifts[k]; the device undoes it while scattering, before the classic 5/3 "/2".  Checked against the oracle's restated
    rule followed by its scatter + inverse, and (lossless, 0 < shift < 31) against the original frames."""
    rng = np.random.default_rng(seed)
    fp, ip = fwd_inv_params(w, h, c, bits, False, L, reversible, oracle, tile)
    frames = np.stack([raw_bytes(synth(rng, h, w, c, bits, False, "smooth" if f else "noise")) for f in range(nframes)])
    blocks, _ = ctx.forward_blocks(fp, frames, cb[0], cb[1])
    shift6 = bool(reversible)
    if reversible:
        back = ((blocks >> 6) * 2).astype(np.int32)
        ip.fuse_t1_halve = 1
    else:
        back = np.stack([M.t1_emulate(blocks[f], False) for f in range(nframes)]).astype(np.int32)
    # what an encoder with a MaxShift ROI leaves in the code-blocks: samples inside the region carry mag << shift, every
    # background magnitude stays below 2^shift (that is the MaxShift condition, so small shifts are exercised with a
    # region that holds every sample whose magnitude reaches the threshold)
    sent = back.copy()
    lossless_ok = True
    for f in range(nframes):
        off = 0
        for (x0, y0, x1, y1) in tile_list(oracle, fp):
            n = (x1 - x0) * (y1 - y0)
            for comp in range(c):
                sh = int(shifts[comp])
                seg = sent[f][off:off + n]
                if 0 < sh < 31:
                    mag = np.abs(seg.astype(np.int64))
                    region = (rng.random(n) < 0.3) | (mag >= (1 << sh))
                    up = mag << sh
                    ok = up < (1 << 31)
                    region &= ok & (mag > 0)
                    if np.any((mag >= (1 << sh)) & ~region):
                        lossless_ok = False
                    seg[region] = (np.sign(seg[region]) * up[region]).astype(np.int32)
                elif sh >= 31:
                    lossless_ok = False
                off += n
    px = ctx.inverse_blocks(ip, np.ascontiguousarray(sent), cb[0], cb[1], roi_maxshift=shifts)
    for f in range(nframes):
        planes, off = [], 0
        for (x0, y0, x1, y1) in tile_list(oracle, fp):
            tw, th = x1 - x0, y1 - y0
            for comp in range(c):
                blk = oracle.inverse_max_shift(sent[f][off:off + tw * th], int(shifts[comp]))
                planes.append(oracle.scatter_blocks(blk, tw, th, L, cb[0], cb[1]).reshape(-1))
                off += tw * th
        want_px = oracle.inverse(ip, np.concatenate(planes))
        assert np.array_equal(px[f], want_px), f"inverse from ROI-scaled blocks, frame {f}"
        if reversible and lossless_ok:
            assert np.array_equal(px[f], frames[f]), "lossless identity through MaxShift ROI blocks"
    # no shifts at all == the plain call
    assert np.array_equal(ctx.inverse_blocks(ip, np.ascontiguousarray(back), cb[0], cb[1], roi_maxshift=[0] * c),
                          ctx.inverse_blocks(ip, np.ascontiguousarray(back), cb[0], cb[1]))


def check_tall_chunks(ctx, oracle, w, h, c, bits, L, reversible, chunk, seed=41):
    """Job partitioning of the persistent launch with the chunk heights only big batches reach (ring_chunks in j2k_b200.cu
    sizes chunks for a job target, so a single test frame gets 8-row-pair chunks): J2K_RING_TARGET_JOBS=1 asks for one
    chunk per strip, J2K_RING_CHUNK caps it, so the frame is cut into ceil(rows / 2 / chunk) chunks of `chunk` row pairs.
    The knobs are read when the plan is built: the geometry must be new to the context."""
    import os
    old = {k: os.environ.get(k) for k in ("J2K_RING_TARGET_JOBS", "J2K_RING_CHUNK", "J2K_RING_CHUNK_DEEP")}
    os.environ.update(J2K_RING_TARGET_JOBS="1", J2K_RING_CHUNK=str(chunk))
    os.environ.pop("J2K_RING_CHUNK_DEEP", None)
    try:
        check_pipeline(ctx, oracle, w, h, c, bits, False, L, reversible, kind="noise", seed=seed)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def check_blocks_roi_general(ctx, oracle, w, h, c, bits, L, reversible, tile=(0, 0), cb=(32, 32), seed=37, nframes=2, masked=True, maxshift=None):
    """The whole ROI tail of decodeCodeBlock on the device (t2/tile_decoder.go:723-742): MaxShift (Srgn = 0), the classic 5/3
    "/2", then general scaling (Srgn = 1): blocks the region touches are divided by 2^shift - every sample (rectangle form,
    applyInverseGeneralScaling :1082-1090) or the samples of their mask (applyInverseGeneralScalingMasked :1093-1111), with
    Go's truncating division.  Checked against the oracle's restatements applied per block in the reference's order, followed
    by its scatter and inverse."""
    rng = np.random.default_rng(seed)
    fp, ip = fwd_inv_params(w, h, c, bits, False, L, reversible, oracle, tile)
    frames = np.stack([raw_bytes(synth(rng, h, w, c, bits, False, "smooth" if f else "noise")) for f in range(nframes)])
    blocks, _ = ctx.forward_blocks(fp, frames, cb[0], cb[1])
    if reversible:
        sent = ((blocks >> 6) * 2).astype(np.int32)      # T1 output of the classic lossless path: one half bit
        ip.fuse_t1_halve = 1
    else:
        sent = np.stack([M.t1_emulate(blocks[f], False) for f in range(nframes)]).astype(np.int32)
    # per tile-component block tables, in the order the library numbers the blocks
    tiles = tile_list(oracle, fp)
    tabs, nblk = [], 0
    for (x0, y0, x1, y1) in tiles:
        t = oracle.codeblock_layout(x1 - x0, y1 - y0, L, cb[0], cb[1])
        tabs.append(t)
        nblk += len(t) * c
    shifts = np.zeros((nframes, nblk), np.int32)
    mask = np.zeros(sent.shape, np.uint8) if masked else None
    want_planes = []
    for f in range(nframes):
        planes, off, bi = [], 0, 0
        for (x0, y0, x1, y1), t in zip(tiles, tabs):
            tw, th = x1 - x0, y1 - y0
            for comp in range(c):
                seg = sent[f][off:off + tw * th].copy()
                if maxshift is not None:
                    seg = oracle.inverse_max_shift(seg, int(maxshift[comp]))
                if reversible:
                    seg = np.where(seg < 0, -((-seg.astype(np.int64)) // 2), seg // 2).astype(np.int32)  # Go "/ 2" (t2/tile_decoder.go:989-993)
                for b in t:
                    n = int(b.width) * int(b.height)
                    o = int(b.offset)
                    sh = int(rng.integers(0, 9)) if rng.random() < 0.6 else 0
                    shifts[f, bi] = sh
                    mk = None
                    if masked:
                        mk = (rng.random(n) < 0.5).astype(np.uint8)
                        mask[f, off + o:off + o + n] = mk
                    if sh > 0:
                        seg[o:o + n] = oracle.inverse_general_scaling(seg[o:o + n], sh, mk)
                    bi += 1
                planes.append(oracle.scatter_blocks(seg, tw, th, L, cb[0], cb[1]).reshape(-1))
                off += tw * th
        want_planes.append(np.concatenate(planes))
    px = ctx.inverse_blocks(ip, np.ascontiguousarray(sent), cb[0], cb[1], roi_maxshift=maxshift, block_scale_shift=shifts, sample_mask=mask)
    ip2 = fwd_inv_params(w, h, c, bits, False, L, reversible, oracle, tile)[1]   # the oracle side already halved: no fused "/2"
    for f in range(nframes):
        want_px = oracle.inverse(ip2, want_planes[f])
        assert np.array_equal(px[f], want_px), f"inverse from general-scaling ROI blocks, frame {f}"


def check_one_producer_forward(ctx, oracle, w, h, bits, L, nframes, tile=(0, 0), chunk=0, capfd=None, arm_on="2"):
    """ICT + 9/7 forward of a batch of raw RGB frames through the one-producer level-1 kernel (fwd3w_kernel, j2k_ring.cuh)
    and, for the same frames, through the component-split jobs it replaces (J2K_FWD3W=0): both bit-identical to the oracle
    (encoder.go:277-288 + dwt97.go:47-190 + encoder.go:2311-2329).  Environment knobs are read when a plan is built, so each
    arm uses a geometry / batch size new to the context (a frame more for the second arm)."""
    import os
    import re
    old = {k: os.environ.get(k) for k in ("J2K_FWD3W", "J2K_RING_CHUNK", "J2K_B200_TRACE")}
    try:
        os.environ["J2K_B200_TRACE"] = "1"
        if chunk:
            os.environ["J2K_RING_CHUNK"] = str(chunk)
        rng = np.random.default_rng(w * 7 + h)
        frames = np.stack([raw_bytes(synth(rng, h, w, 3, bits, False, "noise" if f & 1 else "smooth")) for f in range(nframes + 1)])
        fp, _ = fwd_inv_params(w, h, 3, bits, False, L, False, oracle, tile)
        want = [oracle.forward(fp, frames[f]) for f in range(nframes + 1)]
        for arm, n in ((arm_on, nframes), ("0", nframes + 1)):  # 2: also for launches below the size the plan builder asks for (1)
            os.environ["J2K_FWD3W"] = arm
            co = ctx.forward_batch(fp, frames[:n])
            for f in range(n):
                nd = int(np.count_nonzero(co[f] != want[f]))
                assert nd == 0, f"J2K_FWD3W={arm}, frame {f}: {nd} differing coefficients"
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    if capfd is not None:
        err = capfd.readouterr().err
        if "[j2k]" in err:  # the trace is latched at first use
            x3 = [int(m) for m in re.findall(r"fwd ring .* x3=(\d)", err)]
            assert x3 and x3[0] == 1 and x3[-1] == 0, err


def check_failed_device(oracle, monkeypatch, lib_path=None, devices=(0, 0)):
    import j2kb200
    import pytest
    rng = np.random.default_rng(77)
    w, h, n = 96, 64, 5
    frames = np.stack([raw_bytes(synth(rng, h, w, 1, 12, False, "smooth")) for _ in range(n)])
    fp, ip = fwd_inv_params(w, h, 1, 12, False, 3, True, oracle)
    want = np.stack([oracle.forward(fp, frames[f]) for f in range(n)])
    with j2kb200.Context(devices=list(devices), lib_path=lib_path) as c2:
        assert c2.device_count == 2 and not c2.device_failed(0) and not c2.device_failed(1)
        assert np.array_equal(c2.forward_batch(fp, frames), want)        # both slots healthy
        monkeypatch.setenv("J2K_FAULT_DEVICE", "1")
        got = c2.forward_batch(fp, frames)                                # slot 1 fails: its block is re-run on slot 0
        monkeypatch.delenv("J2K_FAULT_DEVICE")
        assert np.array_equal(got, want)
        assert c2.device_failed(1) and not c2.device_failed(0)
        assert "removed from the round-robin" in c2.lib.j2k_last_error(c2.h).decode()
        assert np.array_equal(c2.inverse_batch(ip, want), frames.reshape(n, -1))   # later calls: slot 0 only
        assert np.array_equal(c2.forward_batch(fp, frames), want)
        monkeypatch.setenv("J2K_FAULT_DEVICE", "0")
        with pytest.raises(j2kb200.J2KError) as e:                        # the last device goes: an error, not a hang
            c2.forward_batch(fp, frames)
        monkeypatch.delenv("J2K_FAULT_DEVICE")
        assert e.value.code == abi.J2K_ERR_CUDA and c2.device_failed(0)
        with pytest.raises(j2kb200.J2KError) as e:
            c2.forward_batch(fp, frames)
        assert "no usable device" in str(e.value)
