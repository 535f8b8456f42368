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
                   seed=1, mct=None, steps_kind="openjpeg", want_planes=True):
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
    else:
        px = ctx.inverse(ip, back_in)
        opx = oracle.inverse(ip, back_in)
    nd = int(np.count_nonzero(px != opx))
    assert nd == 0, f"inverse: {nd} differing bytes"
    if reversible:
        assert np.array_equal(px, raw), "lossless identity"
    return got, px
