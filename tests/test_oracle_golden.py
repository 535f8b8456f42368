"""Oracle vs golden vectors produced by OpenJPEG 2.5.4 (tests/golden/make_golden.py).

OpenJPEG is the library whose arithmetic the reference's JPEG 2000 path clones
(jpeg2000/encoder.go:1790); its decoded pixels pin the oracle's
9/7 + quantization + dequantization chain end to end and the 5/3 low-pass bands.
"""
import json
import os
import sys

import numpy as np
import pytest

from j2kb200 import abi

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
import np_mirror as M  # noqa: E402

G = np.load(os.path.join(HERE, "golden", "openjpeg_254.npz"))
META = json.load(open(os.path.join(HERE, "golden", "openjpeg_254.json")))["cases"]


def _cases(kind):
    return [k for k, v in META.items() if v["kind"] == kind]


def _raw(a, prec):
    return np.ascontiguousarray(a.astype(np.uint8 if prec <= 8 else "<u2")).view(np.uint8).reshape(-1)


def _steps(oracle, name, L, prec):
    enc, _ = oracle.openjpeg_quant_params(L, prec)
    qcd = G[name + "_qcd"].tobytes()
    assert qcd[0] == 0x42  # scalar expounded, 2 guard bits (jpeg2000/quantization.go:219-220)
    assert qcd[1:] == b"".join(bytes([int(e) >> 8, int(e) & 0xFF]) for e in enc), "QCD bytes differ from OpenJPEG"
    return oracle.runtime_quant_steps(enc, L, prec), oracle.decode_quant_steps(enc, L, prec, False)


@pytest.mark.parametrize("name", _cases("mono97"))
def test_mono_97_chain_equals_openjpeg(oracle, name):
    c = META[name]
    a, want = G[name + "_in"], G[name + "_dec"]
    enc_steps, dec_steps = _steps(oracle, name, c["levels"], c["prec"])
    fp = abi.fwd_params(c["w"], c["h"], 1, c["prec"], False, num_levels=c["levels"], reversible=False, steps=enc_steps)
    q = oracle.forward(fp, _raw(a, c["prec"]))
    v = M.t1_emulate(q)
    ip = abi.inv_params(c["w"], c["h"], 1, c["prec"], False, num_levels=c["levels"], reversible=False, steps=dec_steps)
    px = oracle.inverse(ip, v)
    got = px.view(np.uint8 if c["prec"] <= 8 else "<u2").reshape(c["h"], c["w"])
    assert np.count_nonzero(got != want) == 0


@pytest.mark.parametrize("name", _cases("rgb97"))
def test_rgb_ict_97_chain_equals_openjpeg(oracle, name):
    c = META[name]
    a, want = G[name + "_in"], G[name + "_dec"]
    L, w, h = c["levels"], c["w"], c["h"]
    enc_steps, dec_steps = _steps(oracle, name, L, 8)
    fp = abi.fwd_params(w, h, 3, 8, False, num_levels=L, reversible=False, steps=enc_steps, mct_mode=abi.MCT_ICT)
    q = oracle.forward(fp, _raw(a, 8)).reshape(3, h, w)
    rects = oracle.band_rects(w, h, 0, 0, L)
    planes = []
    for comp in range(3):
        f = M.t1_emulate(q[comp]).astype(np.float32)
        for (ox, oy, bw, bh), st in zip(rects, dec_steps):  # t2/tile_decoder.go:970-987
            f[oy:oy + bh, ox:ox + bw] *= np.float32(0.5 * st)
        planes.append(oracle.inv97(f, L))
    y, u, v = planes
    F = np.float32
    # OpenJPEG's own tail (opj_mct_decode_real on un-rounded float32 samples); the Go tail differs on purpose
    #  and is covered by test_oracle_kat / the GPU parity tests, not by this golden.
    r = y + v * F(1.402)
    g = y - u * F(0.34413) - v * F(0.71414)
    b = y + u * F(1.772)
    got = np.stack([np.clip(np.rint(p.astype(np.float64)) + 128, 0, 255) for p in (r, g, b)], -1).astype(np.uint8)
    assert np.count_nonzero(got != want) == 0


@pytest.mark.parametrize("name", _cases("mono53"))
def test_mono_53_ll_bands_equal_openjpeg(oracle, name):
    c = META[name]
    a = G[name + "_in"].astype(np.int32)
    dc = 1 << (c["prec"] - 1)
    for n in c["ll_levels"]:
        want = G[f"{name}_ll{n}"].astype(np.int32)
        co = oracle.fwd53(a - dc, n)
        lw, lh = oracle.ll_dimensions(c["w"], c["h"], n)
        assert want.shape == (lh, lw)
        got = np.clip(co[:lh, :lw] + dc, 0, 2 * dc - 1)
        assert np.array_equal(got, want), (name, n)
    # and the full forward path is the identity through the inverse path
    fp = abi.fwd_params(c["w"], c["h"], 1, c["prec"], False, num_levels=c["levels"], reversible=True)
    ip = abi.inv_params(c["w"], c["h"], 1, c["prec"], False, num_levels=c["levels"], reversible=True)
    raw = _raw(G[name + "_in"], c["prec"])
    assert np.array_equal(oracle.inverse(ip, oracle.forward(fp, raw)), raw)


def test_rgb_rct_53_identity(oracle):
    a = G["c53_88x72_L3_in"]
    fp = abi.fwd_params(88, 72, 3, 8, False, num_levels=3, reversible=True, mct_mode=abi.MCT_RCT)
    ip = abi.inv_params(88, 72, 3, 8, False, num_levels=3, reversible=True, mct_mode=abi.MCT_RCT)
    raw = _raw(a, 8)
    assert np.array_equal(oracle.inverse(ip, oracle.forward(fp, raw)), raw)


def _interop():
    man = json.load(open(os.path.join(HERE, "golden", "interop", "manifest.json")))
    return man["fixtures"]


@pytest.mark.parametrize("fx", _interop(), ids=lambda f: f["name"])
def test_interop_raws_lossless_identity(oracle, fx):
    # jpeg2000/htj2k/interop_manifest_test.go:43-74 demands decode == input.raw; the sample-domain
    # half of that contract is forward(5/3 [+RCT]) -> inverse == identity, for classic and HTJ2K flags.
    raw = np.fromfile(os.path.join(HERE, "golden", "interop", fx["name"] + ".raw"), np.uint8)
    C = fx["components"]
    for htj2k in (False, True):
        bd = fx["bitsAllocated"] if htj2k else fx["bitsStored"]  # htj2k/codec.go:150 vs lossless/codec.go:157
        mct = abi.MCT_RCT if C == 3 else abi.MCT_NONE
        fp = abi.fwd_params(fx["width"], fx["height"], C, bd, fx["signed"], num_levels=5, reversible=True,
                            htj2k=htj2k, mct_mode=mct, fuse_t1_shift=True)
        co = oracle.forward(fp, raw)
        if not htj2k:
            assert np.all((co & 63) == 0)
            co = (co >> 6) * 2  # what T1 hands back: the value with one half-bit (t1/decoder.go:630-647)
        ip = abi.inv_params(fx["width"], fx["height"], C, bd, fx["signed"], num_levels=5, reversible=True,
                            htj2k=htj2k, mct_mode=mct, fuse_t1_halve=True)
        assert np.array_equal(oracle.inverse(ip, co), raw)


def test_ct1_lossless_and_lossy_paths(oracle):
    # C1: CT1_J2KI, 512x512 signed 16-bit
    ob = np.load(os.path.join(HERE, "golden", "ct1_j2ki.npz"))["offset_binary"]
    s16 = (ob.astype(np.int32) - 32768).astype("<i2")
    raw = s16.view(np.uint8).reshape(-1)
    fp = abi.fwd_params(512, 512, 1, 16, True, num_levels=5, reversible=True)
    ip = abi.inv_params(512, 512, 1, 16, True, num_levels=5, reversible=True)
    assert np.array_equal(oracle.inverse(ip, oracle.forward(fp, raw)), raw)
    enc, _ = oracle.openjpeg_quant_params(5, 16)
    fp = abi.fwd_params(512, 512, 1, 16, True, num_levels=5, reversible=False, steps=oracle.runtime_quant_steps(enc, 5, 16))
    ip = abi.inv_params(512, 512, 1, 16, True, num_levels=5, reversible=False, steps=oracle.decode_quant_steps(enc, 5, 16))
    back = oracle.inverse(ip, M.t1_emulate(oracle.forward(fp, raw))).view("<i2").astype(np.int32)
    assert np.abs(back - s16.reshape(-1)).max() <= 1  # all passes kept: reconstruction within 1 LSB
