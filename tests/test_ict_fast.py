"""The float32 fast path of the inverse ICT in the level-1 inverse ring kernel (`InvRing::ict_fast_row`, j2k_ring.cuh) against the
reference's float64 arithmetic (colorspace/ict.go:16-21: R = Round(y + 1.402 cr), G = Round(y - 0.34413 cb - 0.71414 cr),
B = Round(y + 1.772 cb), math.Round = half away from zero).

CPU (`-m "not gpu"`): a numpy model of the kernel's operation sequence - every fma evaluated exactly in float64 (the operands
are short enough, see `fma32`) and rounded once to float32 - is compared with the Go formula
  * for R and B over EVERY (y, chroma) pair of the fast path's domain |y|, |cb|, |cr| < 512 (no guard: the result must always be
    right, including the exact ties cr = +-250, cb = +-125, +-375);
  * for G over structured and random triples: wherever the distance-to-tie guard passes the result must be right, and the guard
    must pass for all but a tiny fraction of samples.
GPU (`-m gpu`): crafted coefficient planes whose level-1 inverse lands on ties and near-ties go through the real kernel and are
compared with the oracle (float64 path), together with the noise-image cases of test_gpu_parity.py.
"""
import numpy as np
import pytest

import parity_cases as PC
from j2kb200 import abi

MG = np.float32(12582912.0)
BIAS = np.float32(2.0 ** -20)
LIMIT = 512


def fma32(a, b, c):
    """float32 fma of arrays: a * b + c is exact in float64 for the operand sizes of this path (a 24-bit constant or a float32
    value, a 10-bit integer-valued sample, an addend whose bits overlap the product's within 53 bits), then ONE rounding."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)


def rint_magic(t):
    b = (t + MG).astype(np.float32)
    return (b - MG).astype(np.float32), b


def go_round(v):
    """math.Round: half away from zero."""
    return np.where(v >= 0, np.floor(v + 0.5), np.ceil(v - 0.5)).astype(np.int64)


def go_ict_inverse(y, cb, cr):
    y, cb, cr = (v.astype(np.float64) for v in (y, cb, cr))
    r = go_round(y + 1.402 * cr)
    g = go_round(y - 0.34413 * cb - 0.71414 * cr)
    b = go_round(y + 1.772 * cb)
    return r, g, b


def fast_rb(y, c, hi, lo):
    hi, lo = np.float32(hi), np.float32(lo)
    t = fma32(np.full_like(y, lo), c, fma32(np.full_like(y, hi), c, y))
    tb = fma32(t, np.full_like(t, BIAS), t)
    r, _ = rint_magic(tb)
    return r.astype(np.int64)


def fast_g(y, cb, cr):
    t = fma32(np.full_like(y, np.float32(-0.71414)), cr, fma32(np.full_like(y, np.float32(-0.34413)), cb, y))
    r, _ = rint_magic(t)
    dist = np.abs((t.astype(np.float64) - r.astype(np.float64)))
    mag = np.maximum(np.abs(y), np.maximum(np.abs(cb), np.abs(cr)))
    ok = dist < (np.float32(0.5) - mag * np.float32(2.0 ** -21))
    return r.astype(np.int64), ok


def test_r_and_b_are_exact_over_the_whole_domain():
    v = np.arange(-(LIMIT - 1), LIMIT, dtype=np.float32)
    y, c = np.meshgrid(v, v, indexing="ij")
    y, c = y.ravel(), c.ravel()
    zero = np.zeros_like(y)
    want_r, _, _ = go_ict_inverse(y, zero, c)
    _, _, want_b = go_ict_inverse(y, c, zero)
    assert np.array_equal(fast_rb(y, c, 1.375, 0.027), want_r)
    assert np.array_equal(fast_rb(y, c, 1.75, 0.022), want_b)
    # the ties really occur and really are rounded away from zero by the reference
    ties = (c == 125) & (y == 0)
    assert want_b[ties][0] == 222 and go_ict_inverse(zero[:1], -c[ties][:1], zero[:1])[2][0] == -222


def test_g_is_exact_wherever_the_guard_passes_and_the_guard_rarely_fails():
    rng = np.random.default_rng(11)
    n = 4_000_000
    for scale in (40, 140, 500):
        y = np.clip(np.rint(rng.normal(0, scale, n)), -(LIMIT - 1), LIMIT - 1).astype(np.float32)
        cb = np.clip(np.rint(rng.normal(0, scale, n)), -(LIMIT - 1), LIMIT - 1).astype(np.float32)
        cr = np.clip(np.rint(rng.normal(0, scale, n)), -(LIMIT - 1), LIMIT - 1).astype(np.float32)
        got, ok = fast_g(y, cb, cr)
        _, want, _ = go_ict_inverse(y, cb, cr)
        assert np.array_equal(got[ok], want[ok])
        assert (~ok).mean() < 2e-3
    # structured: every (cb, cr) pair at a few y, among them the exact decimal ties of G (34413 cb + 71414 cr = 50000 mod 1e5)
    v = np.arange(-(LIMIT - 1), LIMIT, dtype=np.float32)
    cb, cr = (a.ravel() for a in np.meshgrid(v, v, indexing="ij"))
    ties = 0
    for yy in (-300.0, -1.0, 0.0, 77.0, 511.0):
        y = np.full_like(cb, np.float32(yy))
        got, ok = fast_g(y, cb, cr)
        _, want, _ = go_ict_inverse(y, cb, cr)
        assert np.array_equal(got[ok], want[ok])
        tie = ((34413 * cb.astype(np.int64) + 71414 * cr.astype(np.int64)) % 100000) == 50000
        assert not ok[tie].any()        # a real tie is always left to the float64 path
        ties += int(tie.sum())
    assert ties > 0


def _tie_planes(w, h, rng):
    """Level-shifted Y, Cb, Cr planes (integers) full of decimal ties and near-ties of all three outputs."""
    y = rng.integers(-128, 128, (h, w))
    cb = rng.choice(np.array([125, -125, 375, -375, 124, 126, 0, 250, -250, 57, -91]), (h, w))
    cr = rng.choice(np.array([250, -250, 249, 251, 0, 125, -125, 33, -78, 500, -500]), (h, w))
    return y, cb, cr


def _check_ties(ctx, oracle, w, h, seed):
    """A one-level 9/7 inverse whose output IS a chosen integer image: the oracle's forward transform of the chosen Y / Cb / Cr
    planes, quantized to a fine step (2^-6, HTJ2K scaling so that dequantization is float32(q) * step), reproduces them to
    within float noise far below 0.5 - after the half-even rounding the inverse ICT sees exactly the crafted ties."""
    rng = np.random.default_rng(seed)
    planes = _tie_planes(w, h, rng)
    step = 2.0 ** -6
    # the reference's 9/7 pair has gain 2 on each high-pass direction (dwt97.go:19-22: encode K, decode 2/K), normally absorbed by
    # the encode-side sub-band gains: here the decode steps of HL / LH / HH carry the 1/2, 1/2, 1/4
    steps = [step, step / 2, step / 2, step / 4]
    q = [np.rint(oracle.fwd97(p.astype(np.float32), 1) / np.float32(step)).astype(np.int32) for p in planes]
    ip = abi.inv_params(w, h, 3, 8, False, num_levels=1, reversible=False, htj2k=True, mct_mode=abi.MCT_ICT, steps=steps)
    co = np.concatenate([v.reshape(-1) for v in q])
    got = ctx.inverse(ip, co)
    want = oracle.inverse(ip, co)
    assert np.array_equal(got, want)
    # the crafted values survive the transform pair (so the ties are really exercised)
    lw, lh = (w + 1) // 2, (h + 1) // 2
    for v, p in zip(q, planes):
        f = v.astype(np.float32).reshape(h, w) * np.float32(step)
        f[:lh, lw:] *= np.float32(0.5); f[lh:, :lw] *= np.float32(0.5); f[lh:, lw:] *= np.float32(0.25)
        back = np.rint(oracle.inv97(f, 1)).astype(np.int64)
        assert (back == p).mean() > 0.99


def test_tie_planes_through_the_emulator(oracle):
    import emu_lib
    import j2kb200
    with j2kb200.Context(lib_path=emu_lib.build()) as ectx:
        _check_ties(ectx, oracle, 64, 12, seed=1)


@pytest.mark.gpu
def test_tie_planes_on_the_gpu(ctx, oracle):
    for i, (w, h) in enumerate(((256, 64), (512, 96), (1024, 40))):
        _check_ties(ctx, oracle, w, h, seed=10 + i)
