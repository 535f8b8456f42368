"""The quantizer's division: q = rint((c / float32(step)) * scale)  (encoder.go:2311-2329).

The kernels never divide.  They compute the IEEE quotient c / step' (step' = step / scale, a power-of-two rescale, exact) from
the correctly rounded reciprocal r = RN(1 / step') with two fused multiply-adds (Markstein):

    q0 = RN(c * r);  e = RN(fma(-q0, step', c)) = c - q0 * step'  exactly;  q = RN(fma(e, r, q0))

(`FwdRing::run` packed form, `FwdFast::quant_vec`, `div_by_step` in go-dicom-codec_b200/csrc).  This file is the proof the
kernel comments cite:

  * CPU (`-m "not gpu"`): the three-operation sequence, emulated with exact rational arithmetic and one correct float32
    rounding per operation, equals RN(c / step') for every one of the 2048 step mantissas the reference can produce
    (quantization.go:130-154: 11 fraction bits) at several exponents, both scales, on random, tie, boundary and extreme c;
    the one divisor class the published theorem excludes (all-ones significand) and steps with extreme exponents are routed to
    the true division by the host (`markstein_safe` in j2k_b200.cu) - checked through the emulator build of the product sources;
  * GPU (`-m gpu`): every mantissa x {classic 64, HTJ2K 1} x a spread of exponents through the real ring kernel (aligned
    frame) and the per-level kernels (odd frame), against the oracle's true division, bit for bit.
"""
from fractions import Fraction

import numpy as np
import pytest

import parity_cases as PC
from j2kb200 import abi


# ---------------------------------------------------------------- exact float32 model (CPU)

def rn32(x: Fraction) -> Fraction:
    """Round a rational to the nearest float32 (ties to even, gradual underflow as the FMA unit does); returns the float32
    value as a Fraction."""
    if x == 0:
        return Fraction(0)
    s = -1 if x < 0 else 1
    a = abs(x)
    e = a.numerator.bit_length() - a.denominator.bit_length()  # floor(log2 a) or one more
    if Fraction(2) ** e > a:
        e -= 1
    assert e <= 127, "overflow"
    ulp = Fraction(2) ** (max(e, -126) - 23)
    q = a / ulp                     # in [2^23, 2^24) for normal results
    n = q.numerator // q.denominator
    rem = q - n
    if rem > Fraction(1, 2) or (rem == Fraction(1, 2) and (n & 1)):
        n += 1
    return s * n * ulp


def f32(x) -> Fraction:
    return Fraction(float(np.float32(x)))


def markstein(c: Fraction, step: Fraction) -> Fraction:
    r = rn32(1 / step)
    q0 = rn32(c * r)
    e = rn32(c - q0 * step)         # fma(-q0, step, c): one rounding of the exact value
    return rn32(q0 + e * r)         # fma(e, r, q0)


def interesting_c(rng, step: Fraction, n_random=12):
    """Coefficients that stress the quotient: random, exact multiples, half-way points of rint, neighbours of those in
    float32, tiny and large magnitudes."""
    out = []
    for k in (0, 1, 2, 3, 7, 100, 4095, 65535):
        for half in (0, Fraction(1, 2)):
            v = f32(float((k + half) * step))
            out += [v, -v]
            fv = np.float32(float(v))
            for nb in (np.nextafter(fv, np.float32(np.inf)), np.nextafter(fv, np.float32(-np.inf))):
                out.append(Fraction(float(nb)))
    for _ in range(n_random):
        out.append(f32(rng.normal(0, 300.0)))
        out.append(f32(rng.uniform(-1, 1) * 2.0 ** rng.integers(-20, 24)))
    lo, hi = Fraction(2) ** -40, Fraction(2) ** 30   # the coefficient domain stated at markstein_safe (j2k_b200.cu)
    return [c for c in out if lo <= abs(c) <= hi]


def test_markstein_sequence_is_the_ieee_quotient_for_every_reference_step():
    """All 2048 mantissas; the sequence is invariant under power-of-two scaling of step' (away from under/overflow), so the two
    scales and the exponents only move the operands around inside the stated domain."""
    rng = np.random.default_rng(20)
    bad = n = 0
    for mant in range(2048):
        scale, ex = ((64, -9), (1, 0), (64, 3), (1, 11))[mant % 4]
        step = f32((1.0 + mant / 2048.0) * 2.0 ** ex)   # float32(stepSize): exact, 11 fraction bits
        step_eff = step / scale                          # power-of-two rescale: exact
        for c in interesting_c(rng, step_eff, n_random=4):
            want = rn32(c / step_eff)
            # the reference's own expression: RN(c / step) * scale, the multiply is exact
            assert rn32(c / step) * scale == want
            bad += markstein(c, step_eff) != want
            n += 1
    assert bad == 0 and n > 100000


def test_all_ones_significand_divisor():
    """The published correctness theorem for this sequence (Markstein 1990; Cornea, Harrison, Tang 2002) excludes divisors
    whose significand is all ones.  With r the CORRECTLY rounded reciprocal (the host computes it with an IEEE division) a
    sample of 20 000 dividends shows no miss for that divisor either; `markstein_safe` nevertheless sends the class to the
    true division, so the product claims nothing the theorem does not cover (test_unsafe_steps_take_the_true_division)."""
    step = Fraction(float(np.nextafter(np.float32(2.0), np.float32(0.0))))  # 0x3FFFFFFF
    rng = np.random.default_rng(5)
    for _ in range(20000):
        c = f32(rng.uniform(1, 2) * 2.0 ** rng.integers(0, 12))
        assert markstein(c, step) == rn32(c / step)


def _steps_for(mants, ex, L):
    """3L+1 runtime steps (double) built from 4 mantissas: band k uses mants[k % 4]."""
    return [float(np.float32((1.0 + mants[k % 4] / 2048.0) * 2.0 ** ex)) for k in range(3 * L + 1)]


def _check_quant(ctx, oracle, w, h, steps, htj2k, seed, bits=12, c=1):
    L = (len(steps) - 1) // 3
    rng = np.random.default_rng(seed)
    img = PC.synth(rng, h, w, c, bits, False, "noise")
    fp = abi.fwd_params(w, h, c, bits, False, num_levels=L, reversible=False, htj2k=htj2k,
                        mct_mode=abi.MCT_ICT if c == 3 else abi.MCT_NONE, steps=steps)
    got = ctx.forward(fp, PC.raw_bytes(img))
    want = oracle.forward(fp, PC.raw_bytes(img))
    assert np.abs(want.astype(np.int64)).max() < 2 ** 31 - 1, "test case leaves the int32 domain (Go wraps, F2I saturates)"
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, (w, h, c, htj2k, steps, bad[:8], got[bad[:8]], want[bad[:8]])


def test_unsafe_steps_take_the_true_division(oracle):
    """Through the emulator build of the product sources: an all-ones significand, a step with 23 fraction bits and extreme
    exponents are accepted and quantized exactly like the reference's division."""
    import emu_lib
    import j2kb200
    ones = float(np.nextafter(np.float32(2.0), np.float32(0.0)))
    with j2kb200.Context(lib_path=emu_lib.build()) as ectx:
        for steps in ([ones] * 4, [ones / 64, 0.7853981852531433, 3.0000002384185791, ones * 4],
                      [2.0 ** -9, 2.0 ** 45, 1.0, ones]):  # (quotients stay inside int32: Go's overflow wrap is out of the domain)
            for htj2k in (False, True):
                _check_quant(ectx, oracle, 64, 16, steps, htj2k, seed=3)     # ring geometry
                _check_quant(ectx, oracle, 37, 11, steps, htj2k, seed=4)     # per-level kernels


@pytest.mark.gpu
@pytest.mark.parametrize("htj2k", [False, True])
def test_every_step_mantissa_through_the_kernels(ctx, oracle, htj2k):
    """All 2048 mantissas (four per call, one per band), exponents from far below to far above the coefficient range, ring
    kernel (128 x 32, L = 1: LL is quantized too) and per-level kernels (odd 61 x 23), vs the oracle's division."""
    seed = 100
    for ex in (-12, -6, -2, 0, 1, 4, 9):
        for m0 in range(0, 2048, 4):
            if ex not in (-2, 1) and m0 % 64:
                continue  # the full mantissa sweep at two exponents, every 16th group elsewhere
            steps = _steps_for((m0, m0 + 1, m0 + 2, m0 + 3), ex, 1)
            _check_quant(ctx, oracle, 128, 32, steps, htj2k, seed)
            if m0 % 16 == 0:
                _check_quant(ctx, oracle, 61, 23, steps, htj2k, seed + 1)
            seed += 2


@pytest.mark.gpu
def test_unsafe_steps_on_the_gpu(ctx, oracle):
    ones = float(np.nextafter(np.float32(2.0), np.float32(0.0)))
    for steps in ([ones] * 7, [ones / 64, 0.7853981852531433, 3.0000002384185791, ones * 4, 1e-2, 17.25, 0.1],
                  [2.0 ** -7, 2.0 ** 45, 1.0, ones, 5.5, 2e-2, 3e7]):
        for htj2k in (False, True):
            _check_quant(ctx, oracle, 256, 64, steps, htj2k, seed=7)
            _check_quant(ctx, oracle, 128, 48, steps, htj2k, seed=8, bits=8, c=3)
            _check_quant(ctx, oracle, 77, 35, steps, htj2k, seed=9)
