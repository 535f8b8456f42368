"""The reference's own known-answer tests for the path, replayed against the oracle.

Each test names the reference test it restates (paths relative to the reference root).
"""
import numpy as np
import pytest

from j2kb200 import abi


def test_qcd_bytes_l5_8bit(oracle):
    # jpeg2000/quantization_test.go:68-86 and jpeg2000/openjpeg_lossless_flow_test.go:70-92
    enc, _ = oracle.openjpeg_quant_params(5, 8)
    got = b"".join(bytes([int(e) >> 8, int(e) & 0xFF]) for e in enc)
    assert got.hex() == "772076f076f076c06f006f006ee067506750676850055005504757d357d35762"


def test_runtime_steps_use_encoded_values(oracle):
    # jpeg2000/quantization_test.go:88-104
    enc, steps = oracle.openjpeg_quant_params(5, 8)
    expn, mant = int(enc[0]) >> 11, int(enc[0]) & 0x7FF
    want = np.ldexp(1.0 + mant / 2048.0, 8 - expn)
    assert steps[0] != want
    assert oracle.runtime_quant_steps(enc, 5, 8)[0] == want


def test_quality_monotonic_ll_step(oracle):
    # jpeg2000/quantization_test.go:106-123
    ll = [oracle.quality_quant_params(q, 5, 16)[1][0] for q in (1, 20, 50, 80, 90, 95, 99)]
    assert all(b < a for a, b in zip(ll, ll[1:]))


def test_encode_decode_step_roundtrip(oracle):
    # jpeg2000/quantization_test.go:52-66 (5 % tolerance)
    enc, steps = oracle.quality_quant_params(80, 5, 16)
    dec = oracle.decode_quant_steps(enc, 5, 16, reversible=False)
    assert np.all(np.abs(dec - steps) <= 0.05 * np.abs(steps))


def test_quantizer_rounds_after_t1_scaling(oracle):
    # jpeg2000/openjpeg_lossless_flow_test.go:94-105: (0.49 / 1.0) * 64 -> 31
    p = abi.fwd_params(1, 1, 1, 8, False, num_levels=1, reversible=False, steps=[1.0, 1.0, 1.0, 1.0])
    # a 1x1 image is never transformed; its single sample is quantized with the LL step
    out = oracle.forward(p, np.array([128], np.uint8))
    assert out[0] == 0
    # the arithmetic itself, through the exported float path
    q = np.float32(np.float32(0.49) / np.float32(1.0)) * np.float32(64)
    assert int(np.rint(np.float64(q))) == 31


def test_float64_to_int32_kat(oracle):
    # jpeg2000/wavelet/dwt97_test.go:424-441
    got = oracle.convert_f64_to_i32([-100.7, -1.4, 0.0, 1.5, 100.3, 1000.8])
    assert got.tolist() == [-101, -1, 0, 2, 100, 1001]


def test_float32_round_half_even(oracle):
    # jpeg2000/wavelet/dwt97.go:483-503 semantics == IEEE round-half-even (lrintf)
    v = np.array([0.5, 1.5, 2.5, -0.5, -1.5, -2.5, 0.49999997, 2.4999998, 1e6 + 0.5, -7.75, 3.0], np.float32)
    assert np.array_equal(oracle.convert_f32_to_i32(v), np.rint(v.astype(np.float64)).astype(np.int32))
    rng = np.random.default_rng(5)
    w = (rng.standard_normal(20000) * 3000).astype(np.float32)
    w[::7] = np.round(w[::7]) + 0.5
    assert np.array_equal(oracle.convert_f32_to_i32(w), np.rint(w.astype(np.float64)).astype(np.int32))


@pytest.mark.parametrize("w,h,levels,x0,y0,ew,eh", [
    (888, 459, 0, 0, 0, 888, 459), (888, 459, 1, 0, 0, 444, 230), (888, 459, 5, 0, 0, 28, 15),
    (512, 512, 4, 0, 0, 32, 32), (2, 1, 10, 0, 0, 1, 1),
    (7, 6, 2, 1, 0, 1, 2), (8, 7, 2, 0, 1, 2, 1), (-1, 7, 2, 0, 0, 0, 0),
])
def test_ll_dimensions(oracle, w, h, levels, x0, y0, ew, eh):
    # jpeg2000/wavelet/layout_test.go:5-56
    assert oracle.ll_dimensions(w, h, levels, x0, y0) == (ew, eh)


def test_tile_bounds(oracle):
    # jpeg2000/tile_assembler_test.go:10-193
    for iw, ih, nx, ny in [(256, 256, 1, 1), (512, 512, 2, 2), (600, 400, 3, 2), (500, 300, 2, 2)]:
        p = abi.inv_params(iw, ih, tile_width=256, tile_height=256)
        assert oracle.inv_tile_bounds(p, 0)[0] == nx * ny
    p = abi.inv_params(500, 300, tile_width=256, tile_height=256)
    assert oracle.inv_tile_bounds(p, 1)[1] == [256, 0, 500, 256]
    assert oracle.inv_tile_bounds(p, 2)[1] == [0, 256, 256, 300]
    assert oracle.inv_tile_bounds(p, 3)[1] == [256, 256, 500, 300]
    fp = abi.fwd_params(500, 300, tile_width=256, tile_height=256)
    assert oracle.fwd_tile_bounds(fp, 3) == (4, [256, 256, 500, 300])


def test_ict_int_api_tables(oracle):
    # jpeg2000/colorspace/rgb_test.go:9-194 (loose tables, tolerance 1..2)
    cases = [((0, 0, 0), (0, 0, 0), 1), ((255, 255, 255), (255, 0, 0), 1), ((255, 0, 0), (76, -43, 128), 2),
             ((0, 255, 0), (150, -84, -107), 2)]
    for (r, g, b), want, tol in cases:
        y, cb, cr = oracle.ict_forward([r], [g], [b])
        assert abs(int(y[0]) - want[0]) <= tol and abs(int(cb[0]) - want[1]) <= tol and abs(int(cr[0]) - want[2]) <= tol
    rng = np.random.default_rng(3)
    rgb = rng.integers(-128, 128, (3, 4096)).astype(np.int32)
    y, cb, cr = oracle.ict_forward(*rgb)
    r2, g2, b2 = oracle.ict_inverse(y, cb, cr)
    assert max(np.abs(r2 - rgb[0]).max(), np.abs(g2 - rgb[1]).max(), np.abs(b2 - rgb[2]).max()) <= 2


def test_rct_roundtrip_exact(oracle):
    # jpeg2000/mct_transform_test.go:8-56 pins RCT end-to-end as identity
    rng = np.random.default_rng(4)
    rgb = rng.integers(-32768, 32768, (3, 10000)).astype(np.int32)
    y, cb, cr = oracle.rct_forward(*rgb)
    assert np.array_equal(y, (rgb[0] + 2 * rgb[1] + rgb[2]) >> 2)
    back = oracle.rct_inverse(y, cb, cr)
    assert all(np.array_equal(a, b) for a, b in zip(back, rgb))


@pytest.mark.parametrize("size", [2, 3, 4, 5, 8, 15, 16, 17, 31, 32, 33, 64, 100, 127])
@pytest.mark.parametrize("even", [True, False])
def test_53_1d_perfect_reconstruction(oracle, size, even):
    # jpeg2000/wavelet/dwt53_test.go:9-85
    x = np.array([(i * 7 + 3) % 256 - 128 for i in range(size)], np.int32)
    assert np.array_equal(oracle.inv53_1d(oracle.fwd53_1d(x, even), even), x)


def test_53_odd_start_single_sample(oracle):
    # jpeg2000/wavelet/dwt53_test.go:52-85: n == 1, odd start doubles / halves
    assert oracle.fwd53_1d([21], even=False).tolist() == [42]
    assert oracle.inv53_1d([42], even=False).tolist() == [21]
    assert oracle.inv53_1d([-7], even=False).tolist() == [-3]  # Go `/` truncates toward zero
    assert oracle.fwd53_1d([21], even=True).tolist() == [21]


@pytest.mark.parametrize("w,h,levels,x0,y0", [
    (2, 2, 1, 0, 0), (100, 100, 1, 0, 0), (33, 17, 1, 0, 0), (17, 19, 1, 1, 1), (256, 256, 6, 0, 0),
    (64, 48, 3, 1, 2), (128, 128, 5, 0, 0),
])
def test_53_multilevel_identity(oracle, w, h, levels, x0, y0):
    # jpeg2000/wavelet/dwt53_test.go:88-256, wavelet_256_test.go, validation/dwt_precision_test.go:11-41
    rng = np.random.default_rng(42)
    a = rng.integers(-32768, 32768, (h, w)).astype(np.int32)
    assert np.array_equal(oracle.inv53(oracle.fwd53(a, levels, x0, y0), levels, x0, y0), a)


def test_97_1d_reference_vector_is_float32(oracle):
    # jpeg2000/wavelet/dwt97_test.go:8-20: the float64 wrapper exposes float32 arithmetic.
    x = np.array([0, 17, 33, 71, 129, 251, 502, 777, 1023], np.float64)
    via_f64 = oracle.fwd97_f64(x.reshape(1, -1), 1).reshape(-1)
    via_f32 = oracle.fwd97_1d(x.astype(np.float32), True)
    assert np.array_equal(via_f64, via_f32.astype(np.float64))
    # low-pass DC gain of the OpenJPEG-scaled analysis filter is 1 (invK * sqrt2-normalised K)
    c = oracle.fwd97_1d(np.full(16, 100.0, np.float32), True)
    assert np.allclose(c[:8], 100.0, atol=1e-3) and np.allclose(c[8:], 0.0, atol=1e-3)


def test_97_lossy_roundtrip_gain(oracle):
    # jpeg2000/wavelet/dwt97_test.go:458-500: forward+inverse is lossy; with OpenJPEG's
    # two_invK decode scaling the high-pass round-trip gain is 2, so the
    # un-quantized round trip is NOT the identity - it is the quantizer's band gain that undoes it.
    a = (np.arange(32 * 32) % 256).astype(np.float32).reshape(32, 32)
    r = oracle.convert_f32_to_i32(oracle.inv97(oracle.fwd97(a, 2), 2))
    assert np.any(r != a.astype(np.int32))


def test_signed_sub16_is_not_sign_extended(oracle):
    # encode `if v >= 2^(B-1) { v -= 2^B }` on the raw word; decode `+= 2^B`.
    p = abi.fwd_params(4, 1, 1, 12, True, num_levels=0, reversible=True)
    raw = np.array([0x0FFF, 0x0800, 0x07FF, 0xFFFF], "<u2")
    out = oracle.forward(p, raw.view(np.uint8))
    assert out.tolist() == [-1, -2048, 2047, 0xFFFF - 4096]
    ip = abi.inv_params(4, 1, 1, 12, True, num_levels=0, reversible=True)
    back = oracle.inverse(ip, np.array([-1, -2048, 2047, -5000], np.int32)).view("<u2")
    assert back.tolist() == [0x0FFF, 0x0800, 0x07FF, 0x0800]


def test_inverse_max_shift_rule(oracle):
    """applyInverseMaxShift (t2/tile_decoder.go:1113-1138): threshold rule, shift <= 0 is the identity, shift >= 31 zeroes,
    INT_MIN keeps Go's wrapped magnitude (negative, below every threshold: untouched)."""
    v = np.array([0, 1, -1, 7, 8, -8, 9, -9, 1023, 1024, -1024, 1 << 20, -(1 << 20), (1 << 31) - 1, -(1 << 31)], np.int64).astype(np.int32)
    assert np.array_equal(oracle.inverse_max_shift(v, 0), v)
    assert np.array_equal(oracle.inverse_max_shift(v, -3), v)
    assert not oracle.inverse_max_shift(v, 31).any() and not oracle.inverse_max_shift(v, 200).any()
    want3 = np.array([0, 1, -1, 7, 1, -1, 1, -1, 127, 128, -128, 1 << 17, -(1 << 17), (1 << 28) - 1, -(1 << 31)], np.int64).astype(np.int32)
    assert np.array_equal(oracle.inverse_max_shift(v, 3), want3)
    # an encoder-side MaxShift (ROI magnitudes << s, background below 2^s) is undone exactly
    rng = np.random.default_rng(5)
    for s_ in (1, 5, 12, 20):
        bg = rng.integers(-(1 << s_) + 1, 1 << s_, 1000).astype(np.int32)
        roi = rng.integers(-(1 << (30 - s_)) + 1, 1 << (30 - s_), 1000).astype(np.int32)
        roi[roi == 0] = 1
        sent = np.concatenate([bg, (roi.astype(np.int64) << s_).astype(np.int32)])
        back = oracle.inverse_max_shift(sent, s_)
        assert np.array_equal(back[:1000], bg) and np.array_equal(back[1000:], roi)
