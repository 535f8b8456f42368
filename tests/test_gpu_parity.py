"""GPU parity tests proper: libj2kb200.so (sm_100a) through the C ABI vs the oracle, same seeded inputs.

Bit-exact everywhere (the 9/7 and ICT paths are built without FMA contraction, so they are exact too;
the test would report the max-abs sub-band error otherwise)."""
import json
import os

import numpy as np
import pytest

import parity_cases as PC
from j2kb200 import abi

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("w,h,levels,x0,y0", [
    (16, 16, 1, 0, 0), (64, 64, 3, 0, 0), (17, 19, 3, 0, 0), (33, 17, 2, 1, 0), (20, 9, 4, 0, 1), (64, 48, 3, 1, 2),
    (7, 1, 2, 0, 0), (1, 9, 3, 1, 1), (5, 5, 6, 3, 3), (2, 2, 3, 1, 1), (1, 1, 2, 1, 0), (13, 2, 5, 2, 7),
    (256, 256, 6, 0, 0), (333, 211, 6, 0, 0), (888, 459, 5, 0, 0), (127, 129, 5, 1, 2), (1024, 768, 7, 0, 0), (100, 100, 1, 0, 0),
])
def test_wavelet_api(ctx, oracle, w, h, levels, x0, y0):
    PC.check_wavelet_api(ctx, oracle, w, h, levels, x0, y0, seed=w * 1000 + h)


@pytest.mark.parametrize("w,h,c,bits,signed,L,rev,kind", [
    (512, 512, 1, 16, True, 5, True, "smooth"),      # C1
    (512, 512, 1, 16, False, 5, True, "noise"),
    (1024, 1024, 1, 12, False, 6, False, "smooth"),  # C2 shape at 1/16 area
    (1024, 512, 1, 12, False, 6, False, "noise"),
    (512, 384, 3, 8, False, 5, False, "smooth"),     # C3 (i)
    (512, 384, 3, 8, False, 5, True, "smooth"),      # C3 (ii)
    (333, 211, 1, 12, False, 6, False, "noise"), (127, 129, 3, 8, False, 5, True, "noise"),
    (888, 459, 1, 16, False, 5, True, "smooth"), (640, 480, 3, 16, False, 4, False, "smooth"),
    (300, 200, 2, 8, False, 3, True, "noise"), (300, 200, 4, 12, True, 3, False, "noise"),
    (64, 64, 1, 8, False, 0, True, "noise"), (64, 64, 1, 8, False, 0, False, "noise"), (1, 1, 1, 8, False, 2, False, "noise"),
    (2048, 64, 1, 16, False, 3, True, "smooth"), (64, 2048, 1, 12, False, 3, False, "smooth"),
    (512, 384, 3, 8, True, 4, False, "noise"), (768, 256, 3, 8, False, 5, False, "noise"), (512, 256, 3, 7, False, 3, False, "noise"),  # ICT + 9/7: signed, noise, 7-bit
])
def test_pipeline(ctx, oracle, w, h, c, bits, signed, L, rev, kind):
    PC.check_pipeline(ctx, oracle, w, h, c, bits, signed, L, rev, kind=kind, seed=w + h)


@pytest.mark.parametrize("w,h,c,tile,L,rev", [
    (1024, 768, 1, (256, 256), 5, True), (1000, 700, 1, (256, 256), 5, False), (700, 500, 3, (256, 128), 4, False),
    (513, 257, 3, (256, 256), 5, True), (500, 500, 1, (333, 177), 3, False), (2048, 2048, 3, (1024, 1024), 7, False),
])
def test_tiles(ctx, oracle, w, h, c, tile, L, rev):
    PC.check_pipeline(ctx, oracle, w, h, c, 8, False, L, rev, tile=tile)


def test_htj2k_and_fused_t1_shift(ctx, oracle):
    PC.check_pipeline(ctx, oracle, 400, 360, 1, 16, False, 5, True, fuse=True)
    PC.check_pipeline(ctx, oracle, 400, 360, 1, 16, False, 5, True, htj2k=True, fuse=True)
    PC.check_pipeline(ctx, oracle, 400, 360, 1, 8, False, 5, False, htj2k=True)
    PC.check_pipeline(ctx, oracle, 400, 360, 3, 8, False, 4, False, steps_kind="quality")


@pytest.mark.parametrize("case", ["int_matrix", "q13_matrix", "q13_4comp", "bindings"])
@pytest.mark.parametrize("rev", [True, False])
def test_custom_mct(ctx, oracle, case, rev):
    PC.check_custom_mct(ctx, oracle, 640, 360, 8 if rev else 12, 4, rev, case)


def test_custom_mct_tiled(ctx, oracle):
    PC.check_custom_mct(ctx, oracle, 700, 500, 12, 3, True, "bindings", tile=(256, 256))


@pytest.mark.parametrize("w,h,c,bits,signed,L,rev", [
    (512, 512, 1, 16, True, 5, True), (640, 480, 3, 8, False, 4, False), (1024, 256, 1, 12, False, 5, False), (333, 217, 3, 10, True, 3, True),
])
def test_planar_entry(ctx, oracle, w, h, c, bits, signed, L, rev):
    PC.check_planar(ctx, oracle, w, h, c, bits, signed, L, rev)


def test_package_api(ctx, oracle):
    PC.check_package_api(ctx, oracle, n=1_000_003)


@pytest.mark.parametrize("w,h,c,bits,L,rev,tile,cb", [
    (512, 512, 1, 16, 5, True, (0, 0), (64, 64)),      # C1 with the default 64x64 code-blocks
    (1024, 768, 1, 12, 6, False, (0, 0), (64, 64)),    # C2-shaped
    (640, 480, 3, 8, 5, False, (0, 0), (32, 32)), (513, 257, 3, 8, 5, True, (256, 256), (64, 64)),
    (333, 211, 1, 12, 4, False, (0, 0), (16, 64)), (127, 129, 1, 16, 5, True, (0, 0), (4, 4)), (100, 100, 1, 8, 0, True, (0, 0), (32, 32)),
])
def test_code_block_interface(ctx, oracle, w, h, c, bits, L, rev, tile, cb):
    PC.check_blocks(ctx, oracle, w, h, c, bits, L, rev, tile=tile, cb=cb, nframes=3)


def test_code_block_interface_htj2k_and_errors(ctx, oracle):
    import j2kb200
    PC.check_blocks(ctx, oracle, 400, 360, 1, 12, 4, True, cb=(64, 64), htj2k=True)
    PC.check_blocks(ctx, oracle, 400, 360, 1, 12, 4, False, cb=(64, 64), htj2k=True)
    fp = abi.fwd_params(64, 64, 1, 8, False, num_levels=2)
    with pytest.raises(j2kb200.J2KError) as e:
        ctx.forward_blocks(fp, np.zeros((1, 64 * 64), np.uint8), 48, 64)
    assert "invalid code-block width" in str(e.value)


@pytest.mark.parametrize("w,h,c,bits,L,rev", [
    (4160, 512, 1, 12, 6, False), (1000, 600, 1, 16, 5, True), (1048, 520, 3, 8, 5, False), (1048, 520, 3, 8, 5, True), (2056, 264, 1, 8, 6, False),
])
def test_hybrid_plans(ctx, oracle, w, h, c, bits, L, rev):
    """Widths that stay a multiple of 8 only for the first levels: persistent launch for those, per-level kernels beyond."""
    PC.check_pipeline(ctx, oracle, w, h, c, bits, False, L, rev, seed=w)


def test_interop_raws(ctx, oracle):
    man = json.load(open(os.path.join(HERE, "golden", "interop", "manifest.json")))
    for fx in man["fixtures"]:
        raw = np.fromfile(os.path.join(HERE, "golden", "interop", fx["name"] + ".raw"), np.uint8)
        C = fx["components"]
        mct = abi.MCT_RCT if C == 3 else abi.MCT_NONE
        fp = abi.fwd_params(fx["width"], fx["height"], C, fx["bitsStored"], fx["signed"], num_levels=5, reversible=True, mct_mode=mct)
        ip = abi.inv_params(fx["width"], fx["height"], C, fx["bitsStored"], fx["signed"], num_levels=5, reversible=True, mct_mode=mct)
        co = ctx.forward(fp, raw)
        assert np.array_equal(co, oracle.forward(fp, raw)), fx["name"]
        assert np.array_equal(ctx.inverse(ip, co), raw), fx["name"]


def test_batch_matches_single_and_roundtrip_at_full_size(ctx, oracle):
    # C4-shaped batch: 64 frames 512x512 16-bit 5/3 L=5; frame 0 and 63 vs oracle, all frames by the identity property
    rng = np.random.default_rng(1000)
    n, w, h = 64, 512, 512
    frames = rng.integers(0, 65536, (n, h * w), dtype=np.uint16).view(np.uint8).reshape(n, -1)
    fp = abi.fwd_params(w, h, 1, 16, False, num_levels=5, reversible=True)
    ip = abi.inv_params(w, h, 1, 16, False, num_levels=5, reversible=True)
    co = ctx.forward_batch(fp, frames)
    for f in (0, n - 1):
        assert np.array_equal(co[f], oracle.forward(fp, frames[f]))
    back = ctx.inverse_batch(ip, co)
    assert np.array_equal(back, frames)
    # C2 full size, 9/7 L=6: oracle on the full frame takes seconds
    img = PC.synth(rng, 4096, 4096, 1, 12)
    es, ds = PC.steps_for(oracle, 6, 12)
    fp = abi.fwd_params(4096, 4096, 1, 12, False, num_levels=6, reversible=False, steps=es)
    got = ctx.forward(fp, PC.raw_bytes(img))
    want = oracle.forward(fp, PC.raw_bytes(img))
    assert np.count_nonzero(got != want) == 0


def test_full_size_series_and_tile_block_properties(ctx, oracle):
    """BASELINE sizes the oracle cannot cover in seconds, through size-independent properties."""
    # C4: the whole 2000-frame 512x512 16-bit series, 5/3 lossless: encode -> decode is the identity, frame by frame
    rng = np.random.default_rng(4000)
    n, w, h = 2000, 512, 512
    frames = rng.integers(0, 65536, (n, h * w), dtype=np.uint16).view(np.uint8).reshape(n, -1)
    fp = abi.fwd_params(w, h, 1, 16, False, num_levels=5, reversible=True)
    ip = abi.inv_params(w, h, 1, 16, False, num_levels=5, reversible=True)
    co = ctx.forward_batch(fp, frames)
    assert np.array_equal(ctx.inverse_batch(ip, co), frames)
    for f in (0, 999, 1999):  # and three frames against the oracle
        assert np.array_equal(co[f], oracle.forward(fp, frames[f]))
    del co
    # C5: one GPU's block of the slide, 128 tiles of 1024x1024 RGB (8192 x 16384), 7 levels.
    W5, H5 = 8192, 16384
    pool = [PC.synth(np.random.default_rng(5000 + k), 1024, 1024, 3, 8) for k in range(4)]
    img = np.empty((H5, W5, 3), np.uint8)
    for t in range(128):  # every tile differs: a pool tile, shifted and offset by the tile number
        tyy, txx = t // 8, t % 8
        img[tyy * 1024:(tyy + 1) * 1024, txx * 1024:(txx + 1) * 1024] = np.roll(pool[t % 4], t, axis=1) // 2 + (t % 100)
    raw = img.reshape(-1)
    # (i) RCT + 5/3: identity
    fp = abi.fwd_params(W5, H5, 3, 8, False, 1024, 1024, 7, True, False, abi.MCT_RCT)
    ip = abi.inv_params(W5, H5, 3, 8, False, 1024, 1024, 7, True, False, abi.MCT_RCT)
    co = ctx.forward(fp, raw)
    assert np.array_equal(ctx.inverse(ip, co), raw)
    # every tile is transformed independently: tile 77 alone gives the same coefficients as inside the image
    tx, ty = 77 % 8, 77 // 8
    tile = np.ascontiguousarray(img[ty * 1024:(ty + 1) * 1024, tx * 1024:(tx + 1) * 1024]).reshape(-1)
    fpt = abi.fwd_params(1024, 1024, 3, 8, False, 0, 0, 7, True, False, abi.MCT_RCT)
    t_co = ctx.forward(fpt, tile)
    assert np.array_equal(co[77 * 3 * 1024 * 1024:78 * 3 * 1024 * 1024], t_co)
    assert np.array_equal(t_co, oracle.forward(fpt, tile))
    # (ii) ICT + 9/7 with the OpenJPEG default steps: one tile against the oracle, the whole block by the lossy bound
    es, ds = PC.steps_for(oracle, 7, 8)
    fp = abi.fwd_params(W5, H5, 3, 8, False, 1024, 1024, 7, False, False, abi.MCT_ICT, es)
    ip = abi.inv_params(W5, H5, 3, 8, False, 1024, 1024, 7, False, False, abi.MCT_ICT, ds)
    co = ctx.forward(fp, raw)
    fpt = abi.fwd_params(1024, 1024, 3, 8, False, 0, 0, 7, False, False, abi.MCT_ICT, es)
    assert np.array_equal(co[77 * 3 * 1024 * 1024:78 * 3 * 1024 * 1024], oracle.forward(fpt, tile))
    t1 = np.empty_like(co)
    for k in range(0, co.size, 1 << 24):  # chunked: the stand-in for T1 works in int64
        t1[k:k + (1 << 24)] = PC.M.t1_emulate(co[k:k + (1 << 24)], False)
    back = ctx.inverse(ip, t1)
    worst = 0
    for k in range(0, raw.size, 1 << 26):
        worst = max(worst, int(np.abs(back[k:k + (1 << 26)].astype(np.int16) - raw[k:k + (1 << 26)].astype(np.int16)).max()))
    assert worst <= 12, worst  # the reference's own lossy end-to-end bound (<= 12 at default steps)


def test_async_tickets(ctx, oracle):
    rng = np.random.default_rng(5)
    n, w, h = 8, 256, 256
    fp = abi.fwd_params(w, h, 1, 16, False, num_levels=5, reversible=True)
    fin = ctx.pinned(n * w * h * 2).reshape(n, -1)
    fout = ctx.pinned(n * w * h * 4, np.int32).reshape(n, -1)
    fin[:] = rng.integers(0, 256, fin.shape, dtype=np.uint8)
    t = ctx.submit_forward(fp, fin, fout)
    ctx.wait(t)
    assert np.array_equal(fout[3], oracle.forward(fp, fin[3]))
    ctx.release(fin)
    ctx.release(fout)


def test_errors(ctx):
    import j2kb200
    fp = abi.fwd_params(64, 64, 1, 8, False, num_levels=3)
    with pytest.raises(j2kb200.J2KError) as e:
        ctx.forward(fp, np.zeros(100, np.uint8))
    assert e.value.code == abi.J2K_ERR_SIZE and "insufficient pixel data" in str(e.value)
    bad = abi.fwd_params(0, 64, 1, 8, False)
    with pytest.raises(j2kb200.J2KError):
        ctx.forward(bad, np.zeros(10, np.uint8))


def test_in_process_multi_device_sharding(oracle):
    """j2k_init with every visible device: a host batch is cut into contiguous frame blocks, one per GPU, no collective
    (SURVEY 8e).  Runs on any box; with one GPU it degenerates to the single-device path."""
    import torch

    import j2kb200
    nd = torch.cuda.device_count()
    rng = np.random.default_rng(77)
    n, w, h = 4 * max(nd, 1) + 1, 256, 192  # uneven split on purpose
    frames = rng.integers(0, 4096, (n, h * w), dtype=np.uint16).view(np.uint8).reshape(n, -1)
    es, ds = PC.steps_for(oracle, 4, 12)
    fp = abi.fwd_params(w, h, 1, 12, False, num_levels=4, reversible=False, steps=es)
    fl = abi.fwd_params(w, h, 1, 12, False, num_levels=4, reversible=True)
    il = abi.inv_params(w, h, 1, 12, False, num_levels=4, reversible=True)
    with j2kb200.Context(devices=list(range(nd))) as mctx:
        assert mctx.device_count == nd
        co = mctx.forward_batch(fp, frames)
        for f in range(n):
            assert np.array_equal(co[f], oracle.forward(fp, frames[f])), f"frame {f}"
        assert np.array_equal(mctx.inverse_batch(il, mctx.forward_batch(fl, frames)), frames)
        blocks, nb = mctx.forward_blocks(fl, frames, 32, 32)
        one = j2kb200.Context(devices=[0])
        b1, n1 = one.forward_blocks(fl, frames, 32, 32)
        one.close()
        assert np.array_equal(blocks, b1) and np.array_equal(nb, n1)


def test_c1_ct1_j2ki_image(ctx, oracle):
    """BASELINE config C1 on its own image: the reference's test-data/CT1_J2KI frame (512x512 signed 16-bit, fixture decoded
    by OpenJPEG 2.5.4, tests/golden/make_golden.py), 5/3 lossless 5 levels round trip, and the 9/7 path."""
    ob = np.load(os.path.join(HERE, "golden", "ct1_j2ki.npz"))["offset_binary"]
    raw = (ob.astype(np.int32) - 32768).astype("<i2").view(np.uint8).reshape(-1)
    fp = abi.fwd_params(512, 512, 1, 16, True, num_levels=5, reversible=True)
    ip = abi.inv_params(512, 512, 1, 16, True, num_levels=5, reversible=True)
    co = ctx.forward(fp, raw)
    assert np.array_equal(co, oracle.forward(fp, raw))
    assert np.array_equal(ctx.inverse(ip, co), raw)
    es, ds = PC.steps_for(oracle, 5, 16)
    fp = abi.fwd_params(512, 512, 1, 16, True, num_levels=5, reversible=False, steps=es)
    ip = abi.inv_params(512, 512, 1, 16, True, num_levels=5, reversible=False, steps=ds)
    co = ctx.forward(fp, raw)
    assert np.array_equal(co, oracle.forward(fp, raw))
    t1 = PC.M.t1_emulate(co)
    assert np.array_equal(ctx.inverse(ip, t1), oracle.inverse(ip, t1))


def test_c5_full_slide_single_call(ctx, oracle):
    """BASELINE config C5 at its full size in ONE call: 32768x32768 RGB 8-bit as 1024 tiles of 1024x1024, 7 levels
    (3.2 Gsamples: every offset beyond 2^31).  RCT + 5/3: the round trip is the identity; one tile against the oracle."""
    rng = np.random.default_rng(5005)
    cell = PC.synth(rng, 2048, 2048, 3, 8)                       # 2x2 tiles of distinct content ...
    img = np.tile(cell, (16, 16, 1))                             # ... repeated over the slide
    img[5 * 1024:6 * 1024, 9 * 1024:10 * 1024] = PC.synth(rng, 1024, 1024, 3, 8, kind="noise")  # and one odd tile
    raw = img.reshape(-1)
    W5 = H5 = 32768
    fp = abi.fwd_params(W5, H5, 3, 8, False, 1024, 1024, 7, True, False, abi.MCT_RCT)
    ip = abi.inv_params(W5, H5, 3, 8, False, 1024, 1024, 7, True, False, abi.MCT_RCT)
    co = ctx.forward(fp, raw)
    t = 5 * 32 + 9
    tile = np.ascontiguousarray(img[5 * 1024:6 * 1024, 9 * 1024:10 * 1024]).reshape(-1)
    fpt = abi.fwd_params(1024, 1024, 3, 8, False, 0, 0, 7, True, False, abi.MCT_RCT)
    assert np.array_equal(co[t * 3 * 1024 * 1024:(t + 1) * 3 * 1024 * 1024], oracle.forward(fpt, tile))
    back = ctx.inverse(ip, co)
    del co
    for k in range(0, raw.size, 1 << 28):
        assert np.array_equal(back[k:k + (1 << 28)], raw[k:k + (1 << 28)])
    del back
    # ICT + 9/7 at the same size: the odd tile and the last tile against the oracle
    es, _ = PC.steps_for(oracle, 7, 8)
    fp = abi.fwd_params(W5, H5, 3, 8, False, 1024, 1024, 7, False, False, abi.MCT_ICT, es)
    fpt = abi.fwd_params(1024, 1024, 3, 8, False, 0, 0, 7, False, False, abi.MCT_ICT, es)
    co = ctx.forward(fp, raw)
    assert np.array_equal(co[t * 3 * 1024 * 1024:(t + 1) * 3 * 1024 * 1024], oracle.forward(fpt, tile))
    last = np.ascontiguousarray(img[31 * 1024:, 31 * 1024:]).reshape(-1)
    assert np.array_equal(co[1023 * 3 * 1024 * 1024:], oracle.forward(fpt, last))


def test_package_api_x1(ctx, oracle):
    PC.check_package_api_x1(ctx, oracle)


def test_random_geometry_sweep(ctx, oracle):
    """400 seeded random configurations (size, components, depth, sign, levels, wavelet, tiles, HTJ2K, fused shift)."""
    for k, case in enumerate(PC.random_geometry_cases(400, 20261019, 1100, 300)):
        try:
            PC.check_random_case(ctx, oracle, case, 7000 + k)
        except AssertionError as e:
            raise AssertionError(f"case {k} {case}: {e}") from e


@pytest.mark.parametrize("w,h,c,bits,L,rev,xo,yo,tile,xto,yto", [
    (64, 48, 1, 8, 3, True, 1, 0, (0, 0), 0, 0), (64, 48, 1, 12, 3, False, 0, 1, (0, 0), 0, 0), (70, 50, 3, 8, 2, True, 1, 2, (32, 32), 1, 1),
    (70, 50, 3, 8, 2, False, 3, 5, (32, 32), 2, 3), (33, 17, 1, 16, 4, True, 7, 7, (16, 16), 0, 0), (40, 40, 2, 8, 2, False, 5, 2, (0, 0), 0, 0),    (640, 480, 1, 12, 5, False, 1, 1, (0, 0), 0, 0), (513, 257, 3, 8, 5, True, 3, 2, (256, 256), 1, 2), (1000, 700, 1, 16, 5, True, 11, 6, (256, 256), 5, 3),
])
def test_inverse_with_image_and_tile_offsets(ctx, oracle, w, h, c, bits, L, rev, xo, yo, tile, xto, yto):
    PC.check_inverse_with_offsets(ctx, oracle, w, h, c, bits, L, rev, xo, yo, tile, xto, yto)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("w,h,c,bits,L,rev,nframes,group_ks,lag", [
    (520, 264, 1, 16, 4, True, 24, 128, 1), (520, 264, 1, 12, 4, False, 40, 256, 2), (264, 136, 3, 8, 3, False, 30, 128, 3),
    (264, 136, 3, 8, 3, True, 18, 64, 1),
])
def test_group_pipelined_job_order(ctx, oracle, w, h, c, bits, L, rev, nframes, group_ks, lag, capfd):
    """The optional group-pipelined job order of the persistent launch (ring_schedule) under real concurrency: consumers
    are claimed `lag` item groups behind their producers (lag 1: warps do park on unfinished producers), results unchanged."""
    PC.check_pipelined_order(ctx, oracle, w, h, c, bits, L, rev, nframes, group_ks, lag, capfd)


@pytest.mark.parametrize("w,h,c,bits,L,rev,shifts,tile,cb", [
    (512, 384, 1, 16, 5, True, [7], (0, 0), (64, 64)), (1000, 700, 1, 12, 5, False, [10], (256, 256), (32, 32)),
    (513, 257, 3, 8, 4, True, [3, 0, 9], (256, 256), (64, 64)), (640, 480, 3, 8, 4, False, [0, 5, 31], (0, 0), (32, 64)),
    (127, 129, 1, 8, 3, True, [1], (0, 0), (4, 4)),
])
def test_code_block_interface_roi(ctx, oracle, w, h, c, bits, L, rev, shifts, tile, cb):
    """Decode-side MaxShift ROI fused into the block scatter."""
    PC.check_blocks_roi(ctx, oracle, w, h, c, bits, L, rev, shifts, tile=tile, cb=cb, nframes=3)


@pytest.mark.parametrize("w,h,c,bits,L,rev,chunk", [
    (512, 616, 1, 12, 5, False, 64), (512, 1000, 1, 16, 5, True, 64), (264, 600, 3, 8, 3, False, 64), (264, 520, 3, 8, 3, True, 64),
    (520, 1032, 1, 12, 4, False, 128), (528, 1040, 1, 16, 4, True, 256), (272, 776, 3, 8, 3, False, 128),
])
def test_tall_chunks(ctx, oracle, w, h, c, bits, L, rev, chunk):
    """Chunk heights of 64 row pairs (what the big batches of bench.py and the BASELINE configs run with) and above,
    forced on single frames and compared with the oracle: chunk seams, warm-up rows, partial last chunks."""
    PC.check_tall_chunks(ctx, oracle, w, h, c, bits, L, rev, chunk)


@pytest.mark.parametrize("w,h,c,bits,signed,L,rev", [
    (2140, 1760, 1, 16, False, 6, False), (2140, 1760, 1, 12, False, 5, True), (2022, 2022, 1, 12, False, 6, False), (2022, 2022, 1, 16, True, 5, True),
    (2023, 1001, 1, 8, False, 5, False), (2023, 1001, 1, 8, False, 5, True), (1766, 2140, 1, 10, False, 6, False), (4097, 333, 1, 16, False, 4, True),
    (1000, 600, 1, 16, False, 5, False), (250, 250, 1, 8, True, 3, True),
])
def test_general_alignment_ring_variant(ctx, oracle, w, h, c, bits, signed, L, rev):
    """CR / DX detector sizes (2140 x 1760, 2022 x 2022) and other widths that are not a multiple of 8: the general-alignment
    variant of the persistent kernels (TMA copies from the 16-byte boundary below each row, per-row phases, masked stores)."""
    PC.check_pipeline(ctx, oracle, w, h, c, bits, signed, L, rev, kind="noise", seed=w + h)
    PC.check_pipeline(ctx, oracle, w, h, c, bits, signed, L, rev, kind="smooth", seed=w)


def test_general_alignment_batch_and_device_buffers(ctx, oracle):
    """A batch of odd-sized frames (every frame starts at its own alignment) through the host path and the device path."""
    import torch
    rng = np.random.default_rng(91)
    w, h, n = 2022, 333, 5   # frame bytes = 2022 * 333 * 2: not a multiple of 16, so frames 1.. start unaligned
    fp, ip = PC.fwd_inv_params(w, h, 1, 12, False, 5, True, oracle)
    frames = np.stack([PC.raw_bytes(PC.synth(rng, h, w, 1, 12, False, "noise")) for _ in range(n)])
    co = ctx.forward_batch(fp, frames)
    for f in range(n):
        assert np.array_equal(co[f], oracle.forward(fp, frames[f])), f
    assert np.array_equal(ctx.inverse_batch(ip, co), frames)
    d_in = torch.from_numpy(frames).cuda()
    d_co = torch.empty((n, w * h), dtype=torch.int32, device="cuda")
    d_px = torch.empty_like(d_in)
    ctx.forward_device(fp, n, d_in.data_ptr(), frames.shape[1], d_co.data_ptr())
    ctx.inverse_device(ip, n, d_co.data_ptr(), d_px.data_ptr(), frames.shape[1])
    torch.cuda.synchronize()
    assert np.array_equal(d_co.cpu().numpy(), co) and np.array_equal(d_px.cpu().numpy(), frames)


@pytest.mark.parametrize("w,h,c,bits,L,rev,tile,cb,masked", [(512, 384, 1, 12, 4, True, (0, 0), (32, 32), True), (640, 480, 1, 16, 5, False, (0, 0), (64, 64), False), (300, 200, 3, 8, 3, True, (128, 128), (16, 16), True), (333, 211, 3, 8, 3, False, (0, 0), (32, 32), True)])
def test_code_block_interface_roi_general_scaling(ctx, oracle, w, h, c, bits, L, rev, tile, cb, masked):
    """SURVEY 8f rank 3, second half: inverse general scaling (RGN Srgn = 1) fused into the block scatter, whole-block and masked."""
    PC.check_blocks_roi_general(ctx, oracle, w, h, c, bits, L, rev, tile=tile, cb=cb, masked=masked)
    PC.check_blocks_roi_general(ctx, oracle, w, h, c, bits, L, rev, tile=tile, cb=cb, masked=masked, maxshift=[3] * c, seed=5)


@pytest.mark.parametrize("w,h,bits,L,nframes,tile,chunk", [
    (1024, 600, 8, 4, 3, (0, 0), 0), (512, 520, 8, 3, 9, (0, 0), 64), (1024, 256, 16, 4, 2, (0, 0), 0), (2048, 1024, 8, 3, 2, (512, 512), 0),
    (768, 330, 12, 3, 2, (0, 0), 16), (256, 37, 8, 2, 5, (0, 0), 0), (2048, 2048, 8, 5, 2, (0, 0), 0),
])
def test_one_producer_rgb97_forward(ctx, oracle, w, h, bits, L, nframes, tile, chunk):
    """fwd3w_kernel (one converting producer warp + three single-component consumers per CTA) with real concurrency: many
    CTAs, exchange-ring wrap-around, job triples of the coarser levels waiting on the per-component counters."""
    PC.check_one_producer_forward(ctx, oracle, w, h, bits, L, nframes, tile, chunk)


def test_failed_device_is_removed_from_the_round_robin(oracle, monkeypatch):
    """The same on the real runtime: a context with two slots on GPU 0, slot 1 failing (fault injection), then slot 0."""
    PC.check_failed_device(oracle, monkeypatch)
