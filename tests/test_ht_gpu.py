"""HTJ2K block decoder on the B200 (sm_100a build, through the C ABI) against the oracle: every OpenJPH interop codestream,
corrupted segments, generated streams at every block shape, and C1/C3-sized frames.  Bit-exact, status codes included."""
import numpy as np
import pytest

import ht_cases
import ht_parity as HP

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ht():
    import ht_oracle_lib
    return ht_oracle_lib.HtOracle()


@pytest.mark.parametrize("name,kind", ht_cases.fixtures())
def test_openjph_fixtures(ctx, ht, oracle, name, kind):
    HP.check_fixture(ctx, ht, oracle, name, kind, nframes=3 if "128x128" in name else 1)


@pytest.mark.parametrize("name,kind", ht_cases.fixtures()[::2])
def test_mutated_segments(ctx, ht, oracle, name, kind):
    HP.check_mutations(ctx, ht, oracle, name, kind, rounds=6, seed=11)


@pytest.mark.parametrize("w,h,levels,cbw,cbh", [(64, 64, 0, 64, 64), (70, 37, 1, 32, 32), (33, 65, 1, 64, 64), (40, 24, 0, 4, 4),
                                                (130, 9, 0, 128, 32), (9, 130, 0, 16, 256), (260, 4, 0, 1024, 4), (5, 300, 0, 4, 1024),
                                                (1, 7, 0, 8, 8), (7, 1, 0, 8, 8), (2, 2, 0, 4, 4), (512, 512, 3, 64, 64), (300, 200, 2, 32, 64)])
def test_random_streams(ctx, ht, oracle, w, h, levels, cbw, cbh):
    HP.check_random_streams(ctx, ht, oracle, w, h, levels, cbw, cbh, seed=w * 131 + h)


@pytest.mark.parametrize("w,h,levels,cbw,cbh,bits,density,comps,rev", [
    (64, 64, 0, 64, 64, 12, 0.7, 1, True), (96, 80, 2, 32, 32, 8, 0.3, 1, True), (75, 61, 2, 64, 64, 16, 0.9, 1, True),
    (128, 32, 1, 128, 32, 10, 0.5, 1, False), (40, 40, 1, 16, 16, 8, 0.6, 3, True), (33, 130, 1, 8, 512, 12, 1.0, 1, True),
    (150, 10, 0, 1024, 4, 9, 0.8, 1, True), (48, 48, 2, 64, 64, 8, 0.05, 3, False),
    (512, 512, 5, 64, 64, 16, 0.8, 1, True),       # C1-shaped
    (512, 384, 5, 64, 64, 8, 0.5, 3, False),       # C3-shaped, ICT + 9/7 behind the block decoder
    (2140, 300, 4, 64, 64, 16, 1.0, 1, False),     # odd DX width: edge blocks of every size
    (257, 255, 3, 32, 128, 14, 0.95, 1, True),
])
def test_generated_streams(ctx, ht, oracle, w, h, levels, cbw, cbh, bits, density, comps, rev):
    HP.check_generated(ctx, ht, oracle, w, h, levels, cbw, cbh, bits, density, seed=w + 7 * h, components=comps, reversible=rev)


def test_generated_batch(ctx, ht, oracle):
    HP.check_generated(ctx, ht, oracle, 256, 256, 4, 64, 64, 12, 0.7, seed=5, nframes=4)


def test_error_codes(ctx, ht, oracle):
    HP.check_error_codes(ctx, ht, oracle)


def test_unsupported_block_size_and_bad_offsets(ctx):
    import j2kb200
    from j2kb200 import abi
    ip = abi.inv_params(64, 64, 1, 8, False, num_levels=0, reversible=True, htj2k=True)
    rec = j2kb200.Context.ht_records([0], [4], [8], [7])
    with pytest.raises(Exception, match="4096"):
        ctx.ht_decode_blocks(ip, 1, np.zeros(16, np.uint8), rec, 128, 64)
    rec = j2kb200.Context.ht_records([10], [40], [8], [7])
    with pytest.raises(Exception, match="outside"):
        ctx.ht_decode_blocks(ip, 1, np.zeros(16, np.uint8), rec, 64, 64)


def test_async_ticket_and_pinned_buffers(ctx, ht, oracle):
    """j2k_submit_inverse_ht with library-owned pinned buffers, two tickets in flight, every frame compared"""
    from j2kb200 import abi
    from j2kb200.codec import Context
    fx = ht_cases.load("mono_u16_888x459", "fo_htj2k_lossless", oracle.codeblock_layout)
    h = fx["header"]
    ip = abi.inv_params(h.width, h.height, 1, h.depth[0], h.signed[0], num_levels=h.num_levels, reversible=True, htj2k=True)
    F = 5
    nb = len(fx["offsets"])
    rec = Context.ht_records(np.tile(fx["offsets"], F), np.tile(fx["lengths"], F), np.tile(fx["kmax"], F), np.tile(fx["mmsb"], F))
    stream = ctx.pinned(fx["stream"].size)
    stream[:] = fx["stream"]
    outs = [ctx.pinned(F * fx["raw"].size).reshape(F, -1) for _ in range(2)]
    sts = [ctx.pinned(F * nb * 4, np.int32).reshape(F, nb) for _ in range(2)]
    for o, s_ in zip(outs, sts):
        o[:] = 0
        s_[:] = -9
    t0 = ctx.submit_inverse_ht(ip, F, stream, rec, outs[0], sts[0], h.cbw, h.cbh)
    t1 = ctx.submit_inverse_ht(ip, F, stream, rec, outs[1], sts[1], h.cbw, h.cbh)
    ctx.wait(t1)
    ctx.wait(t0)
    for o, s_ in zip(outs, sts):
        assert not s_.any()
        for f in range(F):
            assert np.array_equal(o[f], fx["raw"]), f
    for a in [stream] + outs + sts:
        ctx.release(a)


# ---- encode side

@pytest.mark.parametrize("name,kind", ht_cases.fixtures())
def test_encode_reproduces_the_openjph_fixture_blocks(ctx, oracle, name, kind):
    HP.check_fixture_encode(ctx, oracle, name, kind)


@pytest.mark.parametrize("w,h,comps,bits,levels,cbw,cbh,rev", [
    (64, 64, 1, 8, 0, 64, 64, True), (96, 80, 1, 12, 2, 32, 32, True), (75, 61, 1, 16, 2, 64, 64, True), (128, 32, 1, 10, 1, 128, 32, False),
    (40, 40, 3, 8, 1, 16, 16, True), (33, 130, 1, 12, 1, 8, 512, True), (150, 10, 1, 9, 0, 1024, 4, True), (48, 48, 3, 8, 2, 64, 64, False),
    (1, 1, 1, 8, 0, 4, 4, True), (2, 7, 1, 8, 0, 4, 4, True),
    (512, 512, 1, 16, 5, 64, 64, True),        # C1-shaped
    (512, 384, 3, 8, 5, 64, 64, False),        # C3-shaped: ICT + 9/7 + quantization in front of the block encoder
    (2140, 300, 1, 16, 4, 64, 64, False),      # odd DX width
    (257, 255, 1, 14, 3, 32, 128, True),
])
def test_encode(ctx, ht, oracle, w, h, comps, bits, levels, cbw, cbh, rev):
    HP.check_encode(ctx, ht, oracle, w, h, comps, bits, levels, cbw, cbh, seed=w * 3 + h, reversible=rev)


def test_encode_batches_tiles_tight_kmax_and_round_trip(ctx, ht, oracle):
    HP.check_encode(ctx, ht, oracle, 256, 256, 1, 12, 4, 64, 64, seed=4, nframes=5)
    HP.check_encode(ctx, ht, oracle, 70, 50, 3, 8, 2, 32, 32, seed=5, tile=(32, 32))
    HP.check_encode(ctx, ht, oracle, 64, 48, 1, 12, 1, 64, 64, seed=6, base=6)
    # encode -> decode on the device == the frames (lossless 5/3)
    from j2kb200 import abi
    stream, rec, fp, frames = HP.check_encode(ctx, ht, oracle, 300, 200, 1, 12, 3, 64, 64, seed=7, nframes=3)
    ip = abi.inv_params(300, 200, 1, 12, False, num_levels=3, reversible=True, htj2k=True)
    px, st = ctx.inverse_ht(ip, 3, np.concatenate([stream, np.zeros(16, np.uint8)]), rec)
    assert not st.any() and np.array_equal(px, frames)


def test_encode_sub_batches_and_small_capacity(ctx, ht, oracle, monkeypatch):
    """more frames than one sub-batch holds (the lagged download path), and a stream buffer that is too small"""
    import j2kb200
    from j2kb200 import abi
    stream, rec, fp, frames = HP.check_encode(ctx, ht, oracle, 1024, 1024, 1, 12, 5, 64, 64, seed=8, nframes=40)
    small = np.empty(stream.size // 2, np.uint8)
    with pytest.raises(j2kb200.J2KError) as e:
        ctx.forward_ht(fp, frames, HP.band_kmax_table(1, 5, 14), out=small)
    assert e.value.code == abi.J2K_ERR_SIZE
    with pytest.raises(j2kb200.J2KError, match="Kmax"):
        ctx.forward_ht(fp, frames[:1], np.full((1, 16), 31, np.uint8))


def test_encode_sharded_over_the_devices_of_a_context(ht, oracle, monkeypatch):
    """j2k_forward_ht shards the frames over the devices of the context in contiguous blocks, each with its own
    lagged sub-batch pipeline.  Every visible device, plus device 0 once more, so that the path runs on a one-GPU box too; the
    segments are located by the records (with several devices they are appended as the host collects them)."""
    import torch

    import j2kb200
    from j2kb200 import abi
    devs = list(range(torch.cuda.device_count())) + [0]
    with j2kb200.Context(devices=devs) as mctx:
        assert mctx.device_count == len(devs)
        stream, rec, fp, frames = HP.check_encode(mctx, ht, oracle, 300, 200, 1, 12, 3, 64, 64, seed=7, nframes=2 * len(devs) + 1, ordered=False)
        ip = abi.inv_params(300, 200, 1, 12, False, num_levels=3, reversible=True, htj2k=True)
        px, st = mctx.inverse_ht(ip, frames.shape[0], np.concatenate([stream, np.zeros(16, np.uint8)]), rec)
        assert not st.any() and np.array_equal(px, frames)
        monkeypatch.setenv("J2K_HT_SUBBATCH_MSAMPLES", "1")   # several sub-batches per device: the lagged collection, interleaved
        HP.check_encode(mctx, ht, oracle, 1024, 1024, 1, 12, 5, 64, 64, seed=8, nframes=9, ordered=False)
        HP.check_encode(mctx, ht, oracle, 128, 96, 3, 8, 2, 32, 32, seed=3, nframes=1, reversible=False, ordered=False)   # fewer frames than devices
