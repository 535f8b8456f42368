"""HTJ2K block decoder: the product sources compiled for the CPU emulator (tests/emu) against the oracle.  CPU only; the
`-m gpu` suite (test_ht_gpu.py) repeats these on the sm_100a build with every fixture and larger sweeps."""
import pytest

import ht_parity as HP


@pytest.fixture(scope="module")
def ectx():
    import emu_lib
    import j2kb200
    c = j2kb200.Context(lib_path=emu_lib.build())
    yield c
    c.close()


@pytest.fixture(scope="module")
def ht():
    import ht_oracle_lib
    return ht_oracle_lib.HtOracle()


@pytest.mark.parametrize("name,kind", [("mono_u8_127x129", "fo_htj2k_lossless"), ("mono_s16_128x128", "fo_htj2k_lossless_rpcl"),
                                       ("mono_u16_128x128", "fo_htj2k_lossless"), ("rgb_u8_127x129", "fo_htj2k_lossless_rpcl")])
def test_fixtures(ectx, ht, oracle, name, kind):
    HP.check_fixture(ectx, ht, oracle, name, kind)


def test_two_frames(ectx, ht, oracle):
    HP.check_fixture(ectx, ht, oracle, "mono_u8_128x128", "fo_htj2k_lossless", nframes=2)


def test_mutated_segments(ectx, ht, oracle):
    HP.check_mutations(ectx, ht, oracle, "mono_u16_128x128", "fo_htj2k_lossless", rounds=3, seed=1)
    HP.check_mutations(ectx, ht, oracle, "rgb_u8_127x129", "fo_htj2k_lossless", rounds=2, seed=2)


@pytest.mark.parametrize("w,h,levels,cbw,cbh", [(64, 64, 0, 64, 64), (70, 37, 1, 32, 32), (33, 65, 1, 64, 64), (40, 24, 0, 4, 4),
                                                (130, 9, 0, 128, 32), (9, 130, 0, 16, 256), (260, 4, 0, 1024, 4), (5, 300, 0, 4, 1024),
                                                (1, 7, 0, 8, 8), (7, 1, 0, 8, 8), (2, 2, 0, 4, 4)])
def test_random_streams(ectx, ht, oracle, w, h, levels, cbw, cbh):
    ok, nb = HP.check_random_streams(ectx, ht, oracle, w, h, levels, cbw, cbh, seed=w * 131 + h)
    assert nb > 0


def test_error_codes(ectx, ht, oracle):
    HP.check_error_codes(ectx, ht, oracle)


@pytest.mark.parametrize("w,h,levels,cbw,cbh,bits,density,comps,rev", [
    (64, 64, 0, 64, 64, 12, 0.7, 1, True), (96, 80, 2, 32, 32, 8, 0.3, 1, True), (75, 61, 2, 64, 64, 16, 0.9, 1, True),
    (128, 32, 1, 128, 32, 10, 0.5, 1, False), (40, 40, 1, 16, 16, 8, 0.6, 3, True), (33, 130, 1, 8, 512, 12, 1.0, 1, True),
    (150, 10, 0, 1024, 4, 9, 0.8, 1, True), (48, 48, 2, 64, 64, 8, 0.05, 3, False),
])
def test_generated_streams(ectx, ht, oracle, w, h, levels, cbw, cbh, bits, density, comps, rev):
    HP.check_generated(ectx, ht, oracle, w, h, levels, cbw, cbh, bits, density, seed=w + 7 * h, components=comps, reversible=rev)


def test_generated_two_frames(ectx, ht, oracle):
    HP.check_generated(ectx, ht, oracle, 72, 56, 2, 32, 32, 12, 0.6, seed=3, nframes=2)


@pytest.mark.parametrize("w,h,comps,bits,levels,cbw,cbh,rev", [
    (64, 64, 1, 8, 0, 64, 64, True), (96, 80, 1, 12, 2, 32, 32, True), (75, 61, 1, 16, 2, 64, 64, True), (128, 32, 1, 10, 1, 128, 32, False),
    (40, 40, 3, 8, 1, 16, 16, True), (33, 130, 1, 12, 1, 8, 512, True), (150, 10, 1, 9, 0, 1024, 4, True), (48, 48, 3, 8, 2, 64, 64, False),
    (1, 1, 1, 8, 0, 4, 4, True), (2, 7, 1, 8, 0, 4, 4, True),
])
def test_encode(ectx, ht, oracle, w, h, comps, bits, levels, cbw, cbh, rev):
    HP.check_encode(ectx, ht, oracle, w, h, comps, bits, levels, cbw, cbh, seed=w * 3 + h, reversible=rev)


def test_encode_frames_tiles_and_tight_kmax(ectx, ht, oracle):
    HP.check_encode(ectx, ht, oracle, 72, 56, 1, 12, 2, 32, 32, seed=4, nframes=3)
    HP.check_encode(ectx, ht, oracle, 70, 50, 3, 8, 2, 32, 32, seed=5, tile=(32, 32))
    # Kmax below the data's magnitudes: the reference shifts bits out of the 32-bit word; the same bytes are expected
    HP.check_encode(ectx, ht, oracle, 64, 48, 1, 12, 1, 64, 64, seed=6, base=6)


def test_encode_sharded_over_the_devices_of_a_context(ht, oracle, monkeypatch):
    """j2k_forward_ht cuts the frames into contiguous blocks, one per device of the context (here: the emulated device twice),
    each with its own lagged sub-batch pipeline; sub-batches of one frame so that every device runs several."""
    import emu_lib
    import j2kb200
    monkeypatch.setenv("J2K_HT_SUBBATCH_MSAMPLES", "0")   # -> one frame per sub-batch
    with j2kb200.Context(devices=[0, 0], lib_path=emu_lib.build()) as c2:
        assert c2.device_count == 2
        HP.check_encode(c2, ht, oracle, 72, 56, 1, 12, 2, 32, 32, seed=4, nframes=7, ordered=False)
        HP.check_encode(c2, ht, oracle, 40, 40, 3, 8, 1, 16, 16, seed=9, nframes=2, reversible=False, ordered=False)
        HP.check_encode(c2, ht, oracle, 64, 48, 1, 12, 1, 64, 64, seed=6, nframes=1, ordered=False)   # fewer frames than devices


def test_encoder_table_matches_the_oracle(ectx, ht):
    import numpy as np
    for which in (0, 1):
        t = np.zeros(2048, np.uint16)
        assert ectx.lib.j2k_ht_enc_table(which, t.ctypes.data) == 2048
        assert np.array_equal(t, ht.enc_table(which))


@pytest.mark.parametrize("name,kind", [("mono_u8_127x129", "fo_htj2k_lossless"), ("mono_s16_128x128", "fo_htj2k_lossless_rpcl"),
                                       ("mono_u16_128x128", "fo_htj2k_lossless"), ("rgb_u8_127x129", "fo_htj2k_lossless_rpcl")])
def test_encode_reproduces_the_openjph_fixture_blocks(ectx, oracle, name, kind):
    HP.check_fixture_encode(ectx, oracle, name, kind)
