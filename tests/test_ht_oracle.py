"""The HTJ2K cleanup-pass block-decoder oracle (oracle/ht_oracle.c) against the reference's own interop contract:
the 14 OpenJPH codestreams of test-data/htj2k/interop decode to input.raw (jpeg2000/htj2k/interop_manifest_test.go:43-74).
CPU only."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "go-dicom-codec_b200"))

import ht_cases  # noqa: E402
import ht_oracle_lib  # noqa: E402
import oracle_lib  # noqa: E402
from j2kb200 import abi  # noqa: E402


@pytest.fixture(scope="module")
def ht():
    return ht_oracle_lib.HtOracle()


@pytest.fixture(scope="module")
def oracle():
    return oracle_lib.Oracle()


def decode_fixture(ht, oracle, name, kind):
    fx = ht_cases.load(name, kind, oracle.codeblock_layout)
    h = fx["header"]
    blocks, status = ht.decode_blocks(fx["stream"], fx["offsets"], fx["lengths"], fx["kmax"], fx["mmsb"], fx["widths"], fx["heights"],
                                      fx["out_offsets"], h.components * fx["plane_samples"])
    assert not status.any(), status
    co = np.concatenate([oracle.scatter_blocks(blocks[c * fx["plane_samples"]:(c + 1) * fx["plane_samples"]], h.width, h.height,
                                               h.num_levels, h.cbw, h.cbh).reshape(-1) for c in range(h.components)])
    ip = abi.inv_params(h.width, h.height, h.components, h.depth[0], h.signed[0], num_levels=h.num_levels, reversible=True, htj2k=True,
                        mct_mode=abi.MCT_RCT if h.mct else abi.MCT_NONE)
    return fx, oracle.inverse(ip, co)


@pytest.mark.parametrize("name,kind", ht_cases.fixtures())
def test_openjph_fixtures_decode_to_input_raw(ht, oracle, name, kind):
    fx, px = decode_fixture(ht, oracle, name, kind)
    h = fx["header"]
    assert h.cb_style & 0x40, "HT code-blocks"
    assert all(b.passes in (0, 1) for comp in fx["blocks"] for b in comp), "OpenJPH lossless emits the cleanup pass only"
    assert np.array_equal(px, fx["raw"])


def test_tables_match_the_generated_product_tables(ht):
    # tools/gen_ht_tables.py builds the packed tables the CUDA decoder indexes; the oracle builds its own at run time
    inc = open(os.path.join(os.path.dirname(HERE), "go-dicom-codec_b200", "csrc", "j2k_ht_tables.inc")).read()
    import re
    for which, cname in enumerate(("HT_VLC_TBL0", "HT_VLC_TBL1", "HT_UVLC_TBL0", "HT_UVLC_TBL1")):
        body = inc.split(cname + "[", 1)[1].split("{", 1)[1].split("}", 1)[0]
        vals = np.array([int(x, 16) for x in re.findall(r"0x[0-9A-Fa-f]+", body)], np.uint16)
        assert np.array_equal(vals, ht.table(which)), cname


def test_empty_and_invalid_blocks(ht):
    rc, out = ht.decode_block(b"", 8, 8, 10, 3)
    assert rc == 0 and not out.any()  # decoder.go:44-46
    rc, out = ht.decode_block(b"\x00\x00\x00\x00", 8, 8, 0, 3)
    assert rc == -1 and not out.any()  # decoder.go:48-50
    rc, out = ht.decode_block(b"\x00\x00\x00\x00", 8, 8, 10, 30)
    assert rc == -2  # openjph_cleanup_decoder.go:125-127
    rc, out = ht.decode_block(b"\x00\x00\x00\x00", 8, 8, 10, 3)
    assert rc == -2  # Scup = 0 < 2: decoder.go:63-65


def test_generator_round_trips_through_the_pinned_decoder(ht):
    """oracle/ht_oracle.c's stream generator (an HT cleanup encoder written against the decoder, not a restatement of the
    reference's encoder) is only trusted through this property: decode(encode(x)) == x for every block shape and magnitude."""
    rng = np.random.default_rng(0)
    for t in range(300):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        if t % 10 == 0:
            w, h = [(1024, 4), (4, 1024), (128, 32), (512, 8), (1, 1), (2, 1), (1, 2), (64, 64), (63, 64), (64, 63)][(t // 10) % 10]
        bits = int(rng.integers(1, 30 if t % 7 == 0 else 17))
        x = (rng.integers(-(1 << bits) + 1, 1 << bits, (h, w)) * (rng.random((h, w)) < rng.random())).astype(np.int32)
        mmsb = min(29, bits - 1 + int(rng.integers(0, 3)))
        data = ht.encode_block(x, mmsb)
        if not x.any():
            assert data == b""
            continue
        rc, y = ht.decode_block(data, w, h, mmsb + 1, mmsb)
        assert rc == 0 and np.array_equal(x, y), (t, w, h, bits)


@pytest.mark.parametrize("name,kind", ht_cases.fixtures())
def test_encoder_oracle_reproduces_the_openjph_block_bytes(ht, oracle, name, kind):
    """htj2k/go_byte_parity_test.go:11-44 at block level: HTEncoder.Encode of every code-block's coefficients gives exactly the
    bytes OpenJPH wrote into the codestream (and an empty block gives no bytes)."""
    fx = ht_cases.load(name, kind, oracle.codeblock_layout)
    for i in range(len(fx["offsets"])):
        o, n = int(fx["offsets"][i]), int(fx["lengths"][i])
        w, h, km, mm = int(fx["widths"][i]), int(fx["heights"][i]), int(fx["kmax"][i]), int(fx["mmsb"][i])
        seg = bytes(fx["stream"][o:o + n])
        rc, blk = ht.decode_block(seg, w, h, km, mm)
        assert rc == 0 and (n == 0 or mm == km - 1)
        assert ht.encode_ref(blk, km) == seg, (i, w, h)


def test_encoder_oracle_round_trips_and_rejects_bad_kmax(ht):
    rng = np.random.default_rng(3)
    for t in range(200):
        w, h = int(rng.integers(1, 65)), int(rng.integers(1, 65))
        bits = int(rng.integers(1, 17))
        x = (rng.integers(-(1 << bits) + 1, 1 << bits, (h, w)) * (rng.random((h, w)) < rng.random())).astype(np.int32)
        kmax = bits + int(rng.integers(0, 3))
        data = ht.encode_ref(x, kmax)
        if not x.any():
            assert data == b""
            continue
        rc, y = ht.decode_block(data, w, h, kmax, kmax - 1)
        assert rc == 0 and np.array_equal(x, y), (t, w, h, bits)
    assert ht.encode_ref(np.ones((4, 4), np.int32), 0) == -1 and ht.encode_ref(np.ones((4, 4), np.int32), 31) == -1   # :201-203
