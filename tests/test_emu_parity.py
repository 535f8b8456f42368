"""Product sources (csrc/*.cu, *.cuh) compiled for the CPU emulator (tests/emu) vs the oracle.

Covers the host logic (plans, tile classes, offset tables, kernel selection) and the kernels' index
arithmetic and lifting order without a GPU.  Sizes are small: the emulator runs one lane at a time.
The `-m gpu` suite repeats these through the real sm_100a build at full sizes."""
import numpy as np
import pytest

import parity_cases as PC
from j2kb200 import abi


@pytest.fixture(scope="module")
def ectx():
    import emu_lib
    import j2kb200
    c = j2kb200.Context(lib_path=emu_lib.build())
    yield c
    c.close()


@pytest.mark.parametrize("w,h,levels,x0,y0", [
    (16, 16, 1, 0, 0), (64, 64, 3, 0, 0), (17, 19, 3, 0, 0), (33, 17, 2, 1, 0), (20, 9, 4, 0, 1), (64, 48, 3, 1, 2),
    (7, 1, 2, 0, 0), (1, 9, 3, 1, 1), (5, 5, 6, 3, 3), (2, 2, 3, 1, 1), (1, 1, 2, 1, 0), (13, 2, 5, 2, 7),
    (130, 70, 5, 0, 0), (257, 3, 2, 0, 0), (3, 300, 3, 1, 0),
])
def test_wavelet_api(ectx, oracle, w, h, levels, x0, y0):
    PC.check_wavelet_api(ectx, oracle, w, h, levels, x0, y0, seed=w * 1000 + h)


@pytest.mark.parametrize("w,h,c,bits,signed,L,rev", [
    (64, 64, 1, 8, False, 3, True), (64, 64, 1, 16, False, 5, True), (61, 47, 1, 16, True, 3, True),
    (64, 64, 1, 12, False, 3, False), (67, 53, 1, 8, False, 2, False), (160, 40, 1, 16, False, 4, False),
    (48, 40, 3, 8, False, 3, True), (48, 40, 3, 8, False, 3, False), (37, 29, 3, 16, False, 2, False),
    (256, 8, 1, 16, False, 2, True), (256, 8, 1, 12, True, 2, False), (40, 40, 2, 8, False, 2, True),
    (24, 24, 1, 8, False, 0, True), (24, 24, 1, 8, False, 0, False), (1, 1, 1, 8, False, 2, False),
])
def test_pipeline(ectx, oracle, w, h, c, bits, signed, L, rev):
    PC.check_pipeline(ectx, oracle, w, h, c, bits, signed, L, rev)


@pytest.mark.parametrize("w,h,c,tile,L,rev", [
    (96, 64, 1, (32, 32), 2, True), (100, 70, 1, (48, 32), 3, False), (70, 50, 3, (32, 32), 2, False),
    (65, 33, 3, (32, 32), 3, True), (50, 50, 1, (33, 17), 3, False), (33, 33, 1, (32, 32), 3, False),
])
def test_tiles(ectx, oracle, w, h, c, tile, L, rev):
    PC.check_pipeline(ectx, oracle, w, h, c, 8, False, L, rev, tile=tile)


def test_htj2k_and_fused_t1_shift(ectx, oracle):
    PC.check_pipeline(ectx, oracle, 40, 36, 1, 16, False, 3, True, fuse=True)
    PC.check_pipeline(ectx, oracle, 40, 36, 1, 16, False, 3, True, htj2k=True, fuse=True)
    PC.check_pipeline(ectx, oracle, 40, 36, 1, 8, False, 3, False, htj2k=True)
    PC.check_pipeline(ectx, oracle, 40, 36, 3, 8, False, 2, False, steps_kind="quality")


@pytest.mark.parametrize("w,h,c,bits,signed,L,rev", [
    (512, 12, 1, 16, False, 2, False), (512, 12, 1, 12, True, 2, True), (384, 10, 1, 8, False, 2, False),
    (384, 10, 1, 8, True, 2, True), (136, 12, 3, 8, False, 2, False), (136, 12, 3, 16, False, 2, True),
    (520, 9, 1, 16, False, 3, False), (128, 16, 3, 8, False, 2, False), (128, 16, 3, 8, False, 2, True), (128, 16, 3, 8, True, 2, False),
    (256, 20, 3, 7, False, 3, False),
    (128, 16, 3, 16, False, 2, False), (128, 16, 3, 12, False, 2, True), (256, 24, 1, 8, False, 2, False),
])
def test_fast_path_shapes(ectx, oracle, w, h, c, bits, signed, L, rev, capfd):
    """Wide, aligned frames: level 1 (and the int32/float32 level 2) run on the fast-path kernel."""
    import os
    os.environ["J2K_B200_TRACE"] = "1"
    PC.check_pipeline(ectx, oracle, w, h, c, bits, signed, L, rev, seed=w)


@pytest.mark.parametrize("case", ["int_matrix", "q13_matrix", "q13_4comp", "bindings"])
@pytest.mark.parametrize("rev", [True, False])
def test_custom_mct(ectx, oracle, case, rev):
    PC.check_custom_mct(ectx, oracle, 40, 24, 8, 2, rev, case)


def test_custom_mct_tiled(ectx, oracle):
    PC.check_custom_mct(ectx, oracle, 50, 40, 12, 2, True, "bindings", tile=(32, 32))


@pytest.mark.parametrize("w,h,c,bits,signed,L,rev", [
    (40, 24, 1, 12, True, 2, True), (40, 24, 3, 8, False, 2, False), (136, 12, 1, 16, False, 2, False), (33, 17, 3, 10, True, 3, True),
])
def test_planar_entry(ectx, oracle, w, h, c, bits, signed, L, rev):
    PC.check_planar(ectx, oracle, w, h, c, bits, signed, L, rev)


def test_package_api(ectx, oracle):
    PC.check_package_api(ectx, oracle, n=5003)


@pytest.mark.parametrize("w,h,c,bits,L,rev,tile,cb", [
    (64, 48, 1, 8, 2, True, (0, 0), (16, 16)), (70, 50, 1, 12, 3, False, (0, 0), (16, 8)), (48, 40, 3, 8, 2, True, (32, 32), (8, 8)),
    (33, 17, 1, 16, 3, True, (0, 0), (4, 4)), (40, 24, 3, 8, 2, False, (0, 0), (64, 64)), (24, 24, 1, 8, 0, True, (0, 0), (8, 8)),
])
def test_code_block_interface(ectx, oracle, w, h, c, bits, L, rev, tile, cb):
    PC.check_blocks(ectx, oracle, w, h, c, bits, L, rev, tile=tile, cb=cb)


def test_code_block_interface_htj2k(ectx, oracle):
    PC.check_blocks(ectx, oracle, 40, 36, 1, 12, 2, True, cb=(16, 16), htj2k=True)
    PC.check_blocks(ectx, oracle, 40, 36, 1, 12, 2, False, cb=(16, 16), htj2k=True)


@pytest.mark.parametrize("w,h,c,bits,signed,L,rev", [
    (520, 24, 1, 16, False, 4, False), (520, 24, 1, 12, False, 4, True), (144, 20, 3, 8, False, 3, False), (144, 20, 3, 8, False, 3, True),
    (272, 18, 1, 8, False, 5, False),
])
def test_hybrid_plans(ectx, oracle, w, h, c, bits, signed, L, rev, capfd):
    """Widths that are a multiple of 8 only for the first levels: the persistent launch takes those, the per-level kernels
    the rest (forward: behind it, inverse: in front of it)."""
    import os
    os.environ["J2K_B200_TRACE"] = "1"
    PC.check_pipeline(ectx, oracle, w, h, c, bits, signed, L, rev, seed=w + L)
    err = capfd.readouterr().err
    if "[j2k]" in err:  # the trace is latched at first use; when it is on, the persistent launch must appear (since round 2 the
        assert " ring " in err  # general-alignment variant may take every level: per-level launches behind it are optional)


def test_package_api_x1(ectx, oracle):
    PC.check_package_api_x1(ectx, oracle)


@pytest.mark.parametrize("k,case", list(enumerate(PC.random_geometry_cases(48, 20261018, 160, 40))))
def test_random_geometry_sweep(ectx, oracle, k, case):
    PC.check_random_case(ectx, oracle, case, 9000 + k)


@pytest.mark.parametrize("w,h,c,bits,L,rev,xo,yo,tile,xto,yto", [
    (64, 48, 1, 8, 3, True, 1, 0, (0, 0), 0, 0), (64, 48, 1, 12, 3, False, 0, 1, (0, 0), 0, 0), (70, 50, 3, 8, 2, True, 1, 2, (32, 32), 1, 1),
    (70, 50, 3, 8, 2, False, 3, 5, (32, 32), 2, 3), (33, 17, 1, 16, 4, True, 7, 7, (16, 16), 0, 0), (40, 40, 2, 8, 2, False, 5, 2, (0, 0), 0, 0),
])
def test_inverse_with_image_and_tile_offsets(ectx, oracle, w, h, c, bits, L, rev, xo, yo, tile, xto, yto):
    PC.check_inverse_with_offsets(ectx, oracle, w, h, c, bits, L, rev, xo, yo, tile, xto, yto)


@pytest.mark.parametrize("w,h,c,bits,L,rev,nframes,group_ks,lag", [
    (64, 32, 1, 16, 3, True, 7, 2, 2), (64, 32, 1, 12, 3, False, 9, 4, 3), (48, 32, 3, 8, 2, False, 5, 4, 1), (48, 32, 3, 8, 2, True, 6, 4, 2),
    (64, 32, 1, 8, 2, False, 3, 2, 5),
])
def test_group_pipelined_job_order(ectx, oracle, w, h, c, bits, L, rev, nframes, group_ks, lag, capfd):
    """Batches long enough for the group-pipelined job order of the persistent launch (ring_schedule): slices of
    (level, item group) interleaved `lag` groups apart; a batch no longer than the lag keeps the level-major list."""
    PC.check_pipelined_order(ectx, oracle, w, h, c, bits, L, rev, nframes, group_ks, lag, capfd)


@pytest.mark.parametrize("w,h,c,bits,L,rev,shifts,tile,cb", [
    (64, 48, 1, 8, 2, True, [5], (0, 0), (16, 16)), (70, 50, 1, 12, 3, False, [9], (0, 0), (16, 8)), (48, 40, 3, 8, 2, True, [4, 0, 7], (32, 32), (8, 8)),
    (40, 24, 3, 8, 2, False, [0, 6, 31], (0, 0), (64, 64)), (33, 17, 1, 16, 3, True, [1], (0, 0), (4, 4)),
])
def test_code_block_interface_roi(ectx, oracle, w, h, c, bits, L, rev, shifts, tile, cb):
    PC.check_blocks_roi(ectx, oracle, w, h, c, bits, L, rev, shifts, tile=tile, cb=cb)


def test_roi_argument_errors(ectx):
    """j2k_inverse_blocks_roi validates the shifts like the encoder validates ROI.Shift (encoder.go:332-334: at most 255)."""
    import j2kb200
    ip = abi.inv_params(32, 32, 1, 8, False, num_levels=2, reversible=True)
    blocks = np.zeros((1, 32 * 32), np.int32)
    for bad in ([-1], [256]):
        with pytest.raises(j2kb200.J2KError) as e:
            ectx.inverse_blocks(ip, blocks, 16, 16, roi_maxshift=bad)
        assert e.value.code == abi.J2K_ERR_INVALID_ARG and "invalid ROI shift" in str(e.value)
    ectx.inverse_blocks(ip, blocks, 16, 16, roi_maxshift=[255])  # shift >= 31 zeroes the blocks: a valid call


@pytest.mark.parametrize("w,h,c,bits,L,rev,chunk", [(64, 88, 1, 12, 2, False, 16), (64, 72, 1, 16, 2, True, 12), (48, 56, 3, 8, 2, False, 16)])
def test_tall_chunks(ectx, oracle, w, h, c, bits, L, rev, chunk):
    PC.check_tall_chunks(ectx, oracle, w, h, c, bits, L, rev, chunk)


@pytest.mark.parametrize("w,h,c,bits,signed,L,rev", [
    (140, 20, 1, 16, False, 2, False), (150, 22, 1, 12, False, 3, False), (134, 18, 1, 8, False, 2, False), (131, 17, 1, 8, False, 3, False),
    (140, 20, 1, 16, True, 2, True), (150, 23, 1, 12, False, 3, True), (134, 18, 1, 8, True, 2, True), (133, 9, 1, 8, False, 2, True),
    (270, 12, 1, 16, False, 4, False), (271, 13, 1, 16, False, 4, True), (535, 10, 1, 16, False, 3, False), (300, 11, 1, 8, False, 5, True),
])
def test_general_alignment_ring_variant(ectx, oracle, w, h, c, bits, signed, L, rev, capfd):
    """Widths that are not a multiple of 8 (row pitch not a multiple of 16 bytes, band rows at odd offsets, odd LL windows):
    the persistent kernels' general-alignment variant (FwdRing / InvRing with UA = 1) takes them, level 1 included, in both
    directions - per-row staging phases, masked partial vectors at the right edge, the extra high-pass mirror of odd windows."""
    import os
    os.environ["J2K_B200_TRACE"] = "1"
    PC.check_pipeline(ectx, oracle, w, h, c, bits, signed, L, rev, kind="noise", seed=w)
    err = capfd.readouterr().err
    if "[j2k]" in err:
        assert "fwd ring" in err and "inv ring" in err


@pytest.mark.parametrize("w,h,c,bits,L,rev,tile,cb,masked", [(64, 48, 1, 12, 2, True, (0, 0), (16, 16), True), (70, 50, 1, 8, 3, False, (0, 0), (16, 8), False), (48, 40, 3, 8, 2, True, (32, 32), (8, 8), True)])
def test_code_block_interface_roi_general_scaling(ectx, oracle, w, h, c, bits, L, rev, tile, cb, masked):
    """SURVEY 8f rank 3, second half: inverse general scaling (RGN Srgn = 1) fused into the block scatter, whole-block and masked."""
    PC.check_blocks_roi_general(ectx, oracle, w, h, c, bits, L, rev, tile=tile, cb=cb, masked=masked)
    PC.check_blocks_roi_general(ectx, oracle, w, h, c, bits, L, rev, tile=tile, cb=cb, masked=masked, maxshift=[3] * c, seed=5)


@pytest.mark.parametrize("w,h,bits,L,nframes,tile,chunk", [
    (512, 44, 8, 3, 2, (0, 0), 0), (256, 40, 8, 2, 3, (0, 0), 8), (512, 24, 16, 3, 1, (0, 0), 0), (256, 64, 8, 2, 2, (128, 32), 0),
    (768, 50, 12, 2, 2, (0, 0), 8), (320, 37, 8, 2, 1, (0, 0), 0),
])
def test_one_producer_rgb97_forward(ectx, oracle, w, h, bits, L, nframes, tile, chunk, capfd):
    """fwd3w_kernel: ICT + 9/7 level 1 of raw RGB frames as one converting producer warp + three single-component consumers
    per CTA (several strips, chunks, frames and tile classes; odd heights; coarser levels as job triples behind it)."""
    PC.check_one_producer_forward(ectx, oracle, w, h, bits, L, nframes, tile, chunk, capfd)


def test_one_producer_rgb97_forward_default_policy(ectx, oracle, capfd):
    """A launch big enough for the plan builder to pick fwd3w_kernel by itself (the emulated device has one SM = four quads)."""
    PC.check_one_producer_forward(ectx, oracle, 256, 200, 8, 2, 8, capfd=capfd, arm_on="1")


def test_failed_device_is_removed_from_the_round_robin(oracle, monkeypatch):
    """SURVEY 5 "a failed GPU is removed from the round-robin": a device slot whose CUDA calls fail (J2K_FAULT_DEVICE injects
    that) is marked failed, the blocking call re-runs its frame block on the remaining slot and still returns every frame
    bit-exact, later calls shard over what is left, and a context with no device left reports an error instead of hanging."""
    PC.check_failed_device(oracle, monkeypatch, __import__("emu_lib").build())
