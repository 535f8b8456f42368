"""Caller-owned PAGEABLE host buffers (a Go []byte from PixelData.GetFrame, a plain numpy array) through the synchronous
entry points: the library moves them through its per-device pinned staging ring (SURVEY 8b "Ownership", `stage_up` /
`stage_drain` in j2k_b200.cu) and must produce exactly what the pinned path produces; the ticketed calls refuse them.

CPU: the emulator build treats every host buffer as pageable under J2K_EMU_PAGEABLE=1 (a subprocess, because the switches are
read once per process), with 4 KB chunks and 2-frame sub-batches so that the chunk ring wraps and several sub-batches overlap.
GPU: numpy arrays against j2k_acquire_buffer memory at sizes that need many 4 MB chunks."""
import os
import subprocess
import sys

import numpy as np
import pytest

import parity_cases as PC
from j2kb200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

EMU_SCRIPT = r"""
import sys
sys.path.insert(0, %(pkg)r); sys.path.insert(0, %(tests)r)
import numpy as np
import emu_lib, j2kb200, oracle_lib, parity_cases as PC
from j2kb200 import abi
orc = oracle_lib.Oracle()
rng = np.random.default_rng(5)
with j2kb200.Context(lib_path=emu_lib.build()) as ctx:
    # lossless series: 7 frames, 2 per sub-batch -> 4 sub-batches, each upload 3 chunks
    w, h, n = 64, 48, 7
    fp, ip = PC.fwd_inv_params(w, h, 1, 16, True, 3, True, orc)
    frames = np.stack([PC.raw_bytes(PC.synth(rng, h, w, 1, 16, True, "noise")) for _ in range(n)])
    co = ctx.forward_batch(fp, frames)
    for f in range(n):
        assert np.array_equal(co[f], orc.forward(fp, frames[f])), f
    back = ctx.inverse_batch(ip, co)
    assert np.array_equal(back, frames)
    # strided frames (a frame stride larger than a frame) in both directions
    wide = np.zeros((n, frames.shape[1] + 40), np.uint8); wide[:, :frames.shape[1]] = frames
    co2 = ctx.forward_batch(fp, wide[:, :frames.shape[1]])
    assert np.array_equal(co2, co)
    # lossy RGB with planes (a second pageable output) and the code-block interface (numbps: a third)
    PC.check_pipeline(ctx, orc, 48, 40, 3, 8, False, 2, False, kind="noise", seed=2)
    PC.check_blocks(ctx, orc, 64, 48, 1, 12, 2, True, cb=(16, 16))
    # the ticketed calls refuse pageable memory (which also proves that this process took the staging path above)
    try:
        ctx.submit_forward(fp, frames, co)
    except j2kb200.J2KError as e:
        assert "pinned" in str(e), str(e)
    else:
        raise AssertionError("submit_forward accepted a pageable buffer")
print("pageable emulator ok")
"""


def test_pageable_buffers_through_the_staging_ring_emulator():
    env = dict(os.environ, J2K_EMU_PAGEABLE="1", J2K_STAGE_CHUNK_KB="4", J2K_SUBBATCH_SAMPLES=str(2 * 64 * 48), J2K_STAGE_THREADS="3")
    script = EMU_SCRIPT % {"pkg": os.path.join(ROOT, "go-dicom-codec_b200"), "tests": os.path.join(ROOT, "tests")}
    out = subprocess.run([sys.executable, "-c", script], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "pageable emulator ok" in out.stdout


@pytest.mark.gpu
def test_pageable_sync_calls(ctx, oracle):
    """forward_batch / inverse_batch with plain numpy arrays == the same calls with pinned buffers, bit for bit, at sizes with
    many chunks per sub-batch and several sub-batches; the oracle checks one frame of each."""
    rng = np.random.default_rng(9)
    for (w, h, c, bits, L, rev, n) in ((2048, 2048, 1, 12, 5, False, 9), (1024, 768, 3, 8, 4, False, 12), (512, 512, 1, 16, 5, True, 70)):
        fp, ip = PC.fwd_inv_params(w, h, c, bits, False, L, rev, oracle)
        one = [PC.raw_bytes(PC.synth(rng, h, w, c, bits, False, "smooth" if k else "noise")) for k in range(3)]
        frames = np.stack([one[k % 3] for k in range(n)])            # pageable
        nc = w * h * c
        p_in = ctx.pinned(frames.size).reshape(frames.shape); p_in[:] = frames
        p_co = ctx.pinned(n * nc * 4, np.int32).reshape(n, nc)
        ctx.forward_batch(fp, p_in, p_co)
        co = ctx.forward_batch(fp, frames)                            # pageable in, pageable out
        assert np.array_equal(co, p_co)
        assert np.array_equal(co[1], oracle.forward(fp, frames[1]))
        back_in = co if rev else np.stack([PC.M.t1_emulate(co[f], False) for f in range(n)])
        p_bi = ctx.pinned(back_in.size * 4, np.int32).reshape(back_in.shape); p_bi[:] = back_in
        p_px = ctx.pinned(frames.size).reshape(frames.shape)
        ctx.inverse_batch(ip, p_bi, p_px)
        px = ctx.inverse_batch(ip, np.ascontiguousarray(back_in))
        assert np.array_equal(px, p_px)
        assert np.array_equal(px[2], oracle.inverse(ip, back_in[2]))
        if rev:
            assert np.array_equal(px, frames)
        for b in (p_in, p_co, p_bi, p_px):
            ctx.release(b)


@pytest.mark.gpu
def test_ticketed_calls_refuse_pageable_buffers(ctx, oracle):
    import j2kb200
    fp, _ = PC.fwd_inv_params(256, 256, 1, 12, False, 3, True, oracle)
    frames = np.zeros((2, 256 * 256 * 2), np.uint8)
    out = np.zeros((2, 256 * 256), np.int32)
    with pytest.raises(j2kb200.J2KError) as e:
        ctx.submit_forward(fp, frames, out)
    assert "pinned" in str(e.value)
