"""Full-size parity of the BASELINE configs against the oracle (round-1 VERDICT "test thinness at full size"):
C3 at 2048 x 2048 in both transforms and both directions, the C2 4096 x 4096 inverse, one full tile row of C5 (32 tiles of
1024 x 1024 RGB, ICT + 9/7, 7 levels) inverse, and two tickets in flight with every frame compared.  All through the C ABI;
the oracle (C port, one thread) needs a few seconds per case."""
import numpy as np
import pytest

import parity_cases as PC
from j2kb200 import abi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rev", [True, False])
@pytest.mark.parametrize("kind", ["smooth", "noise"])
def test_c3_full_size_both_directions(ctx, oracle, rev, kind):
    """C3: 2048 x 2048 x 3 8-bit, (ii) RCT + 5/3 and (i) ICT + 9/7 with the OpenJPEG default steps, L = 5."""
    w = h = 2048
    rng = np.random.default_rng(33 + rev)
    img = PC.synth(rng, h, w, 3, 8, False, kind)
    fp, ip = PC.fwd_inv_params(w, h, 3, 8, False, 5, rev, oracle)
    raw = PC.raw_bytes(img)
    co = ctx.forward(fp, raw)
    want = oracle.forward(fp, raw)
    assert np.count_nonzero(co != want) == 0
    back_in = co if rev else PC.M.t1_emulate(co, False)
    px = ctx.inverse(ip, back_in)                      # packed pixels only: the float32 ICT fast path is eligible
    assert np.array_equal(px, oracle.inverse(ip, back_in))
    px2, planes = ctx.inverse(ip, back_in, want_planes=True)   # with GetImageData planes: the float64 path
    opx, oplanes = oracle.inverse(ip, back_in, want_planes=True)
    assert np.array_equal(px2, opx) and np.array_equal(planes, oplanes)
    if rev:
        assert np.array_equal(px, raw.reshape(-1))


def test_c2_full_size_inverse(ctx, oracle):
    """C2: 4096 x 4096 12-bit, 9/7 L = 6: the inverse of the T1-reconstructed coefficients of a full frame."""
    rng = np.random.default_rng(22)
    img = PC.synth(rng, 4096, 4096, 1, 12)
    fp, ip = PC.fwd_inv_params(4096, 4096, 1, 12, False, 6, False, oracle)
    co = ctx.forward(fp, PC.raw_bytes(img))
    back_in = PC.M.t1_emulate(co, False)
    px = ctx.inverse(ip, back_in)
    assert np.array_equal(px, oracle.inverse(ip, back_in))


def test_c5_full_tile_row_inverse(ctx, oracle):
    """C5: one full tile row of the slide (32768 x 1024 RGB = 32 tiles of 1024 x 1024), ICT + 9/7, 7 levels: forward of the row
    vs the oracle on three tiles, inverse of the WHOLE row vs the oracle (compared, not bounded)."""
    W, H = 32768, 1024
    pool = [PC.synth(np.random.default_rng(5100 + k), 1024, 1024, 3, 8, False, "noise" if k == 3 else "smooth") for k in range(4)]
    img = np.empty((H, W, 3), np.uint8)
    for t in range(32):
        img[:, t * 1024:(t + 1) * 1024] = np.roll(pool[t % 4], 7 * t, axis=0) // 2 + 3 * t
    raw = img.reshape(-1)
    es, ds = PC.steps_for(oracle, 7, 8)
    fp = abi.fwd_params(W, H, 3, 8, False, 1024, 1024, 7, False, False, abi.MCT_ICT, es)
    ip = abi.inv_params(W, H, 3, 8, False, 1024, 1024, 7, False, False, abi.MCT_ICT, ds)
    co = ctx.forward(fp, raw)
    fpt = abi.fwd_params(1024, 1024, 3, 8, False, 0, 0, 7, False, False, abi.MCT_ICT, es)
    n_t = 3 * 1024 * 1024
    for t in (0, 13, 31):
        tile = np.ascontiguousarray(img[:, t * 1024:(t + 1) * 1024]).reshape(-1)
        assert np.array_equal(co[t * n_t:(t + 1) * n_t], oracle.forward(fpt, tile)), t
    back_in = PC.M.t1_emulate(co, False)
    px = ctx.inverse(ip, back_in)
    want = oracle.inverse(ip, back_in)
    assert np.array_equal(px, want)


def test_two_tickets_in_flight_every_frame_compared(ctx, oracle):
    """Ticketed calls: two forward jobs and then two inverse jobs in flight at once, different data in each, every frame of
    every job checked (forward: against the oracle; inverse: lossless identity and oracle)."""
    rng = np.random.default_rng(77)
    n, w, h = 6, 512, 384
    fp, ip = PC.fwd_inv_params(w, h, 1, 16, False, 5, True, oracle)
    fin = [ctx.pinned(n * w * h * 2).reshape(n, -1) for _ in range(2)]
    fco = [ctx.pinned(n * w * h * 4, np.int32).reshape(n, -1) for _ in range(2)]
    fpx = [ctx.pinned(n * w * h * 2).reshape(n, -1) for _ in range(2)]
    for k in range(2):
        fin[k][:] = rng.integers(0, 256, fin[k].shape, dtype=np.uint8)
    t0 = ctx.submit_forward(fp, fin[0], fco[0])
    t1 = ctx.submit_forward(fp, fin[1], fco[1])     # second job enqueued while the first is in flight
    ctx.wait(t1)                                     # waited for out of order on purpose
    ctx.wait(t0)
    for k in range(2):
        for f in range(n):
            assert np.array_equal(fco[k][f], oracle.forward(fp, fin[k][f])), (k, f)
    u0 = ctx.submit_inverse(ip, fco[0], fpx[0])
    u1 = ctx.submit_inverse(ip, fco[1], fpx[1])
    ctx.wait(u0)
    ctx.wait(u1)
    for k in range(2):
        assert np.array_equal(fpx[k], fin[k]), k
        assert np.array_equal(fpx[k][n - 1], oracle.inverse(ip, fco[k][n - 1]))
    for b in fin + fco + fpx:
        ctx.release(b)
