"""Oracle (literal Go restatement, C) == mirrored-extension formulation (numpy), bit for bit.

This is the licence for the CUDA kernels to use mirrored indices instead of the
reference's border special cases (jpeg2000/wavelet/dwt53.go:27-234, dwt97.go:47-287).
"""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import np_mirror as M  # noqa: E402


@pytest.mark.parametrize("even", [True, False])
def test_53_1d_all_lengths(oracle, even):
    rng = np.random.default_rng(11)
    for n in range(1, 49):
        for _ in range(6):
            x = rng.integers(-40000, 40000, n).astype(np.int32)
            f_c = oracle.fwd53_1d(x, even)
            f_m = M.fwd53_1d(x, even)
            assert np.array_equal(f_c, f_m), (n, even)
            assert np.array_equal(oracle.inv53_1d(f_c, even), x), (n, even)
            assert np.array_equal(M.inv53_1d(f_c, even), x), (n, even)
            # inverse formulations agree on arbitrary (non-image) coefficient vectors too
            y = rng.integers(-40000, 40000, n).astype(np.int32)
            assert np.array_equal(oracle.inv53_1d(y, even), M.inv53_1d(y, even)), (n, even)


@pytest.mark.parametrize("even", [True, False])
def test_97_1d_all_lengths_bitwise(oracle, even):
    rng = np.random.default_rng(12)
    for n in range(1, 49):
        for _ in range(4):
            x = (rng.standard_normal(n) * 1000).astype(np.float32)
            f_c = oracle.fwd97_1d(x, even)
            f_m = M.fwd97_1d(x, even)
            assert np.array_equal(f_c.view(np.uint32), f_m.view(np.uint32)), (n, even)
            i_c = oracle.inv97_1d(x, even)
            i_m = M.inv97_1d(x, even)
            assert np.array_equal(i_c.view(np.uint32), i_m.view(np.uint32)), (n, even)


@pytest.mark.parametrize("w,h,levels,x0,y0", [
    (16, 16, 2, 0, 0), (17, 19, 3, 0, 0), (33, 17, 3, 1, 0), (20, 9, 4, 0, 1), (64, 48, 3, 1, 2),
    (7, 1, 2, 0, 0), (1, 9, 3, 1, 1), (5, 5, 6, 3, 3), (2, 2, 3, 1, 1), (1, 1, 2, 1, 0), (13, 2, 5, 2, 7),
])
def test_multilevel_2d(oracle, w, h, levels, x0, y0):
    rng = np.random.default_rng(w * 100 + h)
    a = rng.integers(-2000, 2000, (h, w)).astype(np.int32)
    f_c = oracle.fwd53(a, levels, x0, y0)
    assert np.array_equal(f_c, M.fwd_multilevel(a, levels, x0, y0, "53"))
    assert np.array_equal(oracle.inv53(f_c, levels, x0, y0), a)
    assert np.array_equal(M.inv_multilevel(f_c, levels, x0, y0, "53"), a)
    b = a.astype(np.float32)
    g_c = oracle.fwd97(b, levels, x0, y0)
    assert np.array_equal(g_c.view(np.uint32), M.fwd_multilevel(b, levels, x0, y0, "97").view(np.uint32))
    r_c = oracle.inv97(b, levels, x0, y0)
    assert np.array_equal(r_c.view(np.uint32), M.inv_multilevel(b, levels, x0, y0, "97").view(np.uint32))
