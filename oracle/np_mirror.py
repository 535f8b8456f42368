"""numpy mirror of the oracle, in the formulation the CUDA kernels use.

TEST INFRASTRUCTURE ONLY (see oracle/j2k_oracle.c).  Where j2k_oracle.c follows the
Go text loop-for-loop (border special cases, scratch lines), this file restates
the same arithmetic as *uniform lifting over a whole-sample-symmetric extension*:

    x_ext[i] = x[mirror(i, n)],  low-pass samples at i == p (mod 2), p = origin parity

which is what a GPU thread computes (no border branches).  tests/test_oracle_mirror.py
checks that both formulations agree bit-for-bit for every length 1..48 and both
parities, which is the licence for the kernels to use mirrored indices
(reference: jpeg2000/wavelet/dwt53.go:27-234, jpeg2000/wavelet/dwt97.go:47-287).
"""
from __future__ import annotations

import numpy as np

F = np.float32
ALPHA = F(np.float64(-1.586134342))
BETA = F(np.float64(-0.052980118))
GAMMA = F(np.float64(0.882911075))
DELTA = F(np.float64(0.443506852))
K = F(np.float64(1.230174105))
INVK = F(np.float64(0.812893066))
TWO_INVK = F(np.float64(1.625732422))


def mirror(i, n):
    """Whole-sample symmetric reflection of integer index array i into [0, n)."""
    i = np.asarray(i, dtype=np.int64)
    if n == 1:
        return np.zeros_like(i)
    period = 2 * (n - 1)
    i = np.mod(i, period)
    return np.where(i < n, i, period - i)


def split(n, even):
    return (n + 1) // 2 if even else n // 2


def _ext(x, halo):
    n = x.shape[0]
    idx = mirror(np.arange(-halo, n + halo), n)
    return x[idx]


def fwd53_1d(x, even=True):
    """ISO 15444-1 F.3 reversible lifting on the mirrored extension (dwt53.go:27-103)."""
    x = np.asarray(x, dtype=np.int32)
    n = x.shape[0]
    p = 0 if even else 1
    if n == 1:
        return x.copy() if even else (x * 2).astype(np.int32)
    H = 4
    e = _ext(x, H).astype(np.int64)  # position i <-> e[i + H]
    pos = np.arange(-H, n + H)
    is_low = ((pos - p) % 2) == 0
    d = e.copy()
    hi = np.where(~is_low)[0]
    hi = hi[(hi >= 1) & (hi < len(e) - 1)]
    d[hi] = e[hi] - ((e[hi - 1] + e[hi + 1]) >> 1)
    s = d.copy()
    lo = np.where(is_low)[0]
    lo = lo[(lo >= 2) & (lo < len(e) - 2)]
    s[lo] = d[lo] + ((d[lo - 1] + d[lo + 1] + 2) >> 2)
    core = s[H:H + n]
    low = core[p::2]
    high = core[1 - p::2]
    return np.concatenate([low, high]).astype(np.int32)


def inv53_1d(y, even=True):
    """Inverse of fwd53_1d on the mirrored extension (dwt53.go:123-234)."""
    y = np.asarray(y, dtype=np.int32)
    n = y.shape[0]
    p = 0 if even else 1
    if n == 1:
        if even:
            return y.copy()
        return np.array([int(np.trunc(int(y[0]) / 2))], dtype=np.int32)  # Go `/` truncates
    sn = split(n, even)
    inter = np.empty(n, dtype=np.int64)
    inter[p::2] = y[:sn]
    inter[1 - p::2] = y[sn:]
    H = 4
    e = _ext(inter, H)
    pos = np.arange(-H, n + H)
    is_low = ((pos - p) % 2) == 0
    s = e.copy()
    lo = np.where(is_low)[0]
    lo = lo[(lo >= 1) & (lo < len(e) - 1)]
    s[lo] = e[lo] - ((e[lo - 1] + e[lo + 1] + 2) >> 2)
    x = s.copy()
    hi = np.where(~is_low)[0]
    hi = hi[(hi >= 2) & (hi < len(e) - 2)]
    x[hi] = s[hi] + ((s[hi - 1] + s[hi + 1]) >> 1)
    return x[H:H + n].astype(np.int32)


def _lift(e, sel, c):
    """e[sel] = e[sel] + (e[sel-1] + e[sel+1]) * c with three separate float32 roundings."""
    t = (e[sel - 1] + e[sel + 1]).astype(F)
    t = (t * c).astype(F)
    e[sel] = (e[sel] + t).astype(F)


def fwd97_1d(x, even=True):
    """OpenJPEG float32 9/7 analysis on the mirrored extension (dwt97.go:47-160)."""
    x = np.asarray(x, dtype=F)
    n = x.shape[0]
    if n <= 1:
        return x.copy()
    p = 0 if even else 1
    H = 6
    e = _ext(x, H).copy()
    pos = np.arange(-H, n + H)
    is_low = ((pos - p) % 2) == 0
    idx = np.arange(len(e))
    for k, c in enumerate((ALPHA, BETA, GAMMA, DELTA)):
        want_low = (k % 2) == 1
        sel = idx[(is_low == want_low) & (idx >= k + 1) & (idx < len(e) - k - 1)]
        _lift(e, sel, c)
    core = e[H:H + n]
    low = (core[p::2] * INVK).astype(F)
    high = (core[1 - p::2] * K).astype(F)
    return np.concatenate([low, high]).astype(F)


def inv97_1d(y, even=True):
    """OpenJPEG float32 9/7 synthesis on the mirrored extension (dwt97.go:192-287)."""
    y = np.asarray(y, dtype=F)
    n = y.shape[0]
    if n <= 1:
        return y.copy()
    p = 0 if even else 1
    sn = split(n, even)
    inter = np.empty(n, dtype=F)
    inter[p::2] = (y[:sn] * K).astype(F)
    inter[1 - p::2] = (y[sn:] * TWO_INVK).astype(F)
    H = 6
    e = _ext(inter, H).copy()
    pos = np.arange(-H, n + H)
    is_low = ((pos - p) % 2) == 0
    idx = np.arange(len(e))
    for k, c in enumerate((F(-DELTA), F(-GAMMA), F(-BETA), F(-ALPHA))):
        want_low = (k % 2) == 0
        sel = idx[(is_low == want_low) & (idx >= k + 1) & (idx < len(e) - k - 1)]
        _lift(e, sel, c)
    return e[H:H + n].astype(F)


def _apply_cols(a, fn, even):
    out = a.copy()
    for x in range(a.shape[1]):
        out[:, x] = fn(a[:, x], even)
    return out


def _apply_rows(a, fn, even):
    out = a.copy()
    for y in range(a.shape[0]):
        out[y, :] = fn(a[y, :], even)
    return out


def _windows(w, h, levels, x0, y0):
    win = [(w, h, x0, y0)]
    for _ in range(levels):
        cw, ch, cx, cy = win[-1]
        win.append((split(cw, cx % 2 == 0), split(ch, cy % 2 == 0), (cx + 1) >> 1, (cy + 1) >> 1))
    return win


def fwd_multilevel(a, levels, x0=0, y0=0, kind="53"):
    """Per level: all columns then all rows (dwt53.go:259-301,365-394; dwt97.go:290-322,388-407)."""
    f1 = fwd53_1d if kind == "53" else fwd97_1d
    a = np.array(a, dtype=np.int32 if kind == "53" else F)
    h, w = a.shape
    cw, ch, cx, cy = w, h, x0, y0
    for _ in range(levels):
        if cw <= 1 and ch <= 1:
            break
        win = a[:ch, :cw]
        if ch > 1:
            win = _apply_cols(win, f1, cy % 2 == 0)
        if cw > 1:
            win = _apply_rows(win, f1, cx % 2 == 0)
        a[:ch, :cw] = win
        cw, ch, cx, cy = split(cw, cx % 2 == 0), split(ch, cy % 2 == 0), (cx + 1) >> 1, (cy + 1) >> 1
    return a


def inv_multilevel(a, levels, x0=0, y0=0, kind="53"):
    """Per level, coarsest first: all rows then all columns (dwt53.go:313-355,404-434; dwt97.go:355-385,425-451)."""
    f1 = inv53_1d if kind == "53" else inv97_1d
    a = np.array(a, dtype=np.int32 if kind == "53" else F)
    h, w = a.shape
    win = _windows(w, h, levels, x0, y0)
    for lvl in range(levels - 1, -1, -1):
        cw, ch, cx, cy = win[lvl]
        if cw <= 1 and ch <= 1:
            continue
        sub = a[:ch, :cw]
        if cw > 1:
            sub = _apply_rows(sub, f1, cx % 2 == 0)
        if ch > 1:
            sub = _apply_cols(sub, f1, cy % 2 == 0)
        a[:ch, :cw] = sub
    return a


def t1_emulate(q, htj2k=False):
    """What classic EBCOT hands back when every coding pass is kept: the integer magnitude
    m = |q| >> 6 (6 fractional bits, jpeg2000/t1/encoder.go:203) reconstructed with one
    half-bit, v = sign * (2m + 1) (jpeg2000/t1/decoder.go:630-647); zero stays zero."""
    q = np.asarray(q, dtype=np.int64)
    m = np.abs(q) >> (0 if htj2k else 6)
    v = np.where(m > 0, np.sign(q) * (2 * m + 1), 0)
    return v.astype(np.int32)
