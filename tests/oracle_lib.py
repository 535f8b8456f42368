"""ctypes wrapper around oracle/build/libj2k_oracle.so (the CPU checker).

Test infrastructure: imported only from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "build", "libj2k_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(ORACLE_DIR, "j2k_oracle.c")
    hdr = os.path.join(ROOT, "include", "j2k_b200.h")
    stale = (not os.path.exists(LIB)) or any(
        os.path.exists(f) and os.path.getmtime(f) > os.path.getmtime(LIB) for f in (src, hdr))
    if force or stale:
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.check_call([
            "gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-fvisibility=hidden",
            "-shared", "-o", LIB, src, "-lm", "-lpthread"])
    return LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    def __init__(self):
        import sys
        pkg = os.path.join(ROOT, "go-dicom-codec_b200")
        if pkg not in sys.path:
            sys.path.insert(0, pkg)
        from j2kb200 import abi
        self.abi = abi
        self.lib = C.CDLL(build())
        fwd, inv, bnd = C.c_int(), C.c_int(), C.c_int()
        self.lib.orc_abi_sizes(C.byref(fwd), C.byref(inv), C.byref(bnd))
        assert fwd.value == C.sizeof(abi.FwdParams), (fwd.value, C.sizeof(abi.FwdParams))
        assert inv.value == C.sizeof(abi.InvParams), (inv.value, C.sizeof(abi.InvParams))
        assert bnd.value == C.sizeof(abi.MctBinding)
        self.lib.orc_encode_quant_step.restype = C.c_uint16
        self.lib.orc_encode_quant_step.argtypes = [C.c_double, C.c_int]
        self.lib.orc_quantize_coefficients.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double]
        self.lib.orc_dequantize_coefficients.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_double]
        for n in ("orc_forward_batch", "orc_inverse_batch"):
            getattr(self.lib, n).restype = C.c_int

    # ---- wavelet
    def _ml(self, fn, data, levels, x0, y0, dtype):
        a = np.ascontiguousarray(data, dtype=dtype).copy()
        h, w = a.shape
        getattr(self.lib, fn)(_p(a), C.c_int(w), C.c_int(h), C.c_int(levels), C.c_int(x0), C.c_int(y0))
        return a

    def fwd53(self, data, levels, x0=0, y0=0):
        return self._ml("orc_fwd53_multilevel", data, levels, x0, y0, np.int32)

    def inv53(self, data, levels, x0=0, y0=0):
        return self._ml("orc_inv53_multilevel", data, levels, x0, y0, np.int32)

    def fwd97(self, data, levels, x0=0, y0=0):
        return self._ml("orc_fwd97_multilevel", data, levels, x0, y0, np.float32)

    def inv97(self, data, levels, x0=0, y0=0):
        return self._ml("orc_inv97_multilevel", data, levels, x0, y0, np.float32)

    def fwd97_f64(self, data, levels, x0=0, y0=0):
        return self._ml("orc_fwd97_multilevel_f64", data, levels, x0, y0, np.float64)

    def inv97_f64(self, data, levels, x0=0, y0=0):
        return self._ml("orc_inv97_multilevel_f64", data, levels, x0, y0, np.float64)

    def _1d(self, fn, data, even, dtype):
        a = np.ascontiguousarray(data, dtype=dtype).copy()
        getattr(self.lib, fn)(_p(a), C.c_int(a.size), C.c_int(1 if even else 0))
        return a

    def fwd53_1d(self, data, even=True):
        return self._1d("orc_fwd53_1d", data, even, np.int32)

    def inv53_1d(self, data, even=True):
        return self._1d("orc_inv53_1d", data, even, np.int32)

    def fwd97_1d(self, data, even=True):
        return self._1d("orc_fwd97_1d", data, even, np.float32)

    def inv97_1d(self, data, even=True):
        return self._1d("orc_inv97_1d", data, even, np.float32)

    def _2d(self, fn, data, even_row, even_col, dtype):
        a = np.ascontiguousarray(data, dtype=dtype).copy()
        h, w = a.shape
        getattr(self.lib, fn)(_p(a), C.c_int(w), C.c_int(h), C.c_int(w), C.c_int(int(even_row)), C.c_int(int(even_col)))
        return a

    def fwd53_2d(self, data, even_row=True, even_col=True):
        return self._2d("orc_fwd53_2d", data, even_row, even_col, np.int32)

    def inv53_2d(self, data, even_row=True, even_col=True):
        return self._2d("orc_inv53_2d", data, even_row, even_col, np.int32)

    def fwd97_2d(self, data, even_row=True, even_col=True):
        return self._2d("orc_fwd97_2d", data, even_row, even_col, np.float32)

    def inv97_2d(self, data, even_row=True, even_col=True):
        return self._2d("orc_inv97_2d", data, even_row, even_col, np.float32)

    def ll_dimensions(self, w, h, levels, x0=0, y0=0):
        a, b = C.c_int(), C.c_int()
        self.lib.orc_ll_dimensions(C.c_int(w), C.c_int(h), C.c_int(levels), C.c_int(x0), C.c_int(y0), C.byref(a), C.byref(b))
        return a.value, b.value

    def convert_f32_to_i32(self, data):
        a = np.ascontiguousarray(data, dtype=np.float32)
        out = np.empty(a.shape, np.int32)
        self.lib.orc_convert_f32_to_i32(_p(a), _p(out), C.c_size_t(a.size))
        return out

    def convert_f64_to_i32(self, data):
        a = np.ascontiguousarray(data, dtype=np.float64)
        out = np.empty(a.shape, np.int32)
        self.lib.orc_convert_f64_to_i32(_p(a), _p(out), C.c_size_t(a.size))
        return out

    # ---- colorspace
    def _color(self, fn, a, b, c, out_dtype=np.int32):
        a, b, c = (np.ascontiguousarray(v, dtype=np.int32) for v in (a, b, c))
        o = [np.empty(a.shape, out_dtype) for _ in range(3)]
        getattr(self.lib, fn)(C.c_size_t(a.size), _p(a), _p(b), _p(c), _p(o[0]), _p(o[1]), _p(o[2]))
        return o

    def rct_forward(self, r, g, b):
        return self._color("orc_rct_forward", r, g, b)

    def rct_inverse(self, y, cb, cr):
        return self._color("orc_rct_inverse", y, cb, cr)

    def ict_forward(self, r, g, b):
        return self._color("orc_ict_forward", r, g, b)

    def ict_inverse(self, y, cb, cr):
        return self._color("orc_ict_inverse", y, cb, cr)

    def ict_forward_f32(self, r, g, b):
        return self._color("orc_ict_forward_f32", r, g, b, np.float32)

    # ---- quantization
    def encode_quant_step(self, step, numbps):
        return int(self.lib.orc_encode_quant_step(float(step), int(numbps)))

    def openjpeg_quant_params(self, num_levels, bit_depth):
        n = 3 * max(num_levels, 0) + 1
        enc = np.zeros(n, np.uint16)
        st = np.zeros(n, np.float64)
        self.lib.orc_openjpeg_quant_params(C.c_int(num_levels), C.c_int(bit_depth), _p(enc), _p(st))
        return enc, st

    def quality_quant_params(self, quality, num_levels, bit_depth):
        n = 3 * max(num_levels, 0) + 1
        enc = np.zeros(n, np.uint16)
        st = np.zeros(n, np.float64)
        self.lib.orc_quality_quant_params(C.c_int(quality), C.c_int(num_levels), C.c_int(bit_depth), _p(enc), _p(st))
        return enc, st

    def runtime_quant_steps(self, encoded, num_levels, bit_depth):
        enc = np.ascontiguousarray(encoded, dtype=np.uint16)
        st = np.zeros(enc.size, np.float64)
        self.lib.orc_runtime_quant_steps(_p(enc), C.c_int(enc.size), C.c_int(num_levels), C.c_int(bit_depth), _p(st))
        return st

    def decode_quant_steps(self, encoded, num_levels, bit_depth, reversible=False):
        enc = np.ascontiguousarray(encoded, dtype=np.uint16)
        st = np.zeros(enc.size, np.float64)
        self.lib.orc_decode_quant_steps(_p(enc), C.c_int(enc.size), C.c_int(num_levels), C.c_int(bit_depth), C.c_int(int(reversible)), _p(st))
        return st

    def decode_quant_steps_derived(self, encoded, num_levels, bit_depth, reversible=False):
        st = np.zeros(3 * num_levels + 1, np.float64)
        self.lib.orc_decode_quant_steps_derived(C.c_uint16(int(encoded)), C.c_int(num_levels), C.c_int(bit_depth), C.c_int(int(reversible)), _p(st))
        return st

    def quantize_coefficients(self, data, step):
        a = np.ascontiguousarray(data, dtype=np.int32)
        out = np.empty_like(a)
        self.lib.orc_quantize_coefficients(_p(a), _p(out), a.size, float(step))
        return out

    def dequantize_coefficients(self, data, step):
        a = np.ascontiguousarray(data, dtype=np.int32)
        out = np.empty_like(a)
        self.lib.orc_dequantize_coefficients(_p(a), _p(out), a.size, float(step))
        return out

    def band_rects(self, w, h, x0, y0, levels):
        r = np.zeros((3 * levels + 1, 4), np.int32)
        n = self.lib.orc_band_rects(C.c_int(w), C.c_int(h), C.c_int(x0), C.c_int(y0), C.c_int(levels), _p(r))
        return r[:n]

    # ---- code-block interface
    def codeblock_layout(self, width, height, num_levels, cbw=64, cbh=64):
        n = self.lib.orc_codeblock_layout(C.c_int(width), C.c_int(height), C.c_int(num_levels), C.c_int(cbw), C.c_int(cbh), None, C.c_int(0))
        arr = (self.abi.Cblk * max(n, 1))()
        self.lib.orc_codeblock_layout(C.c_int(width), C.c_int(height), C.c_int(num_levels), C.c_int(cbw), C.c_int(cbh), arr, C.c_int(n))
        return [arr[i] for i in range(n)]

    def gather_blocks(self, plane, num_levels, cbw=64, cbh=64, shift6=False, htj2k=False):
        a = np.ascontiguousarray(plane, dtype=np.int32)
        h, w = a.shape
        n = self.lib.orc_codeblock_layout(C.c_int(w), C.c_int(h), C.c_int(num_levels), C.c_int(cbw), C.c_int(cbh), None, C.c_int(0))
        blocks = np.zeros(a.size, np.int32)
        nb = np.zeros(max(n, 1), np.int32)
        self.lib.orc_gather_blocks(_p(a), C.c_int(w), C.c_int(h), C.c_int(num_levels), C.c_int(cbw), C.c_int(cbh), C.c_int(int(shift6)),
                                   C.c_int(int(htj2k)), _p(blocks), _p(nb))
        return blocks, nb[:n]

    def scatter_blocks(self, blocks, width, height, num_levels, cbw=64, cbh=64):
        b = np.ascontiguousarray(blocks, dtype=np.int32)
        plane = np.empty((height, width), np.int32)
        self.lib.orc_scatter_blocks(_p(b), C.c_int(width), C.c_int(height), C.c_int(num_levels), C.c_int(cbw), C.c_int(cbh), _p(plane))
        return plane

    def inverse_general_scaling(self, data, shift, mask=None):
        a = np.ascontiguousarray(data, dtype=np.int32).copy()
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        self.lib.orc_inverse_general_scaling(_p(a), _p(m) if m is not None else None, C.c_size_t(a.size), C.c_int(int(shift)))
        return a

    def inverse_max_shift(self, data, shift):
        a = np.array(data, dtype=np.int32, copy=True).reshape(-1)
        self.lib.orc_inverse_max_shift(_p(a), C.c_size_t(a.size), C.c_int(shift))
        return a.reshape(np.shape(data))

    # ---- pipelines
    def fwd_tile_bounds(self, p, idx):
        b = (C.c_int32 * 4)()
        n = self.lib.orc_fwd_tile_bounds(C.byref(p), C.c_int(idx), b)
        return n, list(b)

    def inv_tile_bounds(self, p, idx):
        b = (C.c_int32 * 4)()
        cv = (C.c_int32 * 2)()
        n = self.lib.orc_inv_tile_bounds(C.byref(p), C.c_int(idx), b, cv)
        return n, list(b), list(cv)

    def forward(self, p, pixels):
        px = np.ascontiguousarray(pixels)
        out = np.empty(p.width * p.height * p.components, np.int32)
        rc = self.lib.orc_forward(C.byref(p), _p(px), _p(out))
        assert rc == 0
        return out

    def forward_planar(self, p, planes):
        pl = [np.ascontiguousarray(v, dtype=np.int32) for v in planes]
        arr = (C.c_void_p * len(pl))(*[v.ctypes.data for v in pl])
        out = np.empty(p.width * p.height * p.components, np.int32)
        rc = self.lib.orc_forward_planar(C.byref(p), arr, _p(out))
        assert rc == 0
        return out

    def inverse(self, p, coeffs, want_planes=False):
        co = np.ascontiguousarray(coeffs, dtype=np.int32)
        w, h = p.xsiz - p.xosiz, p.ysiz - p.yosiz
        nbytes = w * h * p.components * (1 if p.bit_depth <= 8 else 2)
        px = np.empty(nbytes, np.uint8)
        planes = np.empty((p.components, h, w), np.int32) if want_planes else None
        rc = self.lib.orc_inverse(C.byref(p), _p(co), _p(px), _p(planes) if want_planes else None)
        assert rc == 0
        return (px, planes) if want_planes else px

    def forward_batch(self, p, nframes, pixels, frame_stride, out, threads=1):
        return self.lib.orc_forward_batch(C.byref(p), C.c_int(nframes), _p(pixels), C.c_size_t(frame_stride), _p(out), C.c_int(threads))

    def inverse_batch(self, p, nframes, coeffs, out, frame_stride, threads=1):
        return self.lib.orc_inverse_batch(C.byref(p), C.c_int(nframes), _p(coeffs), _p(out), C.c_size_t(frame_stride), C.c_int(threads))
