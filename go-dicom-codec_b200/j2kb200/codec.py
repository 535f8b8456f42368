"""Context: the ctypes face of libj2kb200.so (numpy in, numpy out; device pointers for resident data)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi


class J2KError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"j2k_b200 error {code}: {msg}")
        self.code = code


def _vp(a):
    return C.c_void_p(a.ctypes.data)


class Context:
    """One library context (one or more CUDA devices).  Fails loudly without a device: no CPU fallback."""

    def __init__(self, devices=None, lib_path: str | None = None):
        self.lib = abi.load(lib_path)
        h = C.c_void_p()
        if devices is None:
            rc = self.lib.j2k_init(C.byref(h), None, 0)
        else:
            arr = (C.c_int * len(devices))(*devices)
            rc = self.lib.j2k_init(C.byref(h), arr, len(devices))
        if rc != 0:
            raise J2KError(rc, self.lib.j2k_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.lib.j2k_shutdown(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        if rc < 0:
            raise J2KError(rc, self.lib.j2k_last_error(self.h).decode())
        return rc

    # ---- bookkeeping
    @property
    def launch_count(self) -> int:
        return int(self.lib.j2k_launch_count(self.h))

    @property
    def device_count(self) -> int:
        return int(self.lib.j2k_device_count(self.h))

    def device_failed(self, slot: int) -> bool:
        """True when the device in `slot` was lost and removed from the round-robin (j2k_device_failed)."""
        return int(self.lib.j2k_device_failed(self.h, slot)) == 1

    def last_timing(self) -> abi.Timing:
        t = abi.Timing()
        self._ck(self.lib.j2k_last_timing(self.h, C.byref(t)))
        return t

    def set_profiling(self, on: bool):
        self._ck(self.lib.j2k_set_profiling(self.h, int(on)))

    def get_profile(self, max_launches: int = 256):
        """[(level, ms)] of the most recent profiled run (level 0 = pointwise kernel)."""
        ms = (C.c_float * max_launches)()
        lv = (C.c_int32 * max_launches)()
        n = self._ck(self.lib.j2k_get_profile(self.h, ms, lv, max_launches))
        return [(int(lv[i]), float(ms[i])) for i in range(min(n, max_launches))]

    def pinned(self, nbytes: int, dtype=np.uint8) -> np.ndarray:
        """A numpy view of library-owned pinned memory (j2k_acquire_buffer)."""
        p = self.lib.j2k_acquire_buffer(self.h, nbytes)
        if not p:
            raise J2KError(abi.J2K_ERR_NOMEM, self.lib.j2k_last_error(self.h).decode())
        buf = (C.c_uint8 * nbytes).from_address(p)
        return np.frombuffer(buf, dtype=np.uint8).view(dtype)

    def release(self, arr: np.ndarray):
        self.lib.j2k_release_buffer(self.h, C.c_void_p(arr.ctypes.data))

    # ---- forward
    def forward(self, p: abi.FwdParams, pixels) -> np.ndarray:
        px = np.ascontiguousarray(pixels).view(np.uint8).reshape(-1)
        n = self.lib.j2k_fwd_coeff_count(C.byref(p))
        out = np.empty(n, np.int32)
        self._ck(self.lib.j2k_forward(self.h, C.byref(p), _vp(px), px.size, _vp(out), out.size))
        return out

    def forward_planar(self, p: abi.FwdParams, planes) -> np.ndarray:
        pl = [np.ascontiguousarray(v, dtype=np.int32) for v in planes]
        arr = (C.c_void_p * len(pl))(*[v.ctypes.data for v in pl])
        out = np.empty(self.lib.j2k_fwd_coeff_count(C.byref(p)), np.int32)
        self._ck(self.lib.j2k_forward_planar(self.h, C.byref(p), arr, _vp(out), out.size))
        return out

    def forward_planar_flat(self, p: abi.FwdParams, planes: np.ndarray) -> np.ndarray:
        """planes: int32 [components, >= H*W] (row stride = plane stride): the cgo-callable twin of forward_planar."""
        pl = np.asarray(planes, dtype=np.int32)
        assert pl.ndim == 2 and pl.strides[1] == 4
        out = np.empty(self.lib.j2k_fwd_coeff_count(C.byref(p)), np.int32)
        self._ck(self.lib.j2k_forward_planar_flat(self.h, C.byref(p), _vp(pl), pl.strides[0] // 4, _vp(out), out.size))
        return out

    def forward_batch(self, p: abi.FwdParams, frames: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """frames: [nframes, frame_bytes] uint8 (row stride = frame stride)."""
        assert frames.ndim == 2 and frames.dtype == np.uint8 and frames.strides[1] == 1
        n = frames.shape[0]
        nc = self.lib.j2k_fwd_coeff_count(C.byref(p))
        if out is None:
            out = np.empty((n, nc), np.int32)
        self._ck(self.lib.j2k_forward_batch(self.h, C.byref(p), n, _vp(frames), frames.strides[0], _vp(out)))
        return out

    def submit_forward(self, p: abi.FwdParams, frames: np.ndarray, out: np.ndarray) -> int:
        return self._ck(self.lib.j2k_submit_forward(self.h, C.byref(p), frames.shape[0], _vp(frames), frames.strides[0], _vp(out)))

    def submit_inverse(self, p: abi.InvParams, coeffs: np.ndarray, out: np.ndarray) -> int:
        return self._ck(self.lib.j2k_submit_inverse(self.h, C.byref(p), coeffs.shape[0], _vp(coeffs), _vp(out), out.strides[0], None))

    def wait(self, ticket: int):
        self._ck(self.lib.j2k_wait(self.h, ticket))

    def forward_device(self, p: abi.FwdParams, nframes: int, d_pixels: int, frame_stride_bytes: int, d_coeffs: int,
                       stream: int = 0, dev: int = 0):
        self._ck(self.lib.j2k_forward_device(self.h, dev, C.byref(p), nframes, C.c_void_p(d_pixels), frame_stride_bytes,
                                             C.c_void_p(d_coeffs), C.c_void_p(stream)))

    # ---- inverse
    def inverse(self, p: abi.InvParams, coeffs, want_planes: bool = False):
        co = np.ascontiguousarray(coeffs, dtype=np.int32).reshape(-1)
        nb = self.lib.j2k_inv_pixel_bytes(C.byref(p))
        px = np.empty(nb, np.uint8)
        w, h = p.xsiz - p.xosiz, p.ysiz - p.yosiz
        planes = np.empty((p.components, h, w), np.int32) if want_planes else None
        self._ck(self.lib.j2k_inverse(self.h, C.byref(p), _vp(co), co.size, _vp(px), px.size,
                                      _vp(planes) if want_planes else None))
        return (px, planes) if want_planes else px

    def inverse_batch(self, p: abi.InvParams, coeffs: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        assert coeffs.ndim == 2 and coeffs.dtype == np.int32 and coeffs.flags.c_contiguous
        n = coeffs.shape[0]
        nb = self.lib.j2k_inv_pixel_bytes(C.byref(p))
        if out is None:
            out = np.empty((n, nb), np.uint8)
        self._ck(self.lib.j2k_inverse_batch(self.h, C.byref(p), n, _vp(coeffs), _vp(out), out.strides[0], None))
        return out

    def inverse_device(self, p: abi.InvParams, nframes: int, d_coeffs: int, d_pixels: int, frame_stride_bytes: int,
                       d_planes: int = 0, stream: int = 0, dev: int = 0):
        self._ck(self.lib.j2k_inverse_device(self.h, dev, C.byref(p), nframes, C.c_void_p(d_coeffs), C.c_void_p(d_pixels),
                                             frame_stride_bytes, C.c_void_p(d_planes) if d_planes else None, C.c_void_p(stream)))

    # ---- wavelet package API
    # ---- code-block interface
    def codeblock_layout(self, width, height, num_levels, cb_width=64, cb_height=64):
        """Blocks of one tile-component plane in the reference's order (encoder.go:2424-2431,3059-3285)."""
        n = self._ck(self.lib.j2k_codeblock_layout(width, height, num_levels, cb_width, cb_height, None, 0))
        arr = (abi.Cblk * max(n, 1))()
        self._ck(self.lib.j2k_codeblock_layout(width, height, num_levels, cb_width, cb_height, arr, n))
        return [arr[i] for i in range(n)]

    def forward_blocks(self, p: abi.FwdParams, frames: np.ndarray, cb_width=64, cb_height=64):
        """frames: [nframes, frame_bytes] uint8 -> (block-major coefficients [nframes, coeff_count], numbps [nframes, nblocks])."""
        assert frames.ndim == 2 and frames.dtype == np.uint8 and frames.strides[1] == 1
        n = frames.shape[0]
        out = np.empty((n, self.lib.j2k_fwd_coeff_count(C.byref(p))), np.int32)
        nb = np.empty((n, self.lib.j2k_fwd_block_count(C.byref(p), cb_width, cb_height)), np.int32)
        self._ck(self.lib.j2k_forward_blocks(self.h, C.byref(p), cb_width, cb_height, n, _vp(frames), frames.strides[0], _vp(out), _vp(nb)))
        return out, nb

    def inverse_blocks(self, p: abi.InvParams, blocks: np.ndarray, cb_width=64, cb_height=64, want_planes: bool = False,
                       roi_maxshift=None, block_scale_shift=None, sample_mask=None):
        """roi_maxshift: per-component MaxShift (RGN Srgn = 0) undone on the device while the blocks are scattered
        (decodeCodeBlock, t2/tile_decoder.go:726-730); None = the blocks carry no ROI scaling.
        block_scale_shift [nframes, nblocks] (+ optional sample_mask [nframes, coeff_count] bytes, block-major): general
        scaling (Srgn = 1, :735-742) of the blocks the region touches."""
        assert blocks.ndim == 2 and blocks.dtype == np.int32 and blocks.flags.c_contiguous
        n = blocks.shape[0]
        nbytes = self.lib.j2k_inv_pixel_bytes(C.byref(p))
        out = np.empty((n, nbytes), np.uint8)
        w, h = p.xsiz - p.xosiz, p.ysiz - p.yosiz
        planes = np.empty((n, p.components, h, w), np.int32) if want_planes else None
        if block_scale_shift is not None:
            bs = np.ascontiguousarray(block_scale_shift, dtype=np.int32)
            mk = None if sample_mask is None else np.ascontiguousarray(sample_mask, dtype=np.uint8)
            roi = None if roi_maxshift is None else np.ascontiguousarray(roi_maxshift, dtype=np.int32)
            self._ck(self.lib.j2k_inverse_blocks_roi_general(self.h, C.byref(p), cb_width, cb_height, n, _vp(blocks),
                                                             _vp(roi) if roi is not None else None, _vp(bs),
                                                             _vp(mk) if mk is not None else None, _vp(out), nbytes,
                                                             _vp(planes) if want_planes else None))
        elif roi_maxshift is None:
            self._ck(self.lib.j2k_inverse_blocks(self.h, C.byref(p), cb_width, cb_height, n, _vp(blocks), _vp(out), nbytes,
                                                 _vp(planes) if want_planes else None))
        else:
            roi = np.ascontiguousarray(roi_maxshift, dtype=np.int32)
            assert roi.size == p.components
            self._ck(self.lib.j2k_inverse_blocks_roi(self.h, C.byref(p), cb_width, cb_height, n, _vp(blocks), _vp(roi), _vp(out), nbytes,
                                                     _vp(planes) if want_planes else None))
        return (out, planes) if want_planes else out

    # ---- HTJ2K block decoding on the device (decode side)
    @staticmethod
    def ht_records(offsets, lengths, kmax, missing_msbs) -> np.ndarray:
        """j2k_ht_cblk records (one per frame and block, code-block interface order) from four parallel arrays."""
        rec = np.zeros(len(offsets), dtype=np.dtype([("offset", "<u8"), ("length", "<u4"), ("kmax", "u1"), ("missing_msbs", "u1"),
                                                      ("reserved", "<u2")]))
        rec["offset"], rec["length"], rec["kmax"], rec["missing_msbs"] = offsets, lengths, kmax, missing_msbs
        assert rec.dtype.itemsize == C.sizeof(abi.HtCblk)
        return rec

    def ht_decode_blocks(self, p: abi.InvParams, nframes: int, stream: np.ndarray, records: np.ndarray, cb_width=64, cb_height=64):
        """HTDecoder.Decode (jpeg2000/htj2k/decoder.go:43-58) for every code-block: -> (block-major planes [nframes, coeffs],
        status [nframes, blocks]); a failing block reads as zeros (t2/tile_decoder.go:718-721)."""
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        nblk = self.lib.j2k_inv_block_count(C.byref(p), cb_width, cb_height)
        assert records.size == nframes * nblk, (records.size, nframes, nblk)
        out = np.empty((nframes, self.lib.j2k_inv_coeff_count(C.byref(p))), np.int32)
        status = np.empty((nframes, nblk), np.int32)
        self._ck(self.lib.j2k_ht_decode_blocks(self.h, C.byref(p), cb_width, cb_height, nframes, _vp(stream), stream.size, _vp(records),
                                               _vp(out), _vp(status)))
        return out, status

    def inverse_ht(self, p: abi.InvParams, nframes: int, stream: np.ndarray, records: np.ndarray, cb_width=64, cb_height=64,
                   want_planes: bool = False):
        """HT block decoding + assembleSubbands + the whole inverse path on the device: cleanup segments in, pixels out.
        -> (pixels [nframes, bytes], status [nframes, blocks]) (+ planes)."""
        stream = np.ascontiguousarray(stream, dtype=np.uint8)
        nblk = self.lib.j2k_inv_block_count(C.byref(p), cb_width, cb_height)
        assert records.size == nframes * nblk, (records.size, nframes, nblk)
        nbytes = self.lib.j2k_inv_pixel_bytes(C.byref(p))
        out = np.empty((nframes, nbytes), np.uint8)
        status = np.empty((nframes, nblk), np.int32)
        w, h = p.xsiz - p.xosiz, p.ysiz - p.yosiz
        planes = np.empty((nframes, p.components, h, w), np.int32) if want_planes else None
        self._ck(self.lib.j2k_inverse_ht(self.h, C.byref(p), cb_width, cb_height, nframes, _vp(stream), stream.size, _vp(records), _vp(out),
                                         nbytes, _vp(planes) if want_planes else None, _vp(status)))
        return (out, status, planes) if want_planes else (out, status)

    def submit_inverse_ht(self, p: abi.InvParams, nframes: int, stream: np.ndarray, records: np.ndarray, out: np.ndarray,
                          status: np.ndarray | None = None, cb_width=64, cb_height=64) -> int:
        """Ticketed inverse_ht: `stream`, `out` (and `status`) come from acquire(); `records` may be any array (it is copied)."""
        return self._ck(self.lib.j2k_submit_inverse_ht(self.h, C.byref(p), cb_width, cb_height, nframes, _vp(stream), stream.size,
                                                       _vp(records), _vp(out), out.strides[0], None,
                                                       _vp(status) if status is not None else None))

    def ht_decode_device(self, p: abi.InvParams, nframes, d_bytes: int, d_records: int, d_out: int, to_planes: bool, d_status: int = 0,
                         cb_width=64, cb_height=64, dev=0, stream=0):
        self._ck(self.lib.j2k_ht_decode_device(self.h, dev, C.byref(p), cb_width, cb_height, nframes, d_bytes, d_records, d_out,
                                               int(to_planes), d_status or None, stream or None))

    # ---- HTJ2K block encoding on the device (encode side)
    def forward_ht(self, p: abi.FwdParams, frames: np.ndarray, kmax, cb_width=64, cb_height=64, out: np.ndarray | None = None):
        """j2k_forward_batch + HTEncoder.Encode (jpeg2000/htj2k/encoder.go:54-68) of every code-block on the device.
        frames [n, frame bytes]; kmax [components, 3 * levels + 1] band precisions (Encoder.bandNumbps).
        -> (stream bytes, records [n * blocks] as a structured array: offset, length, kmax, missing_msbs)."""
        assert frames.ndim == 2 and frames.dtype == np.uint8 and frames.flags.c_contiguous
        n = frames.shape[0]
        km = np.ascontiguousarray(kmax, dtype=np.uint8).reshape(-1)
        assert km.size == p.components * (3 * p.num_levels + 1), km.size
        nblk = self.lib.j2k_fwd_block_count(C.byref(p), cb_width, cb_height)
        if out is None:
            cap = self.lib.j2k_ht_encode_bound(C.byref(p), cb_width, cb_height, int(km.max()) if km.size and 0 < km.max() < 31 else 30, n)
            out = np.empty(max(cap, 16), np.uint8)
        rec = self.ht_records(np.zeros(n * nblk, np.uint64), np.zeros(n * nblk, np.uint32), np.zeros(n * nblk, np.uint8),
                              np.zeros(n * nblk, np.uint8))
        got = C.c_size_t(0)
        self._ck(self.lib.j2k_forward_ht(self.h, C.byref(p), cb_width, cb_height, n, _vp(frames), frames.strides[0], _vp(km), _vp(out),
                                         out.size, C.byref(got), _vp(rec)))
        return out[:got.value], rec

    def gather_blocks_device(self, p: abi.FwdParams, nframes, d_coeffs: int, d_blocks: int, d_numbps: int, cb_width=64, cb_height=64,
                             stream: int = 0, dev: int = 0):
        self._ck(self.lib.j2k_gather_blocks_device(self.h, dev, C.byref(p), cb_width, cb_height, nframes, C.c_void_p(d_coeffs),
                                                   C.c_void_p(d_blocks), C.c_void_p(d_numbps), C.c_void_p(stream)))

    def scatter_blocks_device(self, p: abi.InvParams, nframes, d_blocks: int, d_coeffs: int, cb_width=64, cb_height=64,
                              stream: int = 0, dev: int = 0, roi_maxshift=None):
        if roi_maxshift is None:
            self._ck(self.lib.j2k_scatter_blocks_device(self.h, dev, C.byref(p), cb_width, cb_height, nframes, C.c_void_p(d_blocks),
                                                        C.c_void_p(d_coeffs), C.c_void_p(stream)))
        else:
            roi = np.ascontiguousarray(roi_maxshift, dtype=np.int32)
            assert roi.size == p.components
            self._ck(self.lib.j2k_scatter_blocks_roi_device(self.h, dev, C.byref(p), cb_width, cb_height, nframes, C.c_void_p(d_blocks),
                                                            _vp(roi), C.c_void_p(d_coeffs), C.c_void_p(stream)))

    def _dwt(self, fn, data, levels, x0, y0, dtype):
        a = np.ascontiguousarray(data, dtype=dtype).copy()
        h, w = a.shape
        self._ck(getattr(self.lib, fn)(self.h, _vp(a), w, h, levels, x0, y0))
        return a

    def dwt53_forward(self, data, levels, x0=0, y0=0):
        return self._dwt("j2k_dwt53_forward", data, levels, x0, y0, np.int32)

    def dwt53_inverse(self, data, levels, x0=0, y0=0):
        return self._dwt("j2k_dwt53_inverse", data, levels, x0, y0, np.int32)

    def dwt97_forward(self, data, levels, x0=0, y0=0):
        return self._dwt("j2k_dwt97_forward", data, levels, x0, y0, np.float32)

    def dwt97_inverse(self, data, levels, x0=0, y0=0):
        return self._dwt("j2k_dwt97_inverse", data, levels, x0, y0, np.float32)

    # float64 wrappers of the wavelet package (dwt97.go:340-351,410-421,515-526) and layout.go
    def dwt97_forward_f64(self, data, levels, x0=0, y0=0):
        return self._dwt("j2k_dwt97_forward_f64", data, levels, x0, y0, np.float64)

    def dwt97_inverse_f64(self, data, levels, x0=0, y0=0):
        return self._dwt("j2k_dwt97_inverse_f64", data, levels, x0, y0, np.float64)

    def convert_f64_to_i32(self, data):
        a = np.ascontiguousarray(data, dtype=np.float64)
        out = np.empty(a.shape, np.int32)
        self._ck(self.lib.j2k_convert_f64_to_i32(self.h, _vp(a), _vp(out), a.size))
        return out

    def ll_dimensions(self, width, height, levels, x0=0, y0=0):
        a, b = C.c_int(), C.c_int()
        self._ck(self.lib.j2k_ll_dimensions(width, height, levels, x0, y0, C.byref(a), C.byref(b)))
        return a.value, b.value

    # colorspace/rgb.go
    def convert_rgb_to_ycbcr(self, rgb, width, height):
        a = np.ascontiguousarray(rgb, dtype=np.int32).reshape(-1)
        o = [np.empty(width * height, np.int32) for _ in range(3)]
        self._ck(self.lib.j2k_rgb_to_ycbcr(self.h, _vp(a), width, height, _vp(o[0]), _vp(o[1]), _vp(o[2])))
        return o

    def convert_ycbcr_to_rgb(self, y, cb, cr, width, height):
        y, cb, cr = (np.ascontiguousarray(v, dtype=np.int32) for v in (y, cb, cr))
        out = np.empty(width * height * 3, np.int32)
        self._ck(self.lib.j2k_ycbcr_to_rgb(self.h, _vp(y), _vp(cb), _vp(cr), width, height, _vp(out)))
        return out

    def interleave_components(self, components):
        if len(components) == 0:
            return None  # rgb.go:55-57
        pl = [np.ascontiguousarray(v, dtype=np.int32).reshape(-1) for v in components]
        arr = (C.c_void_p * len(pl))(*[v.ctypes.data for v in pl])
        out = np.empty(pl[0].size * len(pl), np.int32)
        self._ck(self.lib.j2k_interleave_components(self.h, arr, len(pl), pl[0].size, _vp(out)))
        return out

    def deinterleave_components(self, data, num_components):
        a = np.ascontiguousarray(data, dtype=np.int32).reshape(-1)
        if a.size == 0 or num_components == 0:
            return None  # rgb.go:77-79
        n = a.size // num_components
        out = [np.empty(n, np.int32) for _ in range(num_components)]
        arr = (C.c_void_p * num_components)(*[v.ctypes.data for v in out])
        self._ck(self.lib.j2k_deinterleave_components(self.h, _vp(a), n, num_components, arr))
        return out

    def convert_f32_to_i32(self, data):
        a = np.ascontiguousarray(data, dtype=np.float32)
        out = np.empty(a.shape, np.int32)
        self._ck(self.lib.j2k_convert_f32_to_i32(self.h, _vp(a), _vp(out), a.size))
        return out

    # ---- colorspace / quantization package API
    def _c3(self, fn, a, b, c):
        a, b, c = (np.ascontiguousarray(v, dtype=np.int32) for v in (a, b, c))
        o = [np.empty(a.shape, np.int32) for _ in range(3)]
        self._ck(getattr(self.lib, fn)(self.h, a.size, _vp(a), _vp(b), _vp(c), _vp(o[0]), _vp(o[1]), _vp(o[2])))
        return o

    def rct_forward(self, r, g, b):
        return self._c3("j2k_rct_forward", r, g, b)

    def rct_inverse(self, y, cb, cr):
        return self._c3("j2k_rct_inverse", y, cb, cr)

    def ict_forward(self, r, g, b):
        return self._c3("j2k_ict_forward", r, g, b)

    def ict_inverse(self, y, cb, cr):
        return self._c3("j2k_ict_inverse", y, cb, cr)

    def quantize_coefficients(self, data, step):
        a = np.ascontiguousarray(data, dtype=np.int32)
        out = np.empty_like(a)
        self._ck(self.lib.j2k_quantize_coefficients(self.h, _vp(a), _vp(out), a.size, float(step)))
        return out

    def dequantize_coefficients(self, data, step):
        a = np.ascontiguousarray(data, dtype=np.int32)
        out = np.empty_like(a)
        self._ck(self.lib.j2k_dequantize_coefficients(self.h, _vp(a), _vp(out), a.size, float(step)))
        return out


# ---- scalar step tables (host only, no context needed)
def openjpeg_quant_params(num_levels: int, bit_depth: int, lib_path: str | None = None):
    lib = abi.load(lib_path)
    n = 3 * max(num_levels, 0) + 1
    enc, st = np.zeros(n, np.uint16), np.zeros(n, np.float64)
    lib.j2k_quant_openjpeg_params(num_levels, bit_depth, enc.ctypes.data_as(C.POINTER(C.c_uint16)), st.ctypes.data_as(C.POINTER(C.c_double)))
    return enc, st


def quality_quant_params(quality: int, num_levels: int, bit_depth: int, lib_path: str | None = None):
    lib = abi.load(lib_path)
    n = 3 * max(num_levels, 0) + 1
    enc, st = np.zeros(n, np.uint16), np.zeros(n, np.float64)
    lib.j2k_quant_quality_params(quality, num_levels, bit_depth, enc.ctypes.data_as(C.POINTER(C.c_uint16)), st.ctypes.data_as(C.POINTER(C.c_double)))
    return enc, st


def runtime_quant_steps(encoded, num_levels: int, bit_depth: int, lib_path: str | None = None):
    lib = abi.load(lib_path)
    enc = np.ascontiguousarray(encoded, dtype=np.uint16)
    st = np.zeros(enc.size, np.float64)
    lib.j2k_quant_runtime_steps(enc.ctypes.data_as(C.POINTER(C.c_uint16)), enc.size, num_levels, bit_depth, st.ctypes.data_as(C.POINTER(C.c_double)))
    return st


def decode_quant_steps(encoded, num_levels: int, bit_depth: int, reversible: bool = False, lib_path: str | None = None):
    lib = abi.load(lib_path)
    enc = np.ascontiguousarray(encoded, dtype=np.uint16)
    st = np.zeros(enc.size, np.float64)
    lib.j2k_quant_decode_steps(enc.ctypes.data_as(C.POINTER(C.c_uint16)), enc.size, num_levels, bit_depth, int(reversible), st.ctypes.data_as(C.POINTER(C.c_double)))
    return st
