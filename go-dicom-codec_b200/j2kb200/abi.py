"""ctypes view of include/j2k_b200.h and the loader of libj2kb200.so.

The library is the product; there is no CPU fallback.  `load()` raises if the
shared object has not been built (run `python -c "import __graft_entry__ as g; g.build()"`),
and every compute entry point returns J2K_ERR_CUDA when no CUDA device exists.
"""
from __future__ import annotations

import ctypes as C
import os

J2K_MAX_COMPONENTS = 4
J2K_MAX_LEVELS = 10
J2K_MAX_BANDS = 3 * J2K_MAX_LEVELS + 1
J2K_MAX_BINDINGS = 4

J2K_OK = 0
J2K_ERR_INVALID_ARG = -1
J2K_ERR_SIZE = -2
J2K_ERR_CUDA = -3
J2K_ERR_NOMEM = -4
J2K_ERR_UNSUPPORTED = -5
J2K_ERR_TICKET = -6

MCT_NONE, MCT_RCT, MCT_ICT, MCT_CUSTOM_INT, MCT_CUSTOM_Q13, MCT_CUSTOM_FLOAT, MCT_BINDINGS = range(7)


class MctBinding(C.Structure):
    _fields_ = [
        ("n_components", C.c_int32),
        ("component_ids", C.c_int32 * J2K_MAX_COMPONENTS),
        ("element_type", C.c_int32),
        ("has_matrix", C.c_int32),
        ("matrix", C.c_double * (J2K_MAX_COMPONENTS * J2K_MAX_COMPONENTS)),
        ("has_offsets", C.c_int32),
        ("offsets", C.c_int32 * J2K_MAX_COMPONENTS),
    ]


class FwdParams(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("components", C.c_int32),
        ("bit_depth", C.c_int32), ("is_signed", C.c_int32),
        ("tile_width", C.c_int32), ("tile_height", C.c_int32),
        ("num_levels", C.c_int32), ("reversible", C.c_int32), ("htj2k", C.c_int32),
        ("mct_mode", C.c_int32),
        ("mct_matrix", C.c_double * (J2K_MAX_COMPONENTS * J2K_MAX_COMPONENTS)),
        ("mct_has_offsets", C.c_int32),
        ("mct_offsets", C.c_int32 * J2K_MAX_COMPONENTS),
        ("n_bindings", C.c_int32),
        ("bindings", MctBinding * J2K_MAX_BINDINGS),
        ("n_steps", C.c_int32),
        ("steps", C.c_double * J2K_MAX_BANDS),
        ("fuse_t1_shift", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


class InvParams(C.Structure):
    _fields_ = [
        ("xsiz", C.c_int32), ("ysiz", C.c_int32), ("xosiz", C.c_int32), ("yosiz", C.c_int32),
        ("xtsiz", C.c_int32), ("ytsiz", C.c_int32), ("xtosiz", C.c_int32), ("ytosiz", C.c_int32),
        ("components", C.c_int32), ("bit_depth", C.c_int32), ("is_signed", C.c_int32),
        ("num_levels", C.c_int32), ("reversible", C.c_int32), ("htj2k", C.c_int32),
        ("n_steps", C.c_int32),
        ("steps", C.c_double * J2K_MAX_BANDS),
        ("mct_mode", C.c_int32),
        ("mct_matrix", C.c_double * (J2K_MAX_COMPONENTS * J2K_MAX_COMPONENTS)),
        ("mct_has_offsets", C.c_int32),
        ("mct_offsets", C.c_int32 * J2K_MAX_COMPONENTS),
        ("n_bindings", C.c_int32),
        ("bindings", MctBinding * J2K_MAX_BINDINGS),
        ("fuse_t1_halve", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


class Cblk(C.Structure):
    """j2k_cblk: one code-block of a tile-component plane (codeBlockInfo, jpeg2000/encoder.go:3199-3212)."""
    _fields_ = [
        ("x0", C.c_int32), ("y0", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
        ("cbx", C.c_int32), ("cby", C.c_int32), ("band", C.c_int32), ("res", C.c_int32),
        ("offset", C.c_int64),
    ]


class HtCblk(C.Structure):
    """j2k_ht_cblk: one HT code-block as T2 leaves it (cbInfo, jpeg2000/t2/tile_decoder.go:453-526) plus the coding context
    of HTDecoder.SetCodingContext (jpeg2000/htj2k/decoder.go:92-96)."""
    _fields_ = [("offset", C.c_uint64), ("length", C.c_uint32), ("kmax", C.c_uint8), ("missing_msbs", C.c_uint8), ("reserved", C.c_uint16)]


HT_OK, HT_ERR_KMAX, HT_ERR_SEGMENT, HT_ERR_UQ = 0, -1, -2, -3


class Timing(C.Structure):
    _fields_ = [
        ("h2d_ms", C.c_float), ("kernel_ms", C.c_float), ("d2h_ms", C.c_float), ("total_ms", C.c_float),
        ("kernel_launches", C.c_int32), ("reserved", C.c_int32),
    ]


def make_binding(component_ids=(), matrix=None, offsets=None, element_type=0) -> MctBinding:
    b = MctBinding()
    b.n_components = len(component_ids)
    for i, c in enumerate(component_ids):
        b.component_ids[i] = int(c)
    b.element_type = int(element_type)
    if matrix is not None:
        b.has_matrix = 1
        flat = [float(v) for row in matrix for v in row]
        for i, v in enumerate(flat):
            b.matrix[i] = v
    if offsets is not None:
        b.has_offsets = 1
        for i, v in enumerate(offsets):
            b.offsets[i] = int(v)
    return b


def fwd_params(width, height, components=1, bit_depth=8, is_signed=False, tile_width=0, tile_height=0,
               num_levels=5, reversible=True, htj2k=False, mct_mode=MCT_NONE, steps=None,
               mct_matrix=None, mct_offsets=None, bindings=(), fuse_t1_shift=False) -> FwdParams:
    p = FwdParams()
    p.width, p.height, p.components = int(width), int(height), int(components)
    p.bit_depth, p.is_signed = int(bit_depth), int(bool(is_signed))
    p.tile_width, p.tile_height = int(tile_width), int(tile_height)
    p.num_levels, p.reversible, p.htj2k = int(num_levels), int(bool(reversible)), int(bool(htj2k))
    p.mct_mode = int(mct_mode)
    if mct_matrix is not None:
        flat = [float(v) for row in mct_matrix for v in row]
        for i, v in enumerate(flat):
            p.mct_matrix[i] = v
    if mct_offsets is not None:
        p.mct_has_offsets = 1
        for i, v in enumerate(mct_offsets):
            p.mct_offsets[i] = int(v)
    p.n_bindings = len(bindings)
    for i, b in enumerate(bindings):
        p.bindings[i] = b
    if steps is not None:
        p.n_steps = len(steps)
        for i, v in enumerate(steps):
            p.steps[i] = float(v)
    p.fuse_t1_shift = int(bool(fuse_t1_shift))
    return p


def inv_params(width, height, components=1, bit_depth=8, is_signed=False, tile_width=0, tile_height=0,
               num_levels=5, reversible=True, htj2k=False, mct_mode=MCT_NONE, steps=None,
               mct_matrix=None, mct_offsets=None, bindings=(), fuse_t1_halve=False,
               xosiz=0, yosiz=0, xtosiz=0, ytosiz=0) -> InvParams:
    """`width`/`height` are the image size; SIZ fields are derived (Xsiz = XOsiz + width)."""
    p = InvParams()
    p.xosiz, p.yosiz = int(xosiz), int(yosiz)
    p.xsiz, p.ysiz = int(xosiz) + int(width), int(yosiz) + int(height)
    p.xtosiz, p.ytosiz = int(xtosiz), int(ytosiz)
    p.xtsiz = int(tile_width) if tile_width else p.xsiz - p.xtosiz
    p.ytsiz = int(tile_height) if tile_height else p.ysiz - p.ytosiz
    p.components, p.bit_depth, p.is_signed = int(components), int(bit_depth), int(bool(is_signed))
    p.num_levels, p.reversible, p.htj2k = int(num_levels), int(bool(reversible)), int(bool(htj2k))
    p.mct_mode = int(mct_mode)
    if mct_matrix is not None:
        flat = [float(v) for row in mct_matrix for v in row]
        for i, v in enumerate(flat):
            p.mct_matrix[i] = v
    if mct_offsets is not None:
        p.mct_has_offsets = 1
        for i, v in enumerate(mct_offsets):
            p.mct_offsets[i] = int(v)
    p.n_bindings = len(bindings)
    for i, b in enumerate(bindings):
        p.bindings[i] = b
    if steps is not None:
        p.n_steps = len(steps)
        for i, v in enumerate(steps):
            p.steps[i] = float(v)
    p.fuse_t1_halve = int(bool(fuse_t1_halve))
    return p


_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_PKG_DIR), "csrc", "build", "libj2kb200.so")

# every symbol include/j2k_b200.h declares (tests check that the library exports all of them)
EXPORTED_SYMBOLS = [
    "j2k_init", "j2k_shutdown", "j2k_last_error", "j2k_last_error_copy", "j2k_abi_version", "j2k_device_count", "j2k_device_failed", "j2k_visible_devices", "j2k_launch_count",
    "j2k_last_timing", "j2k_set_profiling", "j2k_get_profile", "j2k_acquire_buffer", "j2k_release_buffer",
    "j2k_fwd_pixel_bytes", "j2k_fwd_coeff_count", "j2k_inv_pixel_bytes", "j2k_inv_coeff_count",
    "j2k_fwd_tile_bounds", "j2k_inv_tile_bounds",
    "j2k_forward", "j2k_forward_planar", "j2k_forward_planar_flat", "j2k_forward_batch", "j2k_forward_device",
    "j2k_inverse", "j2k_inverse_batch", "j2k_inverse_device",
    "j2k_submit_forward", "j2k_submit_inverse", "j2k_wait",
    "j2k_codeblock_layout", "j2k_fwd_block_count", "j2k_inv_block_count", "j2k_forward_blocks", "j2k_inverse_blocks",
    "j2k_gather_blocks_device", "j2k_scatter_blocks_device", "j2k_inverse_blocks_roi", "j2k_scatter_blocks_roi_device",
    "j2k_inverse_blocks_roi_general", "j2k_scatter_blocks_roi_general_device",
    "j2k_dwt53_forward", "j2k_dwt53_inverse", "j2k_dwt97_forward", "j2k_dwt97_inverse", "j2k_convert_f32_to_i32",
    "j2k_rct_forward", "j2k_rct_inverse", "j2k_ict_forward", "j2k_ict_inverse",
    "j2k_dwt97_forward_f64", "j2k_dwt97_inverse_f64", "j2k_convert_f64_to_i32", "j2k_ll_dimensions",
    "j2k_rgb_to_ycbcr", "j2k_ycbcr_to_rgb", "j2k_interleave_components", "j2k_deinterleave_components",
    "j2k_quantize_coefficients", "j2k_dequantize_coefficients",
    "j2k_quant_openjpeg_params", "j2k_quant_quality_params", "j2k_quant_runtime_steps", "j2k_quant_decode_steps",
]

_libs = {}


def load(path: str | None = None) -> C.CDLL:
    """Load libj2kb200.so and declare the prototypes.  Raises if it is not built.

    `path` exists for the test-suite's CPU emulator build of the same sources (tests/emu);
    the product always loads LIB_PATH.
    """
    path = path or os.environ.get("J2K_B200_LIB") or LIB_PATH  # the env override is for kernel experiments (profiles/)
    if path in _libs:
        return _libs[path]
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: the CUDA extension is the product and has no fallback; "
            "build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    lib = C.CDLL(path)
    vp, i32p, f32p, u16p, f64p = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(C.c_uint16), C.POINTER(C.c_double)
    sz, ci = C.c_size_t, C.c_int
    FP, IP = C.POINTER(FwdParams), C.POINTER(InvParams)
    sig = {
        "j2k_init": (ci, [C.POINTER(vp), C.POINTER(ci), ci]),
        "j2k_shutdown": (None, [vp]),
        "j2k_last_error": (C.c_char_p, [vp]),
        "j2k_last_error_copy": (C.c_size_t, [vp, C.c_char_p, C.c_size_t]),
        "j2k_abi_version": (ci, []),
        "j2k_device_count": (ci, [vp]),
        "j2k_visible_devices": (ci, []),
        "j2k_device_failed": (ci, [vp, ci]),
        "j2k_launch_count": (C.c_int64, [vp]),
        "j2k_last_timing": (ci, [vp, C.POINTER(Timing)]),
        "j2k_set_profiling": (ci, [vp, ci]),
        "j2k_get_profile": (ci, [vp, f32p, i32p, ci]),
        "j2k_acquire_buffer": (vp, [vp, sz]),
        "j2k_release_buffer": (None, [vp, vp]),
        "j2k_fwd_pixel_bytes": (sz, [FP]),
        "j2k_fwd_coeff_count": (sz, [FP]),
        "j2k_inv_pixel_bytes": (sz, [IP]),
        "j2k_inv_coeff_count": (sz, [IP]),
        "j2k_fwd_tile_bounds": (ci, [FP, ci, i32p]),
        "j2k_inv_tile_bounds": (ci, [IP, ci, i32p]),
        "j2k_forward": (ci, [vp, FP, vp, sz, vp, sz]),
        "j2k_forward_planar": (ci, [vp, FP, C.POINTER(vp), vp, sz]),
        "j2k_forward_planar_flat": (ci, [vp, FP, vp, sz, vp, sz]),
        "j2k_forward_batch": (ci, [vp, FP, ci, vp, sz, vp]),
        "j2k_forward_device": (ci, [vp, ci, FP, ci, vp, sz, vp, vp]),
        "j2k_inverse": (ci, [vp, IP, vp, sz, vp, sz, vp]),
        "j2k_inverse_batch": (ci, [vp, IP, ci, vp, vp, sz, vp]),
        "j2k_inverse_device": (ci, [vp, ci, IP, ci, vp, vp, sz, vp, vp]),
        "j2k_submit_forward": (C.c_int64, [vp, FP, ci, vp, sz, vp]),
        "j2k_submit_inverse": (C.c_int64, [vp, IP, ci, vp, vp, sz, vp]),
        "j2k_wait": (ci, [vp, C.c_int64]),
        "j2k_codeblock_layout": (ci, [ci, ci, ci, ci, ci, C.POINTER(Cblk), ci]),
        "j2k_fwd_block_count": (sz, [FP, ci, ci]),
        "j2k_inv_block_count": (sz, [IP, ci, ci]),
        "j2k_forward_blocks": (ci, [vp, FP, ci, ci, ci, vp, sz, vp, vp]),
        "j2k_inverse_blocks": (ci, [vp, IP, ci, ci, ci, vp, vp, sz, vp]),
        "j2k_gather_blocks_device": (ci, [vp, ci, FP, ci, ci, ci, vp, vp, vp, vp]),
        "j2k_scatter_blocks_device": (ci, [vp, ci, IP, ci, ci, ci, vp, vp, vp]),
        "j2k_inverse_blocks_roi": (ci, [vp, IP, ci, ci, ci, vp, vp, vp, sz, vp]),
        "j2k_inverse_blocks_roi_general": (ci, [vp, IP, ci, ci, ci, vp, vp, vp, vp, vp, sz, vp]),
        "j2k_scatter_blocks_roi_general_device": (ci, [vp, ci, IP, ci, ci, ci, vp, vp, vp, vp, vp, vp]),
        "j2k_scatter_blocks_roi_device": (ci, [vp, ci, IP, ci, ci, ci, vp, vp, vp, vp]),
        "j2k_ht_decode_blocks": (ci, [vp, IP, ci, ci, ci, vp, sz, vp, vp, vp]),
        "j2k_inverse_ht": (ci, [vp, IP, ci, ci, ci, vp, sz, vp, vp, sz, vp, vp]),
        "j2k_submit_inverse_ht": (C.c_int64, [vp, IP, ci, ci, ci, vp, sz, vp, vp, sz, vp, vp]),
        "j2k_ht_decode_device": (ci, [vp, ci, IP, ci, ci, ci, vp, vp, vp, ci, vp, vp]),
        "j2k_ht_table": (ci, [ci, vp]),
        "j2k_forward_ht": (ci, [vp, FP, ci, ci, ci, vp, sz, vp, vp, sz, C.POINTER(sz), vp]),
        "j2k_ht_encode_bound": (sz, [FP, ci, ci, ci, ci]),
        "j2k_ht_encode_device": (ci, [vp, ci, FP, ci, ci, ci, vp, vp, vp, sz, vp, vp, vp]),
        "j2k_ht_enc_table": (ci, [ci, vp]),
        "j2k_dwt53_forward": (ci, [vp, vp, ci, ci, ci, ci, ci]),
        "j2k_dwt53_inverse": (ci, [vp, vp, ci, ci, ci, ci, ci]),
        "j2k_dwt97_forward": (ci, [vp, vp, ci, ci, ci, ci, ci]),
        "j2k_dwt97_inverse": (ci, [vp, vp, ci, ci, ci, ci, ci]),
        "j2k_convert_f32_to_i32": (ci, [vp, vp, vp, sz]),
        "j2k_dwt97_forward_f64": (ci, [vp, vp, ci, ci, ci, ci, ci]),
        "j2k_dwt97_inverse_f64": (ci, [vp, vp, ci, ci, ci, ci, ci]),
        "j2k_convert_f64_to_i32": (ci, [vp, vp, vp, sz]),
        "j2k_ll_dimensions": (ci, [ci, ci, ci, ci, ci, C.POINTER(ci), C.POINTER(ci)]),
        "j2k_rgb_to_ycbcr": (ci, [vp, vp, ci, ci, vp, vp, vp]),
        "j2k_ycbcr_to_rgb": (ci, [vp, vp, vp, vp, ci, ci, vp]),
        "j2k_interleave_components": (ci, [vp, C.POINTER(vp), ci, sz, vp]),
        "j2k_deinterleave_components": (ci, [vp, vp, sz, ci, C.POINTER(vp)]),
        "j2k_rct_forward": (ci, [vp, sz, vp, vp, vp, vp, vp, vp]),
        "j2k_rct_inverse": (ci, [vp, sz, vp, vp, vp, vp, vp, vp]),
        "j2k_ict_forward": (ci, [vp, sz, vp, vp, vp, vp, vp, vp]),
        "j2k_ict_inverse": (ci, [vp, sz, vp, vp, vp, vp, vp, vp]),
        "j2k_quantize_coefficients": (ci, [vp, vp, vp, sz, C.c_double]),
        "j2k_dequantize_coefficients": (ci, [vp, vp, vp, sz, C.c_double]),
        "j2k_quant_openjpeg_params": (ci, [ci, ci, u16p, f64p]),
        "j2k_quant_quality_params": (ci, [ci, ci, ci, u16p, f64p]),
        "j2k_quant_runtime_steps": (ci, [u16p, ci, ci, ci, f64p]),
        "j2k_quant_decode_steps": (ci, [u16p, ci, ci, ci, ci, f64p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here == a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _libs[path] = lib
    return lib
