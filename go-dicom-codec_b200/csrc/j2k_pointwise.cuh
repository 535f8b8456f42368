// j2k_pointwise.cuh — pointwise sm_100a kernels around the level kernels:
//   * prep_kernel      : raw / planar samples -> DC shift -> any forward MCT -> planar working planes
//                        (generic path: Part-2 custom MCT and bindings, planar API, zero-level transforms)
//   * finalize_kernel  : planar int32 samples -> any inverse MCT -> +DC -> clamp -> pack
//   * rect kernels     : per-sub-band quantization / dequantization driven by the literal band
//                        rectangles of bandInfosForResolution (irregular geometries, zero-level tiles)
//   * colour / quantizer API kernels (colorspace package, quantization.go exported functions)
#pragma once
#include "j2k_kernels.cuh"

namespace j2k {

struct MctOp {
    int n;              // components touched
    int ids[4];
    int kind;           // forward: 0 = integer matrix, 1 = Q13; inverse: 0 = integer matrix, 1 = float64 matrix
    int has_matrix;
    int mi[16];         // int32(m) or int32(m * 8192)
    double mf[16];
    int has_off;
    int off[4];
};
struct MctProgram {
    int kind;           // 0 none, 1 RCT, 2 ICT, 3 op list
    int n_ops;
    MctOp ops[4];
};

struct RectTable {      // sub-band rectangles in QCD order (encoder.go:2372-2389, t2/geometry.go:73-92)
    int n;
    int x[J2K_MAX_BANDS_K], y[J2K_MAX_BANDS_K], w[J2K_MAX_BANDS_K], h[J2K_MAX_BANDS_K];
    int mode[J2K_MAX_BANDS_K];
    float step[J2K_MAX_BANDS_K], rcp[J2K_MAX_BANDS_K], scale[J2K_MAX_BANDS_K];
};

// jpeg2000/encoder.go:662-665
__device__ __forceinline__ int mct_fixed_mul(int a, int b) {
    long long t = (long long)a * (long long)b + 4096;
    return (int)(t >> 13);
}

// One thread per pixel: forward head of Encoder.Encode (encoder.go:187-209) without the DWT.
// in_kind: IN_U8 / IN_U16 interleaved, IN_I32 planar (EncodeComponents, encoder.go:221-273).
// out_f32: ICT keeps float32 (encoder.go:206,277-288); everything else int32.
__global__ void __launch_bounds__(256) prep_kernel(const void* in, int in_kind, long long in_frame_stride, long long npix,
                                                   int C, RawFmt raw, const __grid_constant__ MctProgram prog, int* out,
                                                   long long out_frame_stride, int nframes) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= npix * nframes) return;
    long long f = gid / npix, i = gid - f * npix;
    int v[4];
    for (int c = 0; c < C; c++) {
        int s;
        if (in_kind == IN_U8) s = ((const unsigned char*)in)[f * in_frame_stride + i * C + c];
        else if (in_kind == IN_U16) s = ((const unsigned short*)in)[f * in_frame_stride + i * C + c];
        else s = ((const int*)in)[f * in_frame_stride + (long long)c * npix + i];
        if (in_kind != IN_I32) { if (s >= raw.sign_thresh) s -= raw.sign_sub; }
        v[c] = s - raw.dc;
    }
    int* o = out + f * out_frame_stride + i;
    if (prog.kind == 1) {
        int y, u, w;
        mct_forward<53, MCTK_RCT>(v[0], v[1], v[2], y, u, w);
        o[0] = y; o[npix] = u; o[2 * npix] = w;
        return;
    }
    if (prog.kind == 2) {
        float y, u, w;
        mct_forward<97, MCTK_ICT>(v[0], v[1], v[2], y, u, w);
        o[0] = __float_as_int(y); o[npix] = __float_as_int(u); o[2 * npix] = __float_as_int(w);
        return;
    }
    if (prog.kind == 3) {
        for (int k = 0; k < prog.n_ops; k++) {
            const MctOp& op = prog.ops[k];
            if (op.has_off) for (int j = 0; j < op.n; j++) v[op.ids[j]] -= op.off[j];  // encoder.go:481-491,582-594
            int res[4];
            for (int r = 0; r < op.n; r++) {
                if (op.kind == 0) {  // encoder.go:489-505,610-633
                    long long sum = 0;
                    for (int j = 0; j < op.n; j++) sum += (long long)op.mi[r * op.n + j] * (long long)v[op.ids[j]];
                    res[r] = (int)sum;
                } else {             // encoder.go:506-523,635-660
                    unsigned sum = 0;
                    for (int j = 0; j < op.n; j++) sum += (unsigned)mct_fixed_mul(op.mi[r * op.n + j], v[op.ids[j]]);
                    res[r] = (int)sum;
                }
            }
            for (int r = 0; r < op.n; r++) v[op.ids[r]] = res[r];
        }
    }
    for (int c = 0; c < C; c++) o[(long long)c * npix] = v[c];
}

// One thread per pixel: Decoder.applyInverseTransforms + applyInverseDCLevelShift + GetPixelData
// (decoder.go:540-542,620-735,777-962) on assembled planar int32 samples.
__global__ void __launch_bounds__(256) finalize_kernel(const int* in, long long in_frame_stride, long long npix, int C,
                                                       RawFmt raw, const __grid_constant__ MctProgram prog, void* pixels,
                                                       int out_kind, long long out_frame_stride, int* planes_out,
                                                       long long planes_frame_stride, int nframes) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= npix * nframes) return;
    long long f = gid / npix, i = gid - f * npix;
    int v[4];
    for (int c = 0; c < C; c++) v[c] = in[f * in_frame_stride + (long long)c * npix + i];
    if (prog.kind == 1) {
        int r, g, b; mct_inverse<MCTK_RCT>(v[0], v[1], v[2], r, g, b); v[0] = r; v[1] = g; v[2] = b;
    } else if (prog.kind == 2) {
        int r, g, b; mct_inverse<MCTK_ICT>(v[0], v[1], v[2], r, g, b); v[0] = r; v[1] = g; v[2] = b;
    } else if (prog.kind == 3) {
        for (int k = 0; k < prog.n_ops; k++) {
            const MctOp& op = prog.ops[k];
            if (op.has_matrix) {
                int res[4];
                for (int r = 0; r < op.n; r++) {
                    if (op.kind == 0) {  // decoder.go:646-662
                        long long sum = 0;
                        for (int j = 0; j < op.n; j++) sum += (long long)op.mi[r * op.n + j] * (long long)v[op.ids[j]];
                        res[r] = (int)sum;
                    } else {             // decoder.go:664-681,696-711: float64, no FMA, math.Round
                        double sum = 0.0;
                        for (int j = 0; j < op.n; j++) sum = __dadd_rn(sum, __dmul_rn(op.mf[r * op.n + j], (double)v[op.ids[j]]));
                        res[r] = (int)round(sum);
                    }
                }
                for (int r = 0; r < op.n; r++) v[op.ids[r]] = res[r];
            }
            if (op.has_off) for (int j = 0; j < op.n; j++) v[op.ids[j]] += op.off[j];  // decoder.go:683-694,712-722
        }
    }
    for (int c = 0; c < C; c++) {
        if (planes_out) planes_out[f * planes_frame_stride + (long long)c * npix + i] = v[c] + raw.dc;
        if (pixels) {
            int s = int_to_raw(v[c], raw);
            long long e = f * out_frame_stride + i * C + c;
            if (out_kind == IN_U16) ((unsigned short*)pixels)[e] = (unsigned short)s;
            else ((unsigned char*)pixels)[e] = (unsigned char)s;
        }
    }
}

__device__ __forceinline__ int rect_lookup(const RectTable& rt, int x, int y) {
    int hit = -1;
    for (int k = 0; k < rt.n; k++)  // later bands overwrite earlier ones, like the sequential Go loops
        if (x >= rt.x[k] && x < rt.x[k] + rt.w[k] && y >= rt.y[k] && y < rt.y[k] + rt.h[k]) hit = k;
    return hit;
}

// Forward finishing pass over Mallat planes that hold float32 bits (9/7, irregular geometry or
// zero performed levels): applyQuantizationBySubbandFloat (encoder.go:2265-2329) literally.
// all_round != 0: len(stepSizes) == 0 || NumLevels == 0 -> plain rounding (encoder.go:2266-2273).
__global__ void __launch_bounds__(256) quant_rects_kernel(int* planes, const long long* item_off, int n_items, int w, int h,
                                                          int row_stride, const __grid_constant__ RectTable rt, int all_round) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long per = (long long)w * h;
    if (gid >= per * n_items) return;
    int item = (int)(gid / per);
    int r = (int)(gid - (long long)item * per);
    int y = r / w, x = r - y * w;
    int* p = planes + item_off[item] + (long long)y * row_stride + x;
    float f = __int_as_float(*p);
    int q;
    if (all_round) q = __float2int_rn(f);
    else {
        int k = rect_lookup(rt, x, y);
        if (k < 0 || rt.mode[k] == 4) q = 0;  // `quantized` starts zeroed (encoder.go:2275)
        else if (rt.mode[k] == Q_QUANT) q = __float2int_rn(__fmul_rn(div_by_step(f, rt.step[k], rt.rcp[k]), rt.scale[k]));
        else q = __float2int_rn(f);
    }
    *p = q;
}

// 5/3 finishing pass: classic-EBCOT `<<6` (encoder.go:3294-3300) over whole planes.
__global__ void __launch_bounds__(256) shift_planes_kernel(int* planes, const long long* item_off, int n_items, int w, int h,
                                                           int row_stride, int shift) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long per = (long long)w * h;
    if (gid >= per * n_items) return;
    int item = (int)(gid / per);
    int r = (int)(gid - (long long)item * per);
    int y = r / w, x = r - y * w;
    int* p = planes + item_off[item] + (long long)y * row_stride + x;
    *p = (int)((unsigned)*p << shift);
}

// Copies a w x h window between two item-addressed planar buffers (tile extraction / assembly,
// encoder.go:2213-2237, tile_assembler.go:164-175).  cvt: 0 bit copy, 1 int32 -> float32 bits
// (ConvertInt32ToFloat32, dwt97.go:463-469), 2 float32 bits -> rounded int32 (dwt97.go:473-479),
// 3 int32 truncating /2 (t2/tile_decoder.go:989-993).
__global__ void __launch_bounds__(256) copy_window_kernel(const int* src, const long long* src_off, int src_stride, int* dst,
                                                          const long long* dst_off, int dst_stride, int n_items, int w, int h,
                                                          int cvt) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long per = (long long)w * h;
    if (gid >= per * n_items) return;
    int item = (int)(gid / per);
    int r = (int)(gid - (long long)item * per);
    int y = r / w, x = r - y * w;
    int v = src[src_off[item] + (long long)y * src_stride + x];
    if (cvt == 1) v = __float_as_int((float)v);
    else if (cvt == 2) v = __float2int_rn(__int_as_float(v));
    else if (cvt == 3) v = v / 2;
    dst[dst_off[item] + (long long)y * dst_stride + x] = v;
}

// Inverse pre-pass: applyDequantizationBySubbandFloat (t2/tile_decoder.go:925-987) literally, into a
// float32-bits copy of the coefficient planes (irregular geometry or zero performed levels).
__global__ void __launch_bounds__(256) dequant_rects_kernel(const int* coeffs, const long long* src_off, int src_stride, int* dst,
                                                            const long long* dst_off, int dst_stride, int n_items, int w, int h,
                                                            const __grid_constant__ RectTable rt) {
    long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long per = (long long)w * h;
    if (gid >= per * n_items) return;
    int item = (int)(gid / per);
    int r = (int)(gid - (long long)item * per);
    int y = r / w, x = r - y * w;
    float f = (float)coeffs[src_off[item] + (long long)y * src_stride + x];
    // bands never overlap, so at most one rectangle scales a sample
    int k = rect_lookup(rt, x, y);
    if (k >= 0 && rt.mode[k] == DQ_SCALE) f = __fmul_rn(f, rt.scale[k]);
    dst[dst_off[item] + (long long)y * dst_stride + x] = __float_as_int(f);
}

// ---- colorspace package API (colorspace/rct.go:26-49, colorspace/ict.go:24-45)
__global__ void __launch_bounds__(256) color_api_kernel(int op, long long n, const int* a, const int* b, const int* c, int* x,
                                                        int* y, int* z) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int A = a[i], B = b[i], Cc = c[i], X, Y, Z;
    if (op == 0) { mct_forward<53, MCTK_RCT>(A, B, Cc, X, Y, Z); }
    else if (op == 1) { mct_inverse<MCTK_RCT>(A, B, Cc, X, Y, Z); }
    else if (op == 2) {  // ICTForward: float64 + math.Round (ict.go:8-14)
        double R = (double)A, G = (double)B, Bl = (double)Cc;
        X = (int)round(__dadd_rn(__dadd_rn(__dmul_rn(0.299, R), __dmul_rn(0.587, G)), __dmul_rn(0.114, Bl)));
        Y = (int)round(__dadd_rn(__dadd_rn(__dmul_rn(-0.16875, R), -__dmul_rn(0.331260, G)), __dmul_rn(0.5, Bl)));
        Z = (int)round(__dadd_rn(__dadd_rn(__dmul_rn(0.5, R), -__dmul_rn(0.41869, G)), -__dmul_rn(0.08131, Bl)));
    } else { mct_inverse<MCTK_ICT>(A, B, Cc, X, Y, Z); }
    x[i] = X; y[i] = Y; z[i] = Z;
}

// ---- quantization.go exported API (:310-340): RoundToEven(float64(c) / step), RoundToEven(float64(c) * step)
__global__ void __launch_bounds__(256) quant_api_kernel(int op, long long n, const int* in, int* out, double step) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double v = (double)in[i];
    v = op == 0 ? __ddiv_rn(v, step) : __dmul_rn(v, step);
    out[i] = (int)rint(v);
}

// ---- wavelet.ConvertFloat32ToInt32OpenJPEG (dwt97.go:473-503)
__global__ void __launch_bounds__(256) f32_to_i32_kernel(const float* in, int* out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = in[i];
    int r = __float2int_rn(v);
    if (!(fabsf(v) < 2147483648.0f)) {
        // outside int32 the reference truncates its int64 result (integers all: |v| >= 2^31), and beyond int64 the
        // amd64 "integer indefinite" 0x8000000000000000 flows through the remaining arithmetic of dwt97.go:483-503
        if (fabsf(v) < 9223372036854775808.0f) r = (int)(unsigned)(unsigned long long)__float2ll_rz(v);
        else r = (v != v) ? 0 : (v > 0.f ? 1 : -1);
    }
    out[i] = r;
}

// ---- float64 wrappers of the wavelet package (dwt97.go:30-44,181-187,249-261,325-351,410-421): they convert to float32,
// run the float32 transform and convert back; ConvertFloat64ToInt32 (dwt97.go:515-526) rounds half away from zero by
// truncating v +- 0.5.
__global__ void __launch_bounds__(256) f64_to_f32_kernel(const double* in, float* out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (float)in[i];
}
__global__ void __launch_bounds__(256) f32_to_f64_kernel(const float* in, double* out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}
__global__ void __launch_bounds__(256) f64_to_i32_kernel(const double* in, int* out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = in[i];
    const double t = v >= 0 ? __dadd_rn(v, 0.5) : __dadd_rn(v, -0.5);
    // Go's float64 -> int32 conversion of an out-of-range value (or NaN) on amd64 is the "integer indefinite" 0x80000000
    out[i] = (t > -2147483649.0 && t < 2147483648.0) ? (int)t : (int)0x80000000;
}

// ---- colorspace.InterleaveComponents / DeinterleaveComponents (rgb.go:54-98): planes [C][n] <-> interleaved [n][C]
__global__ void __launch_bounds__(256) interleave_kernel(const int* planes, int* out, long long n, int C, int to_interleaved) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // index into the interleaved array
    if (i >= n * C) return;
    const long long p = i / C;
    const int c = (int)(i - p * C);
    if (to_interleaved) out[i] = planes[(long long)c * n + p];
    else out[(long long)c * n + p] = planes[i];
}

// ---- code-block interface: plane (Mallat layout) <-> block-major plane
//
// gather = getSubbandsForResolution + partitionIntoCodeBlocks + codeBlockNumBps (jpeg2000/encoder.go:3059-3285,
// 3349-3362, 3643-3667) in one pass: ONE WARP per code-block copies its rows (128-bit vectors when the block is
// 4-aligned) into the contiguous block and reduces max |v| with shuffles; lane 0 writes cblkNumbps.  scatter =
// TileDecoder.assembleSubbands (jpeg2000/t2/tile_decoder.go:840-883).  Both move 8 B per sample: HBM-bound.
struct BlockEntry {
    long long plane_off;   // first sample of the block inside the frame's coefficient planes (samples)
    long long block_off;   // first sample of the block inside the frame's block-major planes
    int stride, w, h, vec; // plane row stride (= tile width); vec: w % 4 == 0 and both sides 16-byte aligned in every frame
    int comp, pad_;        // component of the block (per-component ROI shift)
};

// Srgn = 0 (MaxShift) ROI shifts per component, 0 = none (t2/tile_decoder.go:726-730)
#define J2K_ROI_MAXC 16
struct RoiShifts {
    int any; int shift[J2K_ROI_MAXC];
    // general scaling (RGN Srgn = 1, t2/tile_decoder.go:735-742): per (frame, block) shift, 0 = block outside the region;
    // optional per-sample mask in block-major order (non-zero = the sample is divided), NULL = the whole block
    const int* block_shift;
    const unsigned char* sample_mask;
};

// applyInverseGeneralScaling (t2/tile_decoder.go:1082-1090): data /= 2^shift, Go's truncating division.  It follows the
// classic 5/3 "/2" in decodeCodeBlock; truncating divisions by positive powers of two commute, so it may run before it.
__device__ __forceinline__ int inverse_general_scaling(int v, int shift) { return shift > 0 ? v / (1 << shift) : v; }

// applyInverseMaxShift (t2/tile_decoder.go:1113-1138): magnitudes at or above 2^shift belong to the ROI and come down by
// `shift`; shift >= 31 zeroes the block.  -mag wraps for INT_MIN as in Go (then mag < thresh: untouched).
__device__ __forceinline__ int inverse_max_shift(int v, int shift) {
    if (shift <= 0) return v;
    if (shift >= 31) return 0;
    int mag = v < 0 ? (int)(0u - (unsigned)v) : v;
    if (mag >= (1 << shift)) {
        mag >>= shift;
        return v < 0 ? -mag : mag;
    }
    return v;
}

__device__ __forceinline__ int go_abs_max(int m, int v) {  // calculateMaxBitplane: abs wraps for INT_MIN, compare is signed
    const int a = v < 0 ? (int)(0u - (unsigned)v) : v;
    return a > m ? a : m;
}

__global__ void __launch_bounds__(128) gather_blocks_kernel(const int* __restrict__ coeffs, long long coeffs_per_frame,
                                                            const BlockEntry* __restrict__ tab, int nblocks, long long total,
                                                            int* __restrict__ blocks, int* __restrict__ numbps, int sub6, int vec_ok) {
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= total) return;
    const long long frame = wid / nblocks;
    const int b = (int)(wid - frame * nblocks);
    const BlockEntry e = tab[b];
    const int* src = coeffs + frame * coeffs_per_frame + e.plane_off;
    int* dst = blocks + frame * coeffs_per_frame + e.block_off;
    int m = 0;
    if (e.vec && vec_ok) {
        const int w4 = e.w >> 2, n4 = w4 * e.h;
#pragma unroll 4
        for (int i = lane; i < n4; i += 32) {
            const int y = i / w4, x = i - y * w4;
            const int4 v = *(const int4*)(src + (long long)y * e.stride + 4 * x);
            m = go_abs_max(go_abs_max(go_abs_max(go_abs_max(m, v.x), v.y), v.z), v.w);
            *(int4*)(dst + 4 * i) = v;
        }
    } else {
        const int n = e.w * e.h;
        for (int i = lane; i < n; i += 32) {
            const int y = i / e.w, x = i - y * e.w;
            const int v = src[(long long)y * e.stride + x];
            m = go_abs_max(m, v);
            dst[i] = v;
        }
    }
    if (numbps) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const int o = __shfl_xor_sync(0xffffffffu, m, d);
            m = o > m ? o : m;
        }
        if (lane == 0) {
            int bits = 0;  // rawMaxBitplane + 1
            for (unsigned t = (unsigned)m; t; t >>= 1) bits++;
            int v = m == 0 ? 0 : bits - (sub6 ? 6 : 0);  // codeBlockNumBps: minus t1NMSEDecFracBits unless HTJ2K
            numbps[frame * nblocks + b] = v < 0 ? 0 : v;
        }
    }
}

__global__ void __launch_bounds__(128) scatter_blocks_kernel(const int* __restrict__ blocks, long long coeffs_per_frame,
                                                             const BlockEntry* __restrict__ tab, int nblocks, long long total,
                                                             int* __restrict__ coeffs, const RoiShifts roi, int vec_ok) {
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= total) return;
    const long long frame = wid / nblocks;
    const int b = (int)(wid - frame * nblocks);
    const BlockEntry e = tab[b];
    const int* src = blocks + frame * coeffs_per_frame + e.block_off;
    int* dst = coeffs + frame * coeffs_per_frame + e.plane_off;
    const int sh = roi.any ? roi.shift[e.comp] : 0;  // warp-uniform
    const int gs = roi.block_shift ? roi.block_shift[frame * nblocks + b] : 0;  // warp-uniform: general scaling of this block
    if (gs > 0) {  // the rare path: one sample at a time, masked or whole block
        const unsigned char* mk = roi.sample_mask ? roi.sample_mask + frame * coeffs_per_frame + e.block_off : nullptr;
        const int n = e.w * e.h;
        for (int i = lane; i < n; i += 32) {
            const int y = i / e.w, x = i - y * e.w;
            int v = inverse_max_shift(src[i], sh);
            if (!mk || mk[i]) v = inverse_general_scaling(v, gs);
            dst[(long long)y * e.stride + x] = v;
        }
        return;
    }
    if (e.vec && vec_ok) {
        const int w4 = e.w >> 2, n4 = w4 * e.h;
#pragma unroll 4
        for (int i = lane; i < n4; i += 32) {
            const int y = i / w4, x = i - y * w4;
            int4 v = *(const int4*)(src + 4 * i);
            if (sh) { v.x = inverse_max_shift(v.x, sh); v.y = inverse_max_shift(v.y, sh); v.z = inverse_max_shift(v.z, sh); v.w = inverse_max_shift(v.w, sh); }
            *(int4*)(dst + (long long)y * e.stride + 4 * x) = v;
        }
    } else {
        const int n = e.w * e.h;
        for (int i = lane; i < n; i += 32) {
            const int y = i / e.w, x = i - y * e.w;
            dst[(long long)y * e.stride + x] = inverse_max_shift(src[i], sh);
        }
    }
}

}  // namespace j2k
