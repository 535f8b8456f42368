// j2k_kernels.cuh — sm_100a kernels of the JPEG 2000 sample-domain path.
//
// Design: every DWT level is ONE kernel that fuses the vertical and the
// horizontal lifting pass.  A warp owns a strip of the level window: 32 lanes x NP sample
// pairs wide, CH row pairs tall (+ halo).  It streams down the strip two rows per iteration:
//   * forward : rows are loaded straight from HBM with coalesced (128-bit when aligned) loads,
//               run through a register sliding-window vertical lifting, and each finished
//               (low row, high row) pair is lifted horizontally with warp shuffles for the
//               lane-crossing neighbours, quantized and stored to its four sub-bands;
//   * inverse : the mirror image (horizontal synthesis with shuffles, then the vertical
//               sliding window, then rounding / colour transform / packing).
// Borders use whole-sample symmetric index mirroring, which tests/test_oracle_mirror.py
// proves bit-identical to the reference's border special cases
// (jpeg2000/wavelet/dwt53.go:27-234, jpeg2000/wavelet/dwt97.go:47-287).
// All float arithmetic uses __fadd_rn/__fmul_rn (never contracted into FMA): the parity
// target is Go on amd64, where every float op is individually rounded.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace j2k {

#define J2K_MAX_BANDS_K 31  // 3 * J2K_MAX_LEVELS + 1 (include/j2k_b200.h)

enum InKind : int { IN_U8 = 0, IN_U16 = 1, IN_I32 = 2, IN_F32 = 3 };
enum QMode : int {
    Q_RAW = 0,    // store the working value as is (int32 for 5/3, float32 bits for 9/7)
    Q_SHIFT = 1,  // 5/3: value << shift                     (encoder.go:3294-3300)
    Q_ROUND = 2,  // 9/7: round half even                    (encoder.go:2320-2321)
    Q_QUANT = 3,  // 9/7: rint((c / step) * scale)           (encoder.go:2323-2324)
    Q_ZERO = 4    // band index beyond len(stepSizes): `quantized` stays zero (encoder.go:2275,2283,2294)
};
enum DqMode : int {
    DQ_RAW = 0,    // working type already (int32 for 5/3, float32 for 9/7)
    DQ_HALVE = 1,  // 5/3: truncating /2                     (t2/tile_decoder.go:989-993)
    DQ_CVT = 2,    // 9/7: float32(q)                        (dwt97.go:463-469)
    DQ_SCALE = 3   // 9/7: float32(q) * float32(scale)       (t2/tile_decoder.go:970-987)
};
enum MctKind : int { MCTK_NONE = 0, MCTK_RCT = 1, MCTK_ICT = 2 };

struct BandIO {
    void* base;            // int32_t* or float*
    const long long* off;  // per-item element offset of the band's plane origin
    long long comp_stride; // elements between the components of one item (NC = 3 kernels)
    int row_stride;        // elements
    int x_off, y_off;      // band origin inside the plane
    int mode;              // QMode (forward) / DqMode (inverse)
    int shift;
    float step, rcp, scale;
};

struct RawFmt {            // interleaved pixel words <-> level-shifted integers
    int pix_stride;        // samples between horizontally adjacent pixels of one component (= components)
    int sign_thresh;       // is_signed: v >= thresh -> v -= sign_sub   (encoder.go:362-364,374-376)
    int sign_sub;
    int dc;                // unsigned: 2^(B-1)                          (encoder.go:3698-3711)
    int clamp_lo, clamp_hi;// decoder.go:787-944
    int wrap_add;          // signed: v < 0 -> v += 2^B
};

struct LevelArgs {
    // level window
    int w, h, px, py, lw, lh, Kx, Ky, hskip, vskip;
    // work decomposition
    int n_items, nchunks, nstrips, chunk_pairs;
    // the interleaved side: forward input / inverse output
    void* x_base;
    const long long* x_off;   // per item, in elements
    long long x_comp_stride;  // NC = 3 planar sources
    int x_row_stride;
    int x_kind;               // InKind
    int x_mode;               // inverse, planar store: 0 = working-type bits, 1 = 9/7 samples rounded half-even to int32
    int vec_x, vec_b;         // 128-bit fast paths allowed on the interleaved / band side
    RawFmt raw;
    int32_t* planes_out;      // inverse final: optional GetImageData planes (decoder.go:738-740)
    const long long* planes_off;
    long long planes_comp_stride;
    int planes_row_stride;
    BandIO ll, hl, lh_, hh;
};

// ------------------------------------------------------------------ helpers

__device__ __forceinline__ int mirror_idx(int i, int n) {
    if ((unsigned)i < (unsigned)n) return i;
    if (n == 1) return 0;
    int period = 2 * (n - 1);
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - i;
}

// 9/7 constants: Go float64 literal -> float32 (jpeg2000/wavelet/dwt97.go:11-22,99,207-212)
#define J2K_ALPHA ((float)-1.586134342)
#define J2K_BETA ((float)-0.052980118)
#define J2K_GAMMA ((float)0.882911075)
#define J2K_DELTA ((float)0.443506852)
#define J2K_K ((float)1.230174105)
#define J2K_INVK ((float)0.812893066)
#define J2K_TWOINVK ((float)1.625732422)

template <int WT> struct Wt;
template <> struct Wt<53> {
    typedef int T;
    static constexpr int LAG = 1, HALO = 1;
};
template <> struct Wt<97> {
    typedef float T;
    static constexpr int LAG = 2, HALO = 2;
};

// x + (l + r) * c with three roundings (dwt97.go:104-116)
__device__ __forceinline__ float lift97(float x, float l, float r, float c) {
    return __fadd_rn(x, __fmul_rn(__fadd_rn(l, r), c));
}

// IEEE-correct c / step from a correctly rounded reciprocal (Markstein): q0 = c*r, e = c - q0*step
// exactly (fma), q = q0 + e*r.  Outside the safe exponent range, and for steps the host marks with rcp == 0
// (markstein_safe in j2k_b200.cu: all-ones significand, extreme exponents), it is the true division.
__device__ __forceinline__ float div_by_step(float c, float step, float rcp) {
    float a = fabsf(c);
    if (rcp != 0.f && a > 1e-30f && a < 1e30f) {
        float q0 = __fmul_rn(c, rcp);
        float e = __fmaf_rn(-q0, step, c);
        return __fmaf_rn(e, rcp, q0);
    }
    return __fdiv_rn(c, step);
}

template <int WT>
__device__ __forceinline__ int quant_store_value(typename Wt<WT>::T v, const BandIO& b) {
    if constexpr (WT == 53) {
        int iv = (int)v;
        return b.mode == Q_SHIFT ? (int)((unsigned)iv << b.shift) : iv;
    } else {
        float f = (float)v;
        if (b.mode == Q_RAW) return __float_as_int(f);
        if (b.mode == Q_ZERO) return 0;
        if (b.mode == Q_QUANT) f = __fmul_rn(div_by_step(f, b.step, b.rcp), b.scale);
        return __float2int_rn(f);  // round half even == Go RoundToEven(float64(f)) for |f| < 2^31
    }
}

template <int WT>
__device__ __forceinline__ typename Wt<WT>::T dequant_load_value(int raw, const BandIO& b) {
    if constexpr (WT == 53) {
        return (typename Wt<WT>::T)(b.mode == DQ_HALVE ? raw / 2 : raw);
    } else {
        if (b.mode == DQ_RAW) return (typename Wt<WT>::T)__int_as_float(raw);
        float f = (float)raw;
        if (b.mode == DQ_SCALE) f = __fmul_rn(f, b.scale);
        return (typename Wt<WT>::T)f;
    }
}

// raw pixel word -> level-shifted integer (convertPixelData + applyDCLevelShift)
__device__ __forceinline__ int raw_to_int(int v, const RawFmt& r) {
    if (v >= r.sign_thresh) v -= r.sign_sub;
    return v - r.dc;
}
// sample -> stored pixel word (applyInverseDCLevelShift + GetPixelData)
__device__ __forceinline__ int int_to_raw(int v, const RawFmt& r) {
    v += r.dc;
    v = max(r.clamp_lo, min(r.clamp_hi, v));
    if (v < 0) v += r.wrap_add;
    return v;
}

// Forward MCT on level-shifted integers; produces the DWT working type.
template <int WT, int MCT>
__device__ __forceinline__ void mct_forward(int r, int g, int b, typename Wt<WT>::T& y, typename Wt<WT>::T& u,
                                            typename Wt<WT>::T& v) {
    if constexpr (MCT == MCTK_RCT) {  // colorspace/rct.go:6-11
        y = (typename Wt<WT>::T)((r + 2 * g + b) >> 2);
        u = (typename Wt<WT>::T)(b - g);
        v = (typename Wt<WT>::T)(r - g);
    } else if constexpr (MCT == MCTK_ICT) {  // encoder.go:277-288 (float32, result kept float32)
        float fr = (float)r, fg = (float)g, fb = (float)b;
        y = (typename Wt<WT>::T)__fadd_rn(__fadd_rn(__fmul_rn(fr, 0.299f), __fmul_rn(fg, 0.587f)), __fmul_rn(fb, 0.114f));
        u = (typename Wt<WT>::T)__fadd_rn(__fadd_rn(__fmul_rn(fr, -0.16875f), __fmul_rn(fg, -0.331260f)), __fmul_rn(fb, 0.5f));
        v = (typename Wt<WT>::T)__fadd_rn(__fadd_rn(__fmul_rn(fr, 0.5f), __fmul_rn(fg, -0.41869f)), __fmul_rn(fb, -0.08131f));
    } else {
        y = (typename Wt<WT>::T)r; u = (typename Wt<WT>::T)g; v = (typename Wt<WT>::T)b;
    }
}

// Inverse MCT on rounded integer samples (decoder.go:725-735).
template <int MCT>
__device__ __forceinline__ void mct_inverse(int y, int u, int v, int& r, int& g, int& b) {
    if constexpr (MCT == MCTK_RCT) {  // colorspace/rct.go:16-21
        g = y - ((u + v) >> 2);
        r = v + g;
        b = u + g;
    } else if constexpr (MCT == MCTK_ICT) {  // colorspace/ict.go:16-21: float64, math.Round (half away from zero)
        double dy = (double)y, du = (double)u, dv = (double)v;
        r = (int)round(__dadd_rn(dy, __dmul_rn(1.402, dv)));
        g = (int)round(__dadd_rn(__dadd_rn(dy, -__dmul_rn(0.34413, du)), -__dmul_rn(0.71414, dv)));
        b = (int)round(__dadd_rn(dy, __dmul_rn(1.772, du)));
    } else {
        r = y; g = u; b = v;
    }
}

// ------------------------------------------------------------------ forward level kernel

template <int WT, int NP, int NC, int IN, int MCT>
struct FwdLevel {
    typedef typename Wt<WT>::T T;
    static constexpr int LAG = Wt<WT>::LAG;
    static constexpr int HLN = (Wt<WT>::HALO + NP - 1) / NP;  // halo lanes per warp side
    static constexpr int VP = (32 - 2 * HLN) * NP;            // valid pairs per warp strip
    static constexpr int NS = 2 * NP;                         // samples per lane per row

    // Loads the lane's NS samples of interleaved row `row` (already mirrored) for all NC components,
    // converts raw words, applies DC shift and the forward MCT.
    static __device__ __forceinline__ void load_row(const LevelArgs& a, long long item_off, int row, int i0, bool fast,
                                                    T (&out)[NC][NS]) {
        int vals[NC][NS];
        if constexpr (IN == IN_U8 || IN == IN_U16) {
            const int ps = a.raw.pix_stride;
            long long rbase = item_off + (long long)row * a.x_row_stride;
            if (fast) {
                // contiguous words: NS pixels x ps samples starting at pixel i0
                if constexpr (NC == 1) {
                    if constexpr (IN == IN_U16) {
                        const unsigned short* p = (const unsigned short*)a.x_base + rbase + i0;
                        if constexpr (NS == 8) {
                            uint4 q = __ldg((const uint4*)p);
                            unsigned w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                            for (int k = 0; k < 4; k++) { vals[0][2 * k] = w4[k] & 0xFFFF; vals[0][2 * k + 1] = w4[k] >> 16; }
                        } else if constexpr (NS == 4) {
                            uint2 q = __ldg((const uint2*)p);
                            vals[0][0] = q.x & 0xFFFF; vals[0][1] = q.x >> 16; vals[0][2] = q.y & 0xFFFF; vals[0][3] = q.y >> 16;
                        } else {
                            unsigned q = __ldg((const unsigned*)p);
                            vals[0][0] = q & 0xFFFF; vals[0][1] = q >> 16;
                        }
                    } else {
                        const unsigned char* p = (const unsigned char*)a.x_base + rbase + i0;
                        if constexpr (NS == 8) {
                            uint2 q = __ldg((const uint2*)p);
                            unsigned w2[2] = {q.x, q.y};
#pragma unroll
                            for (int k = 0; k < 8; k++) vals[0][k] = (w2[k >> 2] >> (8 * (k & 3))) & 0xFF;
                        } else if constexpr (NS == 4) {
                            unsigned q = __ldg((const unsigned*)p);
#pragma unroll
                            for (int k = 0; k < 4; k++) vals[0][k] = (q >> (8 * k)) & 0xFF;
                        } else {
                            unsigned short q = __ldg((const unsigned short*)p);
                            vals[0][0] = q & 0xFF; vals[0][1] = q >> 8;
                        }
                    }
                } else {  // NC == 3 interleaved, contiguous 3*NS words
                    if constexpr (IN == IN_U8) {
                        const unsigned* p = (const unsigned*)((const unsigned char*)a.x_base + rbase + (long long)i0 * 3);
                        constexpr int NW = (3 * NS) / 4;  // NS multiple of 4
                        unsigned wv[NW];
#pragma unroll
                        for (int k = 0; k < NW; k++) wv[k] = __ldg(p + k);
#pragma unroll
                        for (int s = 0; s < NS; s++)
#pragma unroll
                            for (int c = 0; c < 3; c++) {
                                int bi = 3 * s + c;
                                vals[c][s] = (wv[bi >> 2] >> (8 * (bi & 3))) & 0xFF;
                            }
                    } else {
                        const uint2* p = (const uint2*)((const unsigned short*)a.x_base + rbase + (long long)i0 * 3);
                        constexpr int NW = (3 * NS) / 4;  // uint2 = 4 words
                        uint2 wv[NW];
#pragma unroll
                        for (int k = 0; k < NW; k++) wv[k] = __ldg(p + k);
#pragma unroll
                        for (int s = 0; s < NS; s++)
#pragma unroll
                            for (int c = 0; c < 3; c++) {
                                int wi = 3 * s + c;
                                unsigned half = (wi & 2) ? wv[wi >> 2].y : wv[wi >> 2].x;
                                vals[c][s] = (wi & 1) ? (half >> 16) : (half & 0xFFFF);
                            }
                    }
                }
            } else {
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    int xi = mirror_idx(i0 + s, a.w);
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        long long e = rbase + (long long)xi * ps + c;
                        vals[c][s] = (IN == IN_U16) ? (int)__ldg((const unsigned short*)a.x_base + e)
                                                    : (int)__ldg((const unsigned char*)a.x_base + e);
                    }
                }
            }
#pragma unroll
            for (int c = 0; c < NC; c++)
#pragma unroll
                for (int s = 0; s < NS; s++) vals[c][s] = raw_to_int(vals[c][s], a.raw);
        } else {  // planar int32 / float32 words
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int* p = (const int*)a.x_base + item_off + c * a.x_comp_stride + (long long)row * a.x_row_stride;
                if (fast) {
                    if constexpr (NS == 4) {
                        int4 q = __ldg((const int4*)(p + i0));
                        vals[c][0] = q.x; vals[c][1] = q.y; vals[c][2] = q.z; vals[c][3] = q.w;
                    } else if constexpr (NS == 8) {
                        int4 q = __ldg((const int4*)(p + i0));
                        int4 q2 = __ldg((const int4*)(p + i0 + 4));
                        vals[c][0] = q.x; vals[c][1] = q.y; vals[c][2] = q.z; vals[c][3] = q.w;
                        vals[c][4] = q2.x; vals[c][5] = q2.y; vals[c][6] = q2.z; vals[c][7] = q2.w;
                    } else {
                        int2 q = __ldg((const int2*)(p + i0));
                        vals[c][0] = q.x; vals[c][1] = q.y;
                    }
                } else {
#pragma unroll
                    for (int s = 0; s < NS; s++) vals[c][s] = __ldg(p + mirror_idx(i0 + s, a.w));
                }
                if constexpr (IN == IN_I32) {
#pragma unroll
                    for (int s = 0; s < NS; s++) vals[c][s] -= a.raw.dc;
                }
            }
        }
        // to working type (+ MCT)
#pragma unroll
        for (int s = 0; s < NS; s++) {
            if constexpr (IN == IN_F32) {
#pragma unroll
                for (int c = 0; c < NC; c++) out[c][s] = (T)__int_as_float(vals[c][s]);
            } else if constexpr (NC == 3 && MCT != MCTK_NONE) {
                mct_forward<WT, MCT>(vals[0][s], vals[1][s], vals[2][s], out[0][s], out[1][s], out[2][s]);
            } else {
#pragma unroll
                for (int c = 0; c < NC; c++) out[c][s] = (T)vals[c][s];
            }
        }
    }

    // Horizontal analysis of one row held across the warp (lane: NP pairs E/O), in place:
    // on return v[2j] = low_j, v[2j+1] = high_j.
    static __device__ __forceinline__ void hlift(T (&v)[NS], bool hskip) {
        if (hskip) return;
        if constexpr (WT == 97) {
            float e[NP + 1], o[NP + 1];  // o[0] = previous lane's last
#pragma unroll
            for (int j = 0; j < NP; j++) { e[j] = (float)v[2 * j]; o[j + 1] = (float)v[2 * j + 1]; }
            e[NP] = __shfl_down_sync(0xffffffffu, e[0], 1);
#pragma unroll
            for (int j = 0; j < NP; j++) o[j + 1] = lift97(o[j + 1], e[j], e[j + 1], J2K_ALPHA);
            o[0] = __shfl_up_sync(0xffffffffu, o[NP], 1);
#pragma unroll
            for (int j = 0; j < NP; j++) e[j] = lift97(e[j], o[j], o[j + 1], J2K_BETA);
            e[NP] = __shfl_down_sync(0xffffffffu, e[0], 1);
#pragma unroll
            for (int j = 0; j < NP; j++) o[j + 1] = lift97(o[j + 1], e[j], e[j + 1], J2K_GAMMA);
            o[0] = __shfl_up_sync(0xffffffffu, o[NP], 1);
#pragma unroll
            for (int j = 0; j < NP; j++) e[j] = lift97(e[j], o[j], o[j + 1], J2K_DELTA);
#pragma unroll
            for (int j = 0; j < NP; j++) {
                v[2 * j] = (T)__fmul_rn(e[j], J2K_INVK);
                v[2 * j + 1] = (T)__fmul_rn(o[j + 1], J2K_K);
            }
        } else {
            int e[NP + 1], o[NP + 1];
#pragma unroll
            for (int j = 0; j < NP; j++) { e[j] = (int)v[2 * j]; o[j + 1] = (int)v[2 * j + 1]; }
            e[NP] = __shfl_down_sync(0xffffffffu, e[0], 1);
#pragma unroll
            for (int j = 0; j < NP; j++) o[j + 1] = o[j + 1] - ((e[j] + e[j + 1]) >> 1);
            o[0] = __shfl_up_sync(0xffffffffu, o[NP], 1);
#pragma unroll
            for (int j = 0; j < NP; j++) e[j] = e[j] + ((o[j] + o[j + 1] + 2) >> 2);
#pragma unroll
            for (int j = 0; j < NP; j++) { v[2 * j] = (T)e[j]; v[2 * j + 1] = (T)o[j + 1]; }
        }
    }

    // Stores the lane's NP values `val[j]` of one band row; xi0 = band x index of j = 0.
    static __device__ __forceinline__ void store_band(const BandIO& b, long long item_off, int c, int yrow, int xi0, int bw,
                                                      const T (&val)[NP], bool vec) {
        int* p = (int*)b.base + item_off + c * b.comp_stride + (long long)(b.y_off + yrow) * b.row_stride + b.x_off;
        int q[NP];
#pragma unroll
        for (int j = 0; j < NP; j++) q[j] = quant_store_value<WT>(val[j], b);
        if (vec && xi0 >= 0 && xi0 + NP <= bw) {
            if constexpr (NP == 4) { *(int4*)(p + xi0) = make_int4(q[0], q[1], q[2], q[3]); return; }
            else if constexpr (NP == 2) { *(int2*)(p + xi0) = make_int2(q[0], q[1]); return; }
        }
#pragma unroll
        for (int j = 0; j < NP; j++)
            if (xi0 + j >= 0 && xi0 + j < bw) p[xi0 + j] = q[j];
    }

    static __device__ __forceinline__ void run(const LevelArgs& a) {
        const int lane = threadIdx.x & 31;
        const long long job = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const long long njobs = (long long)a.n_items * a.nchunks * a.nstrips;
        if (job >= njobs) return;
        const int strip = (int)(job % a.nstrips);
        const int chunk = (int)((job / a.nstrips) % a.nchunks);
        const int item = (int)(job / ((long long)a.nstrips * a.nchunks));

        const int kx0 = strip * VP - HLN * NP + lane * NP;  // first pair of this lane
        const int i0 = 2 * kx0 - a.px;                       // its first sample index
        const bool lane_out = lane >= HLN && lane < 32 - HLN;
        const bool fast_x = a.vec_x && i0 >= 0 && i0 + NS <= a.w;
        const long long x_off = a.x_off[item];

        const int ky0 = chunk * a.chunk_pairs;
        const int ky1 = min(ky0 + a.chunk_pairs, a.Ky);
        const int hw = a.w - a.lw, hh = a.h - a.lh;
        const long long ll_off = a.ll.off[item], hl_off = a.hl.off[item], lh_off = a.lh_.off[item], hh_off = a.hh.off[item];
        const int xl0 = kx0 - a.px;  // low-band x index of pair j = 0; high-band index is kx0

        // vertical sliding-window state, per component and column
        T pe[NC][NS], po[NC][NS], s1p[NC][NS], d1p[NC][NS], d2p[NC][NS];
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
            for (int s = 0; s < NS; s++) { pe[c][s] = 0; po[c][s] = 0; s1p[c][s] = 0; d1p[c][s] = 0; d2p[c][s] = 0; }

        const int lag = a.vskip ? 0 : LAG;
        for (int t = ky0 - lag; t < ky1 + lag; t++) {
            T e[NC][NS], o[NC][NS];
            const int re = mirror_idx(2 * t - a.py, a.h), ro = mirror_idx(2 * t + 1 - a.py, a.h);
            load_row(a, x_off, re, i0, fast_x, e);
            load_row(a, x_off, ro, i0, fast_x, o);
            T lo[NC][NS], hi[NC][NS];
            if (a.vskip) {
#pragma unroll
                for (int c = 0; c < NC; c++)
#pragma unroll
                    for (int s = 0; s < NS; s++) { lo[c][s] = e[c][s]; hi[c][s] = o[c][s]; }
            } else if constexpr (WT == 97) {
#pragma unroll
                for (int c = 0; c < NC; c++)
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        float d1 = lift97((float)po[c][s], (float)pe[c][s], (float)e[c][s], J2K_ALPHA);   // d1[t-1]
                        float s1 = lift97((float)pe[c][s], (float)d1p[c][s], d1, J2K_BETA);                // s1[t-1]
                        float d2 = lift97((float)d1p[c][s], (float)s1p[c][s], s1, J2K_GAMMA);              // d2[t-2]
                        float s2 = lift97((float)s1p[c][s], (float)d2p[c][s], d2, J2K_DELTA);              // s2[t-2]
                        lo[c][s] = (T)__fmul_rn(s2, J2K_INVK);
                        hi[c][s] = (T)__fmul_rn(d2, J2K_K);
                        pe[c][s] = e[c][s]; po[c][s] = o[c][s]; d1p[c][s] = (T)d1; s1p[c][s] = (T)s1; d2p[c][s] = (T)d2;
                    }
            } else {
#pragma unroll
                for (int c = 0; c < NC; c++)
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        int d = (int)po[c][s] - (((int)pe[c][s] + (int)e[c][s]) >> 1);  // d[t-1]
                        int sv = (int)pe[c][s] + (((int)d1p[c][s] + d + 2) >> 2);       // s[t-1]
                        lo[c][s] = (T)sv; hi[c][s] = (T)d;
                        pe[c][s] = e[c][s]; po[c][s] = o[c][s]; d1p[c][s] = (T)d;
                    }
            }
            const int ky = t - lag;  // finished vertical pair
            if (ky < ky0) continue;  // warm-up (warp-uniform)
            const int yl = ky - a.py, yh = ky;  // band rows of the low / high output row
#pragma unroll
            for (int c = 0; c < NC; c++) {
                hlift(lo[c], a.hskip);
                hlift(hi[c], a.hskip);
                if (!lane_out) continue;
                T lowv[NP], highv[NP];
                if (yl >= 0 && yl < a.lh) {
#pragma unroll
                    for (int j = 0; j < NP; j++) { lowv[j] = lo[c][2 * j]; highv[j] = lo[c][2 * j + 1]; }
                    store_band(a.ll, ll_off, c, yl, xl0, a.lw, lowv, a.vec_b);
                    store_band(a.hl, hl_off, c, yl, kx0, hw, highv, a.vec_b);
                }
                if (yh < hh) {
#pragma unroll
                    for (int j = 0; j < NP; j++) { lowv[j] = hi[c][2 * j]; highv[j] = hi[c][2 * j + 1]; }
                    store_band(a.lh_, lh_off, c, yh, xl0, a.lw, lowv, a.vec_b);
                    store_band(a.hh, hh_off, c, yh, kx0, hw, highv, a.vec_b);
                }
            }
        }
    }
};

template <int WT, int NP, int NC, int IN, int MCT>
__global__ void __launch_bounds__(128) fwd_level_kernel(const __grid_constant__ LevelArgs a) {
    FwdLevel<WT, NP, NC, IN, MCT>::run(a);
}

// ------------------------------------------------------------------ forward level kernel, fast path
//
// Same algorithm and arithmetic as FwdLevel, specialised for the geometry every BASELINE config has on
// its large levels: even column origin (px == 0), no 1-sample dimension, every row / plane / band
// offset a multiple of the vector width.  What changes is only HOW the work is issued:
//   * interior lanes move data with 128-bit (64-bit for u8) loads and stores through lane pointers
//     that are computed once; only lanes that straddle the left / right image border fall back to the
//     mirrored scalar path of FwdLevel;
//   * the two rows of iteration t+1 are fetched into registers before iteration t is computed
//     (software pipelining: two more independent 16-byte loads in flight per lane);
//   * the quantizer is branch-free: q = rint(c / step') with step' = step / scale (a power of two, so
//     (c / step) * scale == c / step' exactly) and the IEEE quotient from the correctly rounded
//     reciprocal (Markstein), 3 flops + 1 convert per coefficient (tests/test_div_markstein.py).
// Detail bands must be Q_QUANT / Q_RAW / Q_SHIFT; anything else stays on FwdLevel.

struct FastQ {  // per band, precomputed on the host
    float step, rcp;  // step' = step / scale, rcp = RN(1 / step')
    int mode, shift;
};

template <int WT, int NP, int NC, int IN, int MCT>
struct FwdFast {
    typedef FwdLevel<WT, NP, NC, IN, MCT> Slow;
    typedef typename Wt<WT>::T T;
    static constexpr int LAG = Wt<WT>::LAG;
    static constexpr int HLN = Slow::HLN, VP = Slow::VP, NS = Slow::NS;
    static constexpr int ES = (IN == IN_U8) ? 1 : (IN == IN_U16 ? 2 : 4);
    static constexpr int NW = NC * NS * ES / 4;  // 32-bit words per lane per row

    static __device__ __forceinline__ void fetch(const unsigned char* p, unsigned (&w)[NW]) {
        if constexpr (NW % 4 == 0) {
#pragma unroll
            for (int k = 0; k < NW / 4; k++) {
                uint4 q = __ldg((const uint4*)p + k);
                w[4 * k] = q.x; w[4 * k + 1] = q.y; w[4 * k + 2] = q.z; w[4 * k + 3] = q.w;
            }
        } else if constexpr (NW % 2 == 0) {
#pragma unroll
            for (int k = 0; k < NW / 2; k++) {
                uint2 q = __ldg((const uint2*)p + k);
                w[2 * k] = q.x; w[2 * k + 1] = q.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < NW; k++) w[k] = __ldg((const unsigned*)p + k);
        }
    }

    // words -> level-shifted working values (+ forward MCT); `sgn` is warp-uniform
    static __device__ __forceinline__ void unpack(const unsigned (&w)[NW], const RawFmt& r, T (&out)[NC][NS]) {
        int v[NC][NS];
#pragma unroll
        for (int s = 0; s < NS; s++)
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int e = NC * s + c;  // element index inside the lane's span
                if constexpr (IN == IN_U8) v[c][s] = (w[e >> 2] >> (8 * (e & 3))) & 0xFF;
                else if constexpr (IN == IN_U16) v[c][s] = (e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xFFFF);
                else v[c][s] = (int)w[e];
            }
        if constexpr (IN == IN_U8 || IN == IN_U16) {
            if (r.sign_sub) {
#pragma unroll
                for (int c = 0; c < NC; c++)
#pragma unroll
                    for (int s = 0; s < NS; s++) v[c][s] = raw_to_int(v[c][s], r);
            } else {
#pragma unroll
                for (int c = 0; c < NC; c++)
#pragma unroll
                    for (int s = 0; s < NS; s++) v[c][s] -= r.dc;
            }
        } else if constexpr (IN == IN_I32) {
#pragma unroll
            for (int c = 0; c < NC; c++)
#pragma unroll
                for (int s = 0; s < NS; s++) v[c][s] -= r.dc;
        }
#pragma unroll
        for (int s = 0; s < NS; s++) {
            if constexpr (IN == IN_F32) {
#pragma unroll
                for (int c = 0; c < NC; c++) out[c][s] = (T)__int_as_float(v[c][s]);
            } else if constexpr (NC == 3 && MCT != MCTK_NONE) {
                mct_forward<WT, MCT>(v[0][s], v[1][s], v[2][s], out[0][s], out[1][s], out[2][s]);
            } else {
#pragma unroll
                for (int c = 0; c < NC; c++) out[c][s] = (T)v[c][s];
            }
        }
    }

    static __device__ __forceinline__ void quant_vec(const T (&v)[NP], const FastQ& q, int (&o)[NP]) {
        if constexpr (WT == 53) {
#pragma unroll
            for (int j = 0; j < NP; j++) o[j] = (int)((unsigned)(int)v[j] << q.shift);  // shift == 0 for Q_RAW
        } else {
            if (q.mode == Q_QUANT) {
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    float c = (float)v[j];
                    float q0 = __fmul_rn(c, q.rcp);
                    float e = __fmaf_rn(-q0, q.step, c);
                    o[j] = __float2int_rn(__fmaf_rn(e, q.rcp, q0));
                }
            } else {
#pragma unroll
                for (int j = 0; j < NP; j++) o[j] = __float_as_int((float)v[j]);
            }
        }
    }

    static __device__ __forceinline__ void store_vec(int* p, const int (&o)[NP]) {
        if constexpr (NP == 4) *(int4*)p = make_int4(o[0], o[1], o[2], o[3]);
        else if constexpr (NP == 2) *(int2*)p = make_int2(o[0], o[1]);
        else *p = o[0];
    }

    static __device__ __forceinline__ void run(const LevelArgs& a, const FastQ (&fq)[4]) {
        const int lane = threadIdx.x & 31;
        const long long job = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const long long njobs = (long long)a.n_items * a.nchunks * a.nstrips;
        if (job >= njobs) return;
        const int strip = (int)(job % a.nstrips);
        const int chunk = (int)((job / a.nstrips) % a.nchunks);
        const int item = (int)(job / ((long long)a.nstrips * a.nchunks));

        const int kx0 = strip * VP - HLN * NP + lane * NP;
        const int i0 = 2 * kx0;  // px == 0
        const bool lane_out = lane >= HLN && lane < 32 - HLN;
        const int hw = a.w - a.lw, hh = a.h - a.lh;
        const bool ld_full = i0 >= 0 && i0 + NS <= a.w;
        const bool st_full = kx0 >= 0 && kx0 + NP <= hw;  // hw <= lw when px == 0
        const long long x_off = a.x_off[item];
        const unsigned char* xlane = (const unsigned char*)a.x_base + (x_off + (long long)i0 * (IN == IN_U8 || IN == IN_U16 ? NC : 1)) * ES;
        const long long xrow_bytes = (long long)a.x_row_stride * ES;
        const long long ll_off = a.ll.off[item], hl_off = a.hl.off[item], lh_off = a.lh_.off[item], hh_off = a.hh.off[item];
        int* p_ll = (int*)a.ll.base + ll_off + (long long)a.ll.y_off * a.ll.row_stride + a.ll.x_off + kx0;
        int* p_hl = (int*)a.hl.base + hl_off + (long long)a.hl.y_off * a.hl.row_stride + a.hl.x_off + kx0;
        int* p_lh = (int*)a.lh_.base + lh_off + (long long)a.lh_.y_off * a.lh_.row_stride + a.lh_.x_off + kx0;
        int* p_hh = (int*)a.hh.base + hh_off + (long long)a.hh.y_off * a.hh.row_stride + a.hh.x_off + kx0;

        const int ky0 = chunk * a.chunk_pairs;
        const int ky1 = min(ky0 + a.chunk_pairs, a.Ky);

        T pe[NC][NS], po[NC][NS], s1p[NC][NS], d1p[NC][NS], d2p[NC][NS];
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
            for (int s = 0; s < NS; s++) { pe[c][s] = 0; po[c][s] = 0; s1p[c][s] = 0; d1p[c][s] = 0; d2p[c][s] = 0; }

        const int t_begin = ky0 - LAG, t_end = ky1 + LAG;
        unsigned ne[NW], no[NW];
        if (ld_full) {
            fetch(xlane + (long long)mirror_idx(2 * t_begin - a.py, a.h) * xrow_bytes, ne);
            fetch(xlane + (long long)mirror_idx(2 * t_begin + 1 - a.py, a.h) * xrow_bytes, no);
        }
        for (int t = t_begin; t < t_end; t++) {
            T e[NC][NS], o[NC][NS];
            if (ld_full) {
                unsigned ce[NW], co[NW];
#pragma unroll
                for (int k = 0; k < NW; k++) { ce[k] = ne[k]; co[k] = no[k]; }
                if (t + 1 < t_end) {
                    fetch(xlane + (long long)mirror_idx(2 * t + 2 - a.py, a.h) * xrow_bytes, ne);
                    fetch(xlane + (long long)mirror_idx(2 * t + 3 - a.py, a.h) * xrow_bytes, no);
                }
                unpack(ce, a.raw, e);
                unpack(co, a.raw, o);
            } else {
                Slow::load_row(a, x_off, mirror_idx(2 * t - a.py, a.h), i0, false, e);
                Slow::load_row(a, x_off, mirror_idx(2 * t + 1 - a.py, a.h), i0, false, o);
            }
            T lo[NC][NS], hi[NC][NS];
            if constexpr (WT == 97) {
#pragma unroll
                for (int c = 0; c < NC; c++)
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        float d1 = lift97((float)po[c][s], (float)pe[c][s], (float)e[c][s], J2K_ALPHA);
                        float s1 = lift97((float)pe[c][s], (float)d1p[c][s], d1, J2K_BETA);
                        float d2 = lift97((float)d1p[c][s], (float)s1p[c][s], s1, J2K_GAMMA);
                        float s2 = lift97((float)s1p[c][s], (float)d2p[c][s], d2, J2K_DELTA);
                        lo[c][s] = (T)__fmul_rn(s2, J2K_INVK);
                        hi[c][s] = (T)__fmul_rn(d2, J2K_K);
                        pe[c][s] = e[c][s]; po[c][s] = o[c][s]; d1p[c][s] = (T)d1; s1p[c][s] = (T)s1; d2p[c][s] = (T)d2;
                    }
            } else {
#pragma unroll
                for (int c = 0; c < NC; c++)
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        int d = (int)po[c][s] - (((int)pe[c][s] + (int)e[c][s]) >> 1);
                        int sv = (int)pe[c][s] + (((int)d1p[c][s] + d + 2) >> 2);
                        lo[c][s] = (T)sv; hi[c][s] = (T)d;
                        pe[c][s] = e[c][s]; po[c][s] = o[c][s]; d1p[c][s] = (T)d;
                    }
            }
            const int ky = t - LAG;
            if (ky < ky0) continue;
            const int yl = ky - a.py, yh = ky;
            const bool row_l = yl >= 0 && yl < a.lh, row_h = yh < hh;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                Slow::hlift(lo[c], false);
                Slow::hlift(hi[c], false);
                if (!lane_out) continue;
                T lowv[NP], highv[NP];
                if (st_full) {
                    int q[NP];
                    if (row_l) {
#pragma unroll
                        for (int j = 0; j < NP; j++) { lowv[j] = lo[c][2 * j]; highv[j] = lo[c][2 * j + 1]; }
                        quant_vec(lowv, fq[0], q);
                        store_vec(p_ll + c * a.ll.comp_stride + (long long)yl * a.ll.row_stride, q);
                        quant_vec(highv, fq[1], q);
                        store_vec(p_hl + c * a.hl.comp_stride + (long long)yl * a.hl.row_stride, q);
                    }
                    if (row_h) {
#pragma unroll
                        for (int j = 0; j < NP; j++) { lowv[j] = hi[c][2 * j]; highv[j] = hi[c][2 * j + 1]; }
                        quant_vec(lowv, fq[2], q);
                        store_vec(p_lh + c * a.lh_.comp_stride + (long long)yh * a.lh_.row_stride, q);
                        quant_vec(highv, fq[3], q);
                        store_vec(p_hh + c * a.hh.comp_stride + (long long)yh * a.hh.row_stride, q);
                    }
                } else {
                    if (row_l) {
#pragma unroll
                        for (int j = 0; j < NP; j++) { lowv[j] = lo[c][2 * j]; highv[j] = lo[c][2 * j + 1]; }
                        Slow::store_band(a.ll, ll_off, c, yl, kx0, a.lw, lowv, false);
                        Slow::store_band(a.hl, hl_off, c, yl, kx0, hw, highv, false);
                    }
                    if (row_h) {
#pragma unroll
                        for (int j = 0; j < NP; j++) { lowv[j] = hi[c][2 * j]; highv[j] = hi[c][2 * j + 1]; }
                        Slow::store_band(a.lh_, lh_off, c, yh, kx0, a.lw, lowv, false);
                        Slow::store_band(a.hh, hh_off, c, yh, kx0, hw, highv, false);
                    }
                }
            }
        }
    }
};

struct FastQ4 { FastQ q[4]; };

template <int WT, int NP, int NC, int IN, int MCT>
__global__ void __launch_bounds__(128, (NP * NC >= 4 && WT == 97) ? 4 : 5)
fwd_fast_kernel(const __grid_constant__ LevelArgs a, const __grid_constant__ FastQ4 fq) {
    FwdFast<WT, NP, NC, IN, MCT>::run(a, fq.q);
}

// ------------------------------------------------------------------ inverse level kernel

// OUT: IN_I32 / IN_F32 = raw store of the working type (int32 for 5/3, float32 for 9/7);
//      IN_U8 / IN_U16  = final stage: round (9/7), inverse MCT, +DC, clamp, pack.
template <int WT, int NP, int NC, int OUT, int MCT>
struct InvLevel {
    typedef typename Wt<WT>::T T;
    static constexpr int LAG = Wt<WT>::LAG;
    static constexpr int HLN = (Wt<WT>::HALO + NP - 1) / NP;
    static constexpr int VP = (32 - 2 * HLN) * NP;
    static constexpr int NS = 2 * NP;
    static constexpr bool FINAL = (OUT == IN_U8 || OUT == IN_U16);

    // Loads NP values of band row `yrow` starting at band x index xi0 (fast) or, per pair j, at the
    // band index derived from the mirrored interleaved position.
    static __device__ __forceinline__ void load_band(const BandIO& b, long long item_off, int c, int yrow, int xi0, bool fast,
                                                     const int (&xi)[NP], T (&out)[NP]) {
        const int* p = (const int*)b.base + item_off + c * b.comp_stride + (long long)(b.y_off + yrow) * b.row_stride + b.x_off;
        int q[NP];
        if (fast) {
            if constexpr (NP == 4) { int4 v = __ldg((const int4*)(p + xi0)); q[0] = v.x; q[1] = v.y; q[2] = v.z; q[NP - 1] = v.w; }
            else if constexpr (NP == 2) { int2 v = __ldg((const int2*)(p + xi0)); q[0] = v.x; q[NP - 1] = v.y; }
            else q[0] = __ldg(p + xi0);
        } else {
#pragma unroll
            for (int j = 0; j < NP; j++) q[j] = __ldg(p + xi[j]);
        }
#pragma unroll
        for (int j = 0; j < NP; j++) out[j] = dequant_load_value<WT>(q[j], b);
    }

    // Horizontal synthesis across the warp: in  v[2j] = low_j, v[2j+1] = high_j; out v = interleaved samples.
    static __device__ __forceinline__ void hsynth(T (&v)[NS], bool hskip) {
        if (hskip) return;
        if constexpr (WT == 97) {
            float s[NP + 1], d[NP + 1];  // d[0] = previous lane's last high, s[NP] = next lane's first low
#pragma unroll
            for (int j = 0; j < NP; j++) {
                s[j] = __fmul_rn((float)v[2 * j], J2K_K);
                d[j + 1] = __fmul_rn((float)v[2 * j + 1], J2K_TWOINVK);
            }
            d[0] = __shfl_up_sync(0xffffffffu, d[NP], 1);
#pragma unroll
            for (int j = 0; j < NP; j++) s[j] = lift97(s[j], d[j], d[j + 1], -J2K_DELTA);
            s[NP] = __shfl_down_sync(0xffffffffu, s[0], 1);
#pragma unroll
            for (int j = 0; j < NP; j++) d[j + 1] = lift97(d[j + 1], s[j], s[j + 1], -J2K_GAMMA);
            d[0] = __shfl_up_sync(0xffffffffu, d[NP], 1);
#pragma unroll
            for (int j = 0; j < NP; j++) s[j] = lift97(s[j], d[j], d[j + 1], -J2K_BETA);
            s[NP] = __shfl_down_sync(0xffffffffu, s[0], 1);
#pragma unroll
            for (int j = 0; j < NP; j++) d[j + 1] = lift97(d[j + 1], s[j], s[j + 1], -J2K_ALPHA);
#pragma unroll
            for (int j = 0; j < NP; j++) { v[2 * j] = (T)s[j]; v[2 * j + 1] = (T)d[j + 1]; }
        } else {
            int s[NP + 1], d[NP + 1];
#pragma unroll
            for (int j = 0; j < NP; j++) { s[j] = (int)v[2 * j]; d[j + 1] = (int)v[2 * j + 1]; }
            d[0] = __shfl_up_sync(0xffffffffu, d[NP], 1);
#pragma unroll
            for (int j = 0; j < NP; j++) s[j] = s[j] - ((d[j] + d[j + 1] + 2) >> 2);
            s[NP] = __shfl_down_sync(0xffffffffu, s[0], 1);
#pragma unroll
            for (int j = 0; j < NP; j++) d[j + 1] = d[j + 1] + ((s[j] + s[j + 1]) >> 1);
#pragma unroll
            for (int j = 0; j < NP; j++) { v[2 * j] = (T)s[j]; v[2 * j + 1] = (T)d[j + 1]; }
        }
    }

    // Stores one finished interleaved row (all NC components) for this lane.
    static __device__ __forceinline__ void store_row(const LevelArgs& a, long long item_off, long long planes_off, int row, int i0,
                                                     bool fast, T (&val)[NC][NS]) {
        if constexpr (!FINAL) {
#pragma unroll
            for (int c = 0; c < NC; c++) {
                int* p = (int*)a.x_base + item_off + c * a.x_comp_stride + (long long)row * a.x_row_stride;
                int q[NS];
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    if constexpr (WT == 97) q[s] = a.x_mode == 1 ? __float2int_rn((float)val[c][s]) : __float_as_int((float)val[c][s]);
                    else q[s] = (int)val[c][s];
                }
                if (fast) {
                    if constexpr (NS == 4) *(int4*)(p + i0) = make_int4(q[0], q[1], q[2], q[NS - 1]);
                    else if constexpr (NS == 8) { *(int4*)(p + i0) = make_int4(q[0], q[1], q[2], q[3]); *(int4*)(p + i0 + 4) = make_int4(q[NS - 4], q[NS - 3], q[NS - 2], q[NS - 1]); }
                    else *(int2*)(p + i0) = make_int2(q[0], q[1]);
                } else {
#pragma unroll
                    for (int s = 0; s < NS; s++)
                        if (i0 + s >= 0 && i0 + s < a.w) p[i0 + s] = q[s];
                }
            }
        } else {
            store_row_final(a, item_off, planes_off, row, i0, fast, val);
        }
    }

    // final stage: round -> inverse MCT -> (+DC, optional planes) -> clamp -> pack
    static __device__ __forceinline__ void store_row_final(const LevelArgs& a, long long item_off, long long planes_off, int row, int i0,
                                                           bool fast, T (&val)[NC][NS]) {
        int iv[NC][NS];
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
            for (int s = 0; s < NS; s++) iv[c][s] = (WT == 97) ? __float2int_rn((float)val[c][s]) : (int)val[c][s];
        if constexpr (NC == 3 && MCT != MCTK_NONE) {
#pragma unroll
            for (int s = 0; s < NS; s++) {
                int r, g, b;
                mct_inverse<MCT>(iv[0][s], iv[1][s], iv[2][s], r, g, b);
                iv[0][s] = r; iv[1][s] = g; iv[2][s] = b;
            }
        }
        if (a.planes_out) {
#pragma unroll
            for (int c = 0; c < NC; c++) {
                int* p = a.planes_out + planes_off + c * a.planes_comp_stride + (long long)row * a.planes_row_stride;
#pragma unroll
                for (int s = 0; s < NS; s++)
                    if (i0 + s >= 0 && i0 + s < a.w) p[i0 + s] = iv[c][s] + a.raw.dc;
            }
        }
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
            for (int s = 0; s < NS; s++) iv[c][s] = int_to_raw(iv[c][s], a.raw);
        const int ps = a.raw.pix_stride;
        long long rbase = item_off + (long long)row * a.x_row_stride;
        if constexpr (NC == 1) if (fast) {
            if constexpr (OUT == IN_U16) {
                unsigned short* p = (unsigned short*)a.x_base + rbase + i0;
                unsigned w[NP];
#pragma unroll
                for (int j = 0; j < NP; j++) w[j] = (unsigned)iv[0][2 * j] | ((unsigned)iv[0][2 * j + 1] << 16);
                if constexpr (NP == 4) *(uint4*)p = make_uint4(w[0], w[1], w[2], w[NP - 1]);
                else if constexpr (NP == 2) *(uint2*)p = make_uint2(w[0], w[NP - 1]);
                else *(unsigned*)p = w[0];
            } else {
                unsigned char* p = (unsigned char*)a.x_base + rbase + i0;
                if constexpr (NS == 8) {
                    unsigned w0 = 0, w1 = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) { w0 |= (unsigned)iv[0][k] << (8 * k); w1 |= (unsigned)iv[0][(4 + k) % NS] << (8 * k); }
                    *(uint2*)p = make_uint2(w0, w1);
                } else if constexpr (NS == 4) {
                    unsigned w0 = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) w0 |= (unsigned)iv[0][k % NS] << (8 * k);
                    *(unsigned*)p = w0;
                } else {
                    *(unsigned short*)p = (unsigned short)(iv[0][0] | (iv[0][1] << 8));
                }
            }
            return;
        }
        if constexpr (NC == 3 && OUT == IN_U8 && NP >= 2) if (fast) {
            constexpr int NW = (3 * NS) / 4;
            unsigned wv[NW];
#pragma unroll
            for (int k = 0; k < NW; k++) wv[k] = 0;
#pragma unroll
            for (int s = 0; s < NS; s++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    int bi = 3 * s + c;
                    wv[bi >> 2] |= (unsigned)iv[c][s] << (8 * (bi & 3));
                }
            unsigned* p = (unsigned*)((unsigned char*)a.x_base + rbase + (long long)i0 * 3);
#pragma unroll
            for (int k = 0; k < NW; k++) p[k] = wv[k];
            return;
        }
        if constexpr (NC == 3 && OUT == IN_U16 && NP >= 2) if (fast) {
            constexpr int NW = (3 * NS) / 2;  // 32-bit words
            unsigned wv[NW];
#pragma unroll
            for (int k = 0; k < NW; k++) wv[k] = 0;
#pragma unroll
            for (int s = 0; s < NS; s++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    int wi = 3 * s + c;
                    wv[wi >> 1] |= (unsigned)iv[c][s] << (16 * (wi & 1));
                }
            uint2* p = (uint2*)((unsigned short*)a.x_base + rbase + (long long)i0 * 3);
#pragma unroll
            for (int k = 0; k < NW / 2; k++) p[k] = make_uint2(wv[2 * k], wv[2 * k + 1]);
            return;
        }
#pragma unroll
        for (int s = 0; s < NS; s++) {
            if (i0 + s < 0 || i0 + s >= a.w) continue;
#pragma unroll
            for (int c = 0; c < NC; c++) {
                long long e = rbase + (long long)(i0 + s) * ps + c;
                if constexpr (OUT == IN_U16) ((unsigned short*)a.x_base)[e] = (unsigned short)iv[c][s];
                else ((unsigned char*)a.x_base)[e] = (unsigned char)iv[c][s];
            }
        }
    }

    static __device__ __forceinline__ void run(const LevelArgs& a) {
        const int lane = threadIdx.x & 31;
        const long long job = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
        const long long njobs = (long long)a.n_items * a.nchunks * a.nstrips;
        if (job >= njobs) return;
        const int strip = (int)(job % a.nstrips);
        const int chunk = (int)((job / a.nstrips) % a.nchunks);
        const int item = (int)(job / ((long long)a.nstrips * a.nchunks));

        const int kx0 = strip * VP - HLN * NP + lane * NP;
        const int i0 = 2 * kx0 - a.px;
        const bool lane_out = lane >= HLN && lane < 32 - HLN;
        const int hw = a.w - a.lw;
        const int xl0 = kx0 - a.px;
        // band-side fast path: all NP low and high indices in range and contiguous
        const bool fast_b = a.vec_b && xl0 >= 0 && xl0 + NP <= a.lw && kx0 >= 0 && kx0 + NP <= hw;
        const bool fast_x = a.vec_x && i0 >= 0 && i0 + NS <= a.w;
        int xil[NP], xih[NP];  // mirrored band indices for the slow path
#pragma unroll
        for (int j = 0; j < NP; j++) {
            int il = mirror_idx(2 * (kx0 + j) - a.px, a.w);      // low-type sample
            int ih = mirror_idx(2 * (kx0 + j) + 1 - a.px, a.w);  // high-type sample
            xil[j] = a.hskip ? 0 : (il - a.px) >> 1;
            xih[j] = a.hskip ? 0 : (ih - (1 - a.px)) >> 1;
        }
        // hskip (w == 1): the only column is low when px == 0, high when px == 1
        const bool has_low = a.lw > 0, has_high = hw > 0;

        const long long x_off = a.x_off[item];
        const long long planes_off = (FINAL && a.planes_out) ? a.planes_off[item] : 0;
        const long long ll_off = a.ll.off[item], hl_off = a.hl.off[item], lh_off = a.lh_.off[item], hh_off = a.hh.off[item];
        const int ky0 = chunk * a.chunk_pairs;
        const int ky1 = min(ky0 + a.chunk_pairs, a.Ky);
        const int hh = a.h - a.lh;

        T dp[NC][NS], s1p[NC][NS], d1p[NC][NS], s2p[NC][NS];
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
            for (int s = 0; s < NS; s++) { dp[c][s] = 0; s1p[c][s] = 0; d1p[c][s] = 0; s2p[c][s] = 0; }

        const int lag = a.vskip ? 0 : LAG;
        for (int t = ky0 - lag; t < ky1 + lag; t++) {
            // vertical positions of this pair in the interleaved column, mirrored, and their band rows
            int yl, yh;
            if (a.vskip) { yl = 0; yh = 0; }
            else {
                int pl = mirror_idx(2 * t - a.py, a.h), ph = mirror_idx(2 * t + 1 - a.py, a.h);
                yl = (pl - a.py) >> 1;
                yh = (ph - (1 - a.py)) >> 1;
            }
            T e[NC][NS], o[NC][NS];  // horizontally synthesized low-type / high-type rows
#pragma unroll
            for (int c = 0; c < NC; c++) {
                T lowv[NP], highv[NP];
#pragma unroll
                for (int j = 0; j < NP; j++) { lowv[j] = 0; highv[j] = 0; }
                if (a.lh > 0) {
                    if (has_low) load_band(a.ll, ll_off, c, yl, xl0, fast_b, xil, lowv);
                    if (has_high) load_band(a.hl, hl_off, c, yl, kx0, fast_b, xih, highv);
                }
#pragma unroll
                for (int j = 0; j < NP; j++) { e[c][2 * j] = lowv[j]; e[c][2 * j + 1] = highv[j]; }
                hsynth(e[c], a.hskip);
#pragma unroll
                for (int j = 0; j < NP; j++) { lowv[j] = 0; highv[j] = 0; }
                if (hh > 0) {
                    if (has_low) load_band(a.lh_, lh_off, c, yh, xl0, fast_b, xil, lowv);
                    if (has_high) load_band(a.hh, hh_off, c, yh, kx0, fast_b, xih, highv);
                }
#pragma unroll
                for (int j = 0; j < NP; j++) { o[c][2 * j] = lowv[j]; o[c][2 * j + 1] = highv[j]; }
                hsynth(o[c], a.hskip);
            }
            T xe[NC][NS], xo[NC][NS];  // finished rows of pair t - lag
            if (a.vskip) {
#pragma unroll
                for (int c = 0; c < NC; c++)
#pragma unroll
                    for (int s = 0; s < NS; s++) { xe[c][s] = e[c][s]; xo[c][s] = o[c][s]; }
            } else if constexpr (WT == 97) {
#pragma unroll
                for (int c = 0; c < NC; c++)
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        float sv = __fmul_rn((float)e[c][s], J2K_K), dv = __fmul_rn((float)o[c][s], J2K_TWOINVK);
                        float s1 = lift97(sv, (float)dp[c][s], dv, -J2K_DELTA);                      // s'[t]
                        float d1 = lift97((float)dp[c][s], (float)s1p[c][s], s1, -J2K_GAMMA);        // d'[t-1]
                        float s2 = lift97((float)s1p[c][s], (float)d1p[c][s], d1, -J2K_BETA);        // s''[t-1]
                        float d2 = lift97((float)d1p[c][s], (float)s2p[c][s], s2, -J2K_ALPHA);       // d''[t-2]
                        xe[c][s] = s2p[c][s];                                                          // s''[t-2]
                        xo[c][s] = (T)d2;
                        dp[c][s] = (T)dv; s1p[c][s] = (T)s1; d1p[c][s] = (T)d1; s2p[c][s] = (T)s2;
                    }
            } else {
#pragma unroll
                for (int c = 0; c < NC; c++)
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        int sv = (int)e[c][s] - (((int)dp[c][s] + (int)o[c][s] + 2) >> 2);  // s[t]
                        int xodd = (int)dp[c][s] + (((int)s1p[c][s] + sv) >> 1);            // x[2(t-1)+1]
                        xe[c][s] = s1p[c][s];
                        xo[c][s] = (T)xodd;
                        dp[c][s] = o[c][s]; s1p[c][s] = (T)sv;
                    }
            }
            const int ky = t - lag;
            if (ky < ky0 || !lane_out) continue;
            const int re = 2 * ky - a.py, ro = 2 * ky + 1 - a.py;
            if (re >= 0 && re < a.h) store_row(a, x_off, planes_off, re, i0, fast_x, xe);
            if (ro >= 0 && ro < a.h) store_row(a, x_off, planes_off, ro, i0, fast_x, xo);
        }
    }
};

template <int WT, int NP, int NC, int OUT, int MCT>
__global__ void __launch_bounds__(128) inv_level_kernel(const __grid_constant__ LevelArgs a) {
    InvLevel<WT, NP, NC, OUT, MCT>::run(a);
}

}  // namespace j2k
