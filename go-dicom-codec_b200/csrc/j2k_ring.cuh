// j2k_ring.cuh — persistent, TMA-staged DWT kernels for sm_100a (the production path on regular geometries).
//
// ONE launch runs every decomposition level of every frame / tile of a batch:
//
//   * the work is a flat, ordered list of jobs (level-major, then item, row chunk, column strip); the
//     warps of a grid sized to the machine (148 SMs x resident CTAs) claim jobs with an atomic counter,
//     so there is no wave quantisation and no per-level launch gap;
//   * a job of level k+1 waits (acquire on a per-item counter) until the jobs of level k that produce
//     its LL input have published their stores (release); jobs are claimed in order, so a waiting warp
//     only ever waits on warps that are already running — no co-residency requirement, no grid sync,
//     and the small deep levels overlap the tail of the big ones;
//   * inside a job a warp owns a strip of the level window: 32 lanes x NP sample pairs wide, `chunk_pairs`
//     row pairs tall (+ halo).  Rows are staged HBM -> shared memory by the TMA engine
//     (cp.async.bulk + mbarrier complete_tx; SASS: UBLKCP / SYNCS) into a warp-private ring of D
//     row-pair stages, D-1 stages always in flight, so the bytes in flight per SM are set by the ring
//     (~7 KB per warp) and not by registers.  Whole-sample symmetric extension at the image borders is
//     an index computation on the staged row (or on the row number for the vertical direction): every
//     global read is a plain, aligned, in-bounds segment;
//   * the arithmetic is the register sliding-window vertical lifting + warp-shuffle horizontal lifting
//     of j2k_kernels.cuh (same operation order, no FMA contraction), fused with unpack / DC shift / MCT
//     in front and the quantizer behind; results are bit-identical to the per-level kernels.
//
// The emulator build (tests/emu, J2K_EMU) runs the same code with the async copy executed synchronously.
#pragma once
#include "j2k_kernels.cuh"

namespace j2k {

#define J2K_RING_MAXSEG 14   // (tile class, level) segments one launch can chain
#ifndef J2K_RING_WARPS
#define J2K_RING_WARPS 4     // warps per CTA
#endif
#ifndef J2K_RING_BYTES
#define J2K_RING_BYTES 12800  // staging bytes per warp: 5 stages of four 544 B rows (u16), 3 stages of four 1056 B rows (float32)
#endif
#define J2K_RING_MAXD 8
#ifndef J2K_RING_L2_HINTS
#define J2K_RING_L2_HINTS 0  // 1: L2 eviction-priority policies on the staged loads and the band stores of the forward kernel
#endif
#if J2K_RING_L2_HINTS
#define J2K_POL_ARG(p) , p
#else
#define J2K_POL_ARG(p)
#endif
#ifndef J2K_RGB_SINGLE_BODY
#define J2K_RGB_SINGLE_BODY 1
#endif
#ifndef J2K_RING_MINB
#define J2K_RING_MINB 3      // resident CTAs per SM the register allocation targets (168 registers: no spills in the 9/7 loop)
#endif
#ifndef J2K_RING_BYTES_UA
#define J2K_RING_BYTES_UA 13056  // general-alignment variant: 3 stages of four 1088 B rows (float32), 5 of four 576 B rows (u16)
#endif
#ifndef J2K_INV_RING_BYTES
#define J2K_INV_RING_BYTES 18432  // inverse: a stage holds four band rows per component (4 stages of 2 x four 576 B rows for halo-free strips)
#endif
#ifndef J2K_INV_RING_BYTES_UA
#define J2K_INV_RING_BYTES_UA 18432  // general-alignment variant: 4 stages of 2 x four 576 B rows (9/7), 3 stages of 2 x four 608 B rows (5/3)
#endif
#ifndef J2K_INV_HALO_FREE
#define J2K_INV_HALO_FREE 1       // 5/3 inverse: strips of 32 storing lanes (InvRing::HF)
#endif
#ifndef J2K_INV_RING_BYTES_53
#define J2K_INV_RING_BYTES_53 13824  // single-component 5/3 inverse: 3 stages, so that 4 CTAs (16 warps) fit one SM
#endif

struct RingSeg {
    // level window (px == 0 always on this path)
    int w, h, py, lw, lh, Kx, Ky;
    int n_items, nchunks, nstrips, chunk_pairs, strip_pairs;
    int first;                         // 1 = image-side variant (raw words / planar api input), 0 = planar working type
    int dep_seg, dep_div, dep_target;  // wait until ctl[done_base(dep_seg) + item / dep_div] == dep_target (dep_seg < 0: none)
    int job_begin, job_end;            // this segment's slice of the job list
    int done_base;                     // index in ctl[] of this segment's per-item completion counters
    int row_bytes;                     // w * bytes per pixel position (multiple of 16)
    int dc;                            // DC level shift applied on load (image-side variant)
    int has_waiters;                   // 1 = some segment's jobs wait on this segment's completion counters
    const unsigned char* x_base;       // interleaved side (forward: source, inverse: destination)
    const long long* x_off;            // per item, elements
    long long x_row_bytes;             // row pitch in bytes (multiple of 16)
    BandIO ll, hl, lh_, hh;
    FastQ q[4];
    float2 rcpE, nstE, rcpO, nstO;     // forward: quantizer pairs (LL, LH) and (HL, HH): reciprocal and negated step' (9/7)
                                       // inverse: rcpE / rcpO hold the dequantizer scale pairs (LL, LH) / (HL, HH)
    // general-alignment variant (UA): bytes every staged lane address is aligned to; store vector width (ints) of the
    // LL / HL / LH / HH rows (forward) or bytes the pixel / plane rows are aligned to (inverse: x_align)
    int load_align, x_align;
    int st_cls[4];
    int dep_mul;                       // inverse: the job waits on dep_mul consecutive counters starting at item * dep_mul
    int x_mode;                        // inverse, planar destination: 1 = store the 9/7 samples rounded half-even to int32
    int32_t* planes_out;               // inverse final: optional GetImageData planes (decoder.go:738-740)
    const long long* planes_off;
    long long planes_comp_stride;
    int planes_row_stride;
    int pad2_;
    // L2 eviction-priority policies (createpolicy encodings): staged loads, LL stores, detail-band / pixel stores
    unsigned long long pol_load, pol_ll, pol_band;
};
#define J2K_L2_EVICT_NORMAL 0x1000000000000000ull
#define J2K_L2_EVICT_FIRST 0x12F0000000000000ull
#define J2K_L2_EVICT_LAST 0x14F0000000000000ull

struct RingArgs {
    int nseg, total_jobs;
    unsigned* ctl;  // [0] job counter, [1] retired-warp counter, [2..] per-(segment,item) completion counters
    int n_ctl;      // entries of ctl (for the self-reset at kernel end)
    float one;      // 1.0f, opaque to the compiler (see addp2 in j2k_ring.cuh)
    // Group-pipelined job order (0 slices: the plain level-major list).  Slice s covers jobs [slice_begin[s],
    // slice_begin[s+1]) = the jobs of segment slice_info[s].x for the items starting at slice_info[s].y.
    int nslice;
    const int* slice_begin;
    const int2* slice_info;
    RawFmt raw;
    RingSeg seg[J2K_RING_MAXSEG];
};
#define J2K_RING_MAXSLICE 256
#ifndef J2K_RING_DEFAULT_LAG
#define J2K_RING_DEFAULT_LAG 0          // groups of level-1 work between a slice and its consumer (0: level-major list)
#endif
#ifndef J2K_RING_DEFAULT_GROUP_KS
#define J2K_RING_DEFAULT_GROUP_KS 16384  // group size in Ki samples of the largest level (one 4096 x 4096 frame)
#endif

// ------------------------------------------------------------------ async-copy / barrier primitives

#ifdef J2K_EMU
#define J2K_SMEM_DECL(name) unsigned char* name = emu::smem()
typedef unsigned char* smem_t;  // "shared address": a host pointer in the emulator
__device__ __forceinline__ smem_t smem_handle(void* p) { return (unsigned char*)p; }
__device__ __forceinline__ unsigned char* smem_ptr(smem_t h) { return h; }
__device__ __forceinline__ void mbar_init(smem_t, int) {}
__device__ __forceinline__ void mbar_fence_init() {}
__device__ __forceinline__ void mbar_expect_tx(smem_t, unsigned) {}
__device__ __forceinline__ void mbar_wait(smem_t, unsigned) {}
__device__ __forceinline__ void bulk_g2s(smem_t dst, const void* src, unsigned bytes, smem_t, unsigned long long = 0) { memcpy(dst, src, bytes); }
__device__ __forceinline__ void stg128_hint(int* p, int a, int b, int c, int d, unsigned long long) { p[0] = a; p[1] = b; p[2] = c; p[3] = d; }
__device__ __forceinline__ void fence_proxy_async() {}
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) { return *p; }
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) { return *p; }
__device__ __forceinline__ void fence_proxy_async_global() {}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) { *p += v; }
__device__ __forceinline__ void backoff() {}
__device__ __forceinline__ bool elect_one() { return emu::lane_id() == 0; }
__device__ __forceinline__ uint4 lds128(smem_t a) { uint4 v; memcpy(&v, a, 16); return v; }
__device__ __forceinline__ uint2 lds64(smem_t a) { uint2 v; memcpy(&v, a, 8); return v; }
__device__ __forceinline__ unsigned lds32(smem_t a) { unsigned v; memcpy(&v, a, 4); return v; }
__device__ __forceinline__ unsigned lds16(smem_t a) { unsigned short v; memcpy(&v, a, 2); return v; }
__device__ __forceinline__ unsigned lds8(smem_t a) { return *a; }
__device__ __forceinline__ void sts32(smem_t a, unsigned v) { memcpy(a, &v, 4); }
__device__ __forceinline__ void sts128f(smem_t a, float2 u, float2 v) { float t[4] = {u.x, u.y, v.x, v.y}; memcpy(a, t, 16); }
__device__ __forceinline__ void mbar_arrive(smem_t) {}
#else
#define J2K_SMEM_DECL(name) extern __shared__ __align__(128) unsigned char name[]
typedef unsigned smem_t;  // 32-bit shared-window address
__device__ __forceinline__ smem_t smem_handle(void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned char* smem_ptr(smem_t h) { return (unsigned char*)__cvta_shared_to_generic((size_t)h); }
__device__ __forceinline__ void mbar_init(smem_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(smem_t bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(smem_t bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
}
// TMA bulk copy global -> shared, completion signalled on `bar` (complete_tx::bytes).  SASS: UBLKCP.
__device__ __forceinline__ void bulk_g2s(smem_t dst, const void* src, unsigned bytes, smem_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// the same with an L2 eviction-priority policy
__device__ __forceinline__ void bulk_g2s(smem_t dst, const void* src, unsigned bytes, smem_t bar, unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void stg128_hint(int* p, int a, int b, int c, int d, unsigned long long pol) {
    asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d), "l"(pol) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_relaxed(const unsigned* p) {
    unsigned v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// generic-proxy accesses before, async-proxy (bulk copy) accesses of global memory after.  SASS: FENCE.VIEW.ASYNC.G alone;
// the state-space-less form adds a MEMBAR.ALL.GPU.
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void backoff() { __nanosleep(400); }
// One lane of the (converged) warp: the form ptxas recognises as warp-uniform single-thread code, so the bulk copies
// it guards take their operands from uniform registers without a per-copy broadcast loop.
__device__ __forceinline__ bool elect_one() {
    unsigned p;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(p));
    return p != 0;
}
__device__ __forceinline__ uint4 lds128(smem_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint2 lds64(smem_t a) {
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds32(smem_t a) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds16(smem_t a) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ unsigned lds8(smem_t a) {
    unsigned v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts32(smem_t a, unsigned v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts128f(smem_t a, float2 u, float2 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(u.x), "f"(u.y), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void mbar_arrive(smem_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
#endif

struct RingWarp {
    smem_t ring;     // J2K_RING_BYTES of staging, 16-byte aligned
    smem_t bars;     // J2K_RING_MAXD mbarriers (8 bytes each)
    unsigned phase;  // bit s = parity the next wait on barrier s uses
    float one;       // 1.0f the compiler cannot see (see lift97x2)
};

// single reflection is enough next to the window; anything else (tiny windows) takes the general form
__device__ __noinline__ int mirror_slow(int i, int n) { return mirror_idx(i, n); }
__device__ __forceinline__ int mirror_fast(int i, int n) {
    if ((unsigned)i < (unsigned)n) return i;
    int r = i < 0 ? -i : 2 * (n - 1) - i;
    if ((unsigned)r < (unsigned)n) return r;
    return mirror_slow(i, n);
}

// ------------------------------------------------------------------ packed fp32x2 arithmetic
//
// sm_100a has two-wide fp32 instructions (PTX add/mul/fma.rn.f32x2, SASS FADD2 / FMUL2 / FFMA2): each half is
// rounded exactly like the scalar instruction, so the lifting stays bit-identical to the reference while the
// issue slots spent on arithmetic are halved.  Vertical lifting pairs two adjacent columns; horizontal lifting
// pairs the two rows (vertical low / high) that one iteration finishes.

#ifdef J2K_EMU
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y)); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return make_float2(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return make_float2(__fmaf_rn(a.x, b.x, c.x), __fmaf_rn(a.y, b.y, c.y)); }
#else
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#endif
__device__ __forceinline__ float2 splat2(float c) { return make_float2(c, c); }
// a + b where a is a product: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (single rounding) even
// under -fmad=false, which would break bit-exactness.  fma(a, one, b) with a 1.0f it cannot see (a kernel
// parameter) rounds once, a * 1 + b = a + b, and cannot absorb the multiply that produced a.
__device__ __forceinline__ float2 addp2(float2 a, float2 b, float one) { return fma2(a, make_float2(one, one), b); }
// x + (l + r) * c, three roundings per half (dwt97.go:104-116)
__device__ __forceinline__ float2 lift97x2(float2 x, float2 l, float2 r, float2 c, float one) {
    return addp2(mul2(add2(l, r), c), x, one);
}
__device__ __forceinline__ float2 shfl_down2(float2 v) {
    return make_float2(__shfl_down_sync(0xffffffffu, v.x, 1), __shfl_down_sync(0xffffffffu, v.y, 1));
}
__device__ __forceinline__ float2 shfl_up2(float2 v) {
    return make_float2(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1));
}

// ------------------------------------------------------------------ packed 16-bit clamp / pack (inverse, final level)
//
// +DC, clamp, two's-complement wrap and packing of GetPixelData (decoder.go:777-944) on two samples per instruction:
// saturate both to int16 (the clamp range lies inside it), clamp to [lo - dc, hi - dc], add dc modulo 2^16, keep the
// low B bits (the un-sign-extended word of a negative signed sample).  SASS: I2IP.S16.S32.SAT, VIMNMX.S16x2, VIADD.16x2.
struct RawPack { unsigned lo2, hi2, dc2, mask2, lo2d, hi2d; float magic_dc; };
__device__ __forceinline__ RawPack make_raw_pack(const RawFmt& r) {
    RawPack p;
    p.lo2 = ((unsigned)(r.clamp_lo - r.dc) & 0xFFFFu) * 0x10001u;
    p.hi2 = ((unsigned)(r.clamp_hi - r.dc) & 0xFFFFu) * 0x10001u;
    p.dc2 = ((unsigned)r.dc & 0xFFFFu) * 0x10001u;
    p.mask2 = ((unsigned)(r.clamp_hi - r.clamp_lo) & 0xFFFFu) * 0x10001u;  // 2^B - 1 in both halves
    // the same clamp on values that carry the DC shift already (pack_clamp2_magic_dc: the shift rides in the rounding constant)
    p.lo2d = ((unsigned)r.clamp_lo & 0xFFFFu) * 0x10001u;
    p.hi2d = ((unsigned)r.clamp_hi & 0xFFFFu) * 0x10001u;
    p.magic_dc = 12582912.0f + (float)r.dc;   // 1.5 * 2^23 + dc, exact
    return p;
}
#ifdef J2K_EMU
__device__ __forceinline__ unsigned pack_clamp2(int a, int b, const RawPack& p) {
    auto one = [&](int v, int sh) {
        v = v < -32768 ? -32768 : (v > 32767 ? 32767 : v);
        const int lo = (short)(p.lo2 >> sh), hi = (short)(p.hi2 >> sh);
        v = v < lo ? lo : (v > hi ? hi : v);
        return (((unsigned)v + (p.dc2 >> sh)) & (p.mask2 >> sh)) & 0xFFFFu;
    };
    return one(a, 0) | (one(b, 16) << 16);
}
#else
__device__ __forceinline__ unsigned pack_clamp2(int a, int b, const RawPack& p) {  // a -> low half, b -> high half
    unsigned w;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(w) : "r"(b), "r"(a));
    w = __vmaxs2(w, p.lo2);
    w = __vmins2(w, p.hi2);
    w = __vadd2(w, p.dc2);
    return w & p.mask2;
}
#endif

// The same on two values that arrive as the bit patterns of (n + 1.5 * 2^23) in float32, |n| < 2^15: the low 16 bits of
// each pattern are n modulo 2^16, so one PRMT replaces the two F2I and the saturating pack.
__device__ __forceinline__ unsigned pack_clamp2_magic(unsigned ua, unsigned ub, const RawPack& p) {  // ua -> low half, ub -> high half
#ifdef J2K_EMU
    return pack_clamp2((int)(short)(ua & 0xFFFFu), (int)(short)(ub & 0xFFFFu), p);
#else
    unsigned w = __byte_perm(ua, ub, 0x5410);
    w = __vmaxs2(w, p.lo2);
    w = __vmins2(w, p.hi2);
    w = __vadd2(w, p.dc2);
    return w & p.mask2;
#endif
}

// ... and with the DC shift folded into the bias (patterns of n + dc + 1.5 * 2^23): clamp to [clamp_lo, clamp_hi] directly,
// no packed add; the mask only matters for signed samples (wrap of the negative words), where dc is 0.
__device__ __forceinline__ unsigned pack_clamp2_magic_dc(unsigned ua, unsigned ub, const RawPack& p) {
#ifdef J2K_EMU
    auto one = [&](unsigned u, int sh) {
        int v = (int)(short)(u & 0xFFFFu);
        const int lo = (short)(p.lo2d >> sh), hi = (short)(p.hi2d >> sh);
        v = v < lo ? lo : (v > hi ? hi : v);
        return ((unsigned)v & (p.mask2 >> sh)) & 0xFFFFu;
    };
    return one(ua, 0) | (one(ub, 16) << 16);
#else
    unsigned w = __byte_perm(ua, ub, 0x5410);
    w = __vmaxs2(w, p.lo2d);
    w = __vmins2(w, p.hi2d);
    return w & p.mask2;
#endif
}

#ifndef J2K_ICT_FAST
#define J2K_ICT_FAST 1       // float32 form of the float64 inverse ICT behind a distance-to-tie guard (InvRing::ict_fast_row)
#endif

// ------------------------------------------------------------------ forward job

// XC = 3: component-split first level of a 3-component frame.  The staged rows hold interleaved RGB pixels, but the job
// produces ONE component (item = 3 * pixel item + component): it converts all three samples of a pixel and evaluates
// only its own row of the ICT, so the warp runs the single-component pipeline (NP = 4, full occupancy) at the price of
// converting the raw bytes three times (they arrive through L2 / shared memory, not HBM).
// UA = 1: the general-alignment variant.  Nothing about the window has to be a multiple of anything: rows may start at any
// byte (the TMA copy then starts at the 16-byte boundary below the wanted segment and the row keeps its own phase, stored
// in the slot's tail), lanes read their spans with the widest load every row's phase allows, band rows are stored with the
// widest vector their alignment allows and masked at the window's right edge, so widths like 2140 or 2022 (CR / DX
// detectors) and odd LL windows stay on the persistent kernel instead of falling to the per-level kernels.
template <int UA> __host__ __device__ constexpr int fwd_ring_bytes() { return UA ? J2K_RING_BYTES_UA : J2K_RING_BYTES; }
template <int UA> __host__ __device__ constexpr int fwd_cta_smem() { return J2K_RING_WARPS * (fwd_ring_bytes<UA>() + J2K_RING_MAXD * 8); }

// Exchange ring of the one-producer level-1 forward (fwd3w_kernel): DX stages, a stage = the four rows (two row pairs) of
// one raw stage as float32 ICT outputs of all three components, [component][row][half][lane] x 16 bytes (conflict-free
// 128-bit accesses; a lane's eight samples are half 0 = samples 0..3, half 1 = samples 4..7).  The producer warp fills a
// stage and arrives on full[stage]; the three consumer warps (one per component) each arrive on empty[stage] when they
// are done reading it.
#ifndef J2K_F3_DX
#define J2K_F3_DX 3
#endif
#define J2K_F3_ROWB 1024
#define J2K_F3_COMPB (4 * J2K_F3_ROWB)
#define J2K_F3_STAGEB (3 * J2K_F3_COMPB)
struct FxPort {
    smem_t data, full, empty;
    unsigned cnt;   // stages produced / consumed so far by this warp (all four warps of the CTA count the same stages)
    int dx;         // stages in the ring
};

// XCH = 1: the job is one COMPONENT of a three-component level 1 whose rows arrive converted (float32 ICT outputs) through
// the CTA's exchange ring instead of the TMA ring; XCH = 2: the producer of that ring (stages the raw pixel rows once,
// converts all three components, no wavelet).
template <int WT, int NP, int NC, int IN, int MCT, int SG, int XC = 1, int UA = 0, int XCH = 0>
struct FwdRing {
    typedef FwdLevel<WT, NP, NC, IN, MCT> Slow;
    typedef typename Wt<WT>::T T;
    static constexpr int LAG = Wt<WT>::LAG;
    static constexpr int HLN = Slow::HLN, NS = Slow::NS;
    static constexpr bool RAWIN = (IN == IN_U8 || IN == IN_U16);
    static constexpr int ES = (IN == IN_U8) ? 1 : (IN == IN_U16 ? 2 : 4);
    static constexpr int PB = ES * (RAWIN ? NC * XC : 1);  // bytes per pixel position of one staged row
    static constexpr int LB = NS * PB;                // bytes per lane per row
    static constexpr int NW = LB / 4;                 // 32-bit words per lane per row
    static constexpr int ROWB = 32 * LB + 32 + (UA ? 32 : 0);  // staged row slot (16 B slack for the alignment phase, 16 B rounding; UA: + the row's own phase, and its value in the last word)
    static constexpr int RPS = 2;                     // row pairs (= loop iterations) per stage: one barrier, one issue per two
    static_assert(!UA || (NC == 1 && XC == 1 && NP == 4), "general alignment: single-component jobs");
    static constexpr int STAGEB = 2 * RPS * ROWB;
    static constexpr int RBYTES = fwd_ring_bytes<UA>();
    static constexpr int D = (RBYTES / STAGEB) > J2K_RING_MAXD ? J2K_RING_MAXD : (RBYTES / STAGEB);
    static constexpr int LDALIGN = (LB % 16 == 0) ? 16 : (LB % 8 == 0 ? 8 : 4);
    static constexpr bool MAGIC = (WT == 97 && RAWIN && SG == 0);  // u8/u16 -> float32 without I2F
    static_assert(D >= 2, "ring too small for this row size");
    static_assert(LB % 4 == 0, "lane span must be whole words");
    static_assert(NC == 1 || RAWIN, "3-component jobs read interleaved raw words");
    static_assert(XC == 1 || (XC == 3 && NC == 1 && MAGIC && MCT == MCTK_ICT), "component split: unsigned raw RGB, float32 ICT");
    static_assert(XCH == 0 || (WT == 97 && NP == 4 && NC == 1 && !UA), "exchange mode: single-component 9/7 jobs");
    static_assert(XCH != 1 || (IN == IN_F32 && XC == 1), "exchange consumer reads float32 rows");
    static_assert(XCH != 2 || XC == 3, "exchange producer stages interleaved raw RGB");
    static constexpr int SROWB = XCH == 1 ? J2K_F3_ROWB : ROWB;   // distance of the rows the job loop reads

    static __device__ __forceinline__ void fetch(smem_t p, unsigned (&w)[NW]) {
        if constexpr (LDALIGN == 16) {
#pragma unroll
            for (int k = 0; k < NW / 4; k++) {
                uint4 q = lds128(p + 16 * k);
                w[4 * k] = q.x; w[4 * k + 1] = q.y; w[4 * k + 2] = q.z; w[4 * k + 3] = q.w;
            }
        } else if constexpr (LDALIGN == 8) {
#pragma unroll
            for (int k = 0; k < NW / 2; k++) {
                uint2 q = lds64(p + 8 * k);
                w[2 * k] = q.x; w[2 * k + 1] = q.y;
            }
        } else {
#pragma unroll
            for (int k = 0; k < NW; k++) w[k] = lds32(p + 4 * k);
        }
    }

    // the same from a lane address at ANY byte phase (UA): lane_off + phase = an aligned part (multiple of the lane vector
    // LDALIGN) plus a warp-uniform byte offset p < LDALIGN.  The lane loads its aligned vectors plus one more with the
    // conflict-free 128- / 64-bit loads of the aligned kernels and extracts its words with funnel shifts (narrow loads at a
    // 16-byte lane stride would be 4-way bank-conflicted: measured 2.6 x the short-scoreboard stalls).
    static __device__ __forceinline__ void fetch_ua(smem_t row, int off, unsigned (&w)[NW]) {
        constexpr int AW = LDALIGN / 4;          // words per aligned vector
        constexpr int NV = LB / LDALIGN + 1;     // aligned vectors that cover the lane's span at any phase
        const int p = off & (LDALIGN - 1);
        const smem_t a = row + (off - p);
        unsigned W[NV * AW];
#pragma unroll
        for (int v = 0; v < NV; v++) {
            if constexpr (LDALIGN == 16) { const uint4 q = lds128(a + 16 * v); W[4 * v] = q.x; W[4 * v + 1] = q.y; W[4 * v + 2] = q.z; W[4 * v + 3] = q.w; }
            else if constexpr (LDALIGN == 8) { const uint2 q = lds64(a + 8 * v); W[2 * v] = q.x; W[2 * v + 1] = q.y; }
            else W[v] = lds32(a + 4 * v);
        }
        const int q = p >> 2, sh = 8 * (p & 3);  // warp-uniform: word offset, bit shift
#pragma unroll
        for (int qq = 0; qq < AW; qq++)
            if (q == qq) {
#pragma unroll
                for (int k = 0; k < NW; k++) w[k] = __funnelshift_r(W[qq + k], W[qq + k + 1], sh);
            }
    }

    // raw words of one lane -> integers (no sign fix, no DC shift yet)
    static __device__ __forceinline__ void words_to_ints(const unsigned (&w)[NW], int (&v)[NC][NS]) {
#pragma unroll
        for (int s = 0; s < NS; s++)
#pragma unroll
            for (int c = 0; c < NC; c++) {
                const int e = NC * s + c;
                if constexpr (IN == IN_U8) v[c][s] = (w[e >> 2] >> (8 * (e & 3))) & 0xFF;
                else if constexpr (IN == IN_U16) v[c][s] = (e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xFFFF);
                else v[c][s] = (int)w[e];
            }
    }

    // integer tail: sign fix, DC shift, MCT, conversion to the working type (encoder.go:341-383,3698-3711,196-209)
    static __device__ __forceinline__ void finish_ints(int (&v)[NC][NS], const RawFmt& r, int dc, T (&out)[NC][NS]) {
        if constexpr (IN == IN_F32) {
#pragma unroll
            for (int c = 0; c < NC; c++)
#pragma unroll
                for (int s = 0; s < NS; s++) out[c][s] = (T)__int_as_float(v[c][s]);
            return;
        }
#pragma unroll
        for (int c = 0; c < NC; c++)
#pragma unroll
            for (int s = 0; s < NS; s++) {
                if constexpr (RAWIN && SG == 1) { if (v[c][s] >= r.sign_thresh) v[c][s] -= r.sign_sub; }
                v[c][s] -= dc;
            }
#pragma unroll
        for (int s = 0; s < NS; s++) {
            if constexpr (NC == 3 && MCT != MCTK_NONE) {
                mct_forward<WT, MCT>(v[0][s], v[1][s], v[2][s], out[0][s], out[1][s], out[2][s]);
            } else {
#pragma unroll
                for (int c = 0; c < NC; c++) out[c][s] = (T)v[c][s];
            }
        }
    }

    // Border strips: the lane span just outside the window (pixels [-NS, 0) and [w, w + NS)) is filled, inside the staged
    // row, with whole-sample-mirrored copies (dwt53.go / dwt97.go border cases == symmetric extension,
    // tests/test_oracle_mirror.py); lanes 0..NS-1 write the left span, lanes NS..2NS-1 the right one.  The TMA copy
    // of a border strip never covers those bytes, so the plain stores cannot race with a refill.
    static __device__ __forceinline__ void fix_halo(smem_t stage, int lane, bool fix_l, bool fix_r, int w, int vb) {
        if (lane < 2 * NS) {
            const bool left = lane < NS;
            const int k = left ? lane : lane - NS;
            if (left ? fix_l : fix_r) {
                const int pi = left ? -(k + 1) : w + k;
                const int so = mirror_fast(pi, w) * PB - vb, d_o = pi * PB - vb;
                unsigned char* st = smem_ptr(stage);
#pragma unroll
                for (int r = 0; r < 2 * RPS; r++) {  // every row of the stage: one pass, one warp barrier per four rows
                    unsigned char* row = st + r * ROWB;
                    if constexpr (UA) row += *(const int*)(row + ROWB - 4);  // this row's phase
#pragma unroll
                    for (int c = 0; c < PB / ES; c++) {
                        if constexpr (ES == 1) row[d_o + c] = row[so + c];
                        else if constexpr (ES == 2) *(unsigned short*)(row + d_o + 2 * c) = *(const unsigned short*)(row + so + 2 * c);
                        else *(unsigned*)(row + d_o + 4 * c) = *(const unsigned*)(row + so + 4 * c);
                    }
                }
            }
        }
        __syncwarp();
    }

    // one staged row of this lane as working values in scalar form (5/3, signed raw words, border lanes)
    // (every lane reads its span, also the lanes past the strip's right halo: the slot is theirs, the junk they compute is
    // never stored and never reaches a storing lane -- no branch in front of the loads, so they hoist freely)
    template <int LA = 16>
    static __device__ __forceinline__ void load_scalar(smem_t row, int lane_off, const RawFmt& raw, int dc, T (&out)[NC][NS]) {
        unsigned wv[NW];
        if constexpr (UA) fetch_ua(row, lane_off + (int)lds32(row + ROWB - 4), wv); else
        fetch(row + lane_off, wv);
        int v[NC][NS];
        words_to_ints(wv, v);
        finish_ints(v, raw, dc, out);
    }

    // the same row as column pairs of float32 (9/7): pair j = samples 2j, 2j+1
    template <int LA = 16>
    static __device__ __forceinline__ void load_pairs(smem_t row, int lane_off, const RawFmt& raw, int dc, float fmagic,
                                                      float one, float2 (&out)[NC][NP], float k0 = 0.f, float k1 = 0.f, float k2 = 0.f) {
        {
            unsigned wv[NW];
            if constexpr (UA) fetch_ua(row, lane_off + (int)lds32(row + ROWB - 4), wv); else
            fetch(row + lane_off, wv);
            if constexpr (XC == 3) {
                // one row of the float32 ICT (encoder.go:277-288): (r * k0 + g * k1) + b * k2 on exact float32(int) inputs
                const float2 nm = splat2(-fmagic);
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    float2 f[3];
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        unsigned p[2];
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const int e = 3 * (2 * j + k) + c;
                            if constexpr (IN == IN_U8) p[k] = __byte_perm(wv[e >> 2], 0x4B000000u, 0x7440u | (e & 3));
                            else p[k] = __byte_perm(wv[e >> 1], 0x4B000000u, (e & 1) ? 0x7432u : 0x7410u);
                        }
                        f[c] = add2(make_float2(__uint_as_float(p[0]), __uint_as_float(p[1])), nm);
                    }
                    out[0][j] = addp2(mul2(f[2], splat2(k2)), addp2(mul2(f[0], splat2(k0)), mul2(f[1], splat2(k1)), one), one);
                }
            } else if constexpr (IN == IN_F32) {
#pragma unroll
                for (int j = 0; j < NP; j++) out[0][j] = make_float2(__uint_as_float(wv[2 * j]), __uint_as_float(wv[2 * j + 1]));
            } else if constexpr (MAGIC) {
                // 0x4B000000 | v is the float 2^23 + v: one exact packed add removes 2^23 + dc
                float2 f[NC][NP];
                const float2 nm = splat2(-fmagic);
#pragma unroll
                for (int j = 0; j < NP; j++)
#pragma unroll
                    for (int c = 0; c < NC; c++) {
                        unsigned p[2];
#pragma unroll
                        for (int k = 0; k < 2; k++) {
                            const int e = NC * (2 * j + k) + c;
                            if constexpr (IN == IN_U8) p[k] = __byte_perm(wv[e >> 2], 0x4B000000u, 0x7440u | (e & 3));
                            else p[k] = __byte_perm(wv[e >> 1], 0x4B000000u, (e & 1) ? 0x7432u : 0x7410u);
                        }
                        f[c][j] = add2(make_float2(__uint_as_float(p[0]), __uint_as_float(p[1])), nm);
                    }
#pragma unroll
                for (int j = 0; j < NP; j++) {
                    if constexpr (NC == 3 && MCT == MCTK_ICT) {  // encoder.go:277-288 on exact float32(int) inputs
                        const float2 fr = f[0][j], fg = f[1][j], fb = f[2][j];
                        out[0][j] = addp2(mul2(fb, splat2(0.114f)), addp2(mul2(fr, splat2(0.299f)), mul2(fg, splat2(0.587f)), one), one);
                        out[1][j] = addp2(mul2(fb, splat2(0.5f)), addp2(mul2(fr, splat2(-0.16875f)), mul2(fg, splat2(-0.331260f)), one), one);
                        out[2][j] = addp2(mul2(fb, splat2(-0.08131f)), addp2(mul2(fr, splat2(0.5f)), mul2(fg, splat2(-0.41869f)), one), one);
                    } else {
#pragma unroll
                        for (int c = 0; c < NC; c++) out[c][j] = f[c][j];
                    }
                }
            } else {
                int v[NC][NS];
                float t[NC][NS];
                words_to_ints(wv, v);
                finish_ints(v, raw, dc, t);
#pragma unroll
                for (int c = 0; c < NC; c++)
#pragma unroll
                    for (int j = 0; j < NP; j++) out[c][j] = make_float2(t[c][2 * j], t[c][2 * j + 1]);
            }
        }
    }

    // exchange consumer: the lane's eight float32 samples of one row (two conflict-free 128-bit loads)
    static __device__ __forceinline__ void load_xch(smem_t p, float2 (&out)[NC][NP]) {
        const uint4 a = lds128(p), b = lds128(p + 512);
        out[0][0] = make_float2(__uint_as_float(a.x), __uint_as_float(a.y));
        out[0][1] = make_float2(__uint_as_float(a.z), __uint_as_float(a.w));
        if constexpr (NP == 4) {
            out[0][2] = make_float2(__uint_as_float(b.x), __uint_as_float(b.y));
            out[0][3] = make_float2(__uint_as_float(b.z), __uint_as_float(b.w));
        }
    }

    static __device__ __forceinline__ void store_vec(int* p, const int (&o)[NP], unsigned long long pol) {
#if J2K_RING_L2_HINTS
        if constexpr (NP == 4) { stg128_hint(p, o[0], o[1], o[2], o[3], pol); return; }
#endif
        (void)pol;
        if constexpr (NP == 4) *(int4*)p = make_int4(o[0], o[1], o[2], o[3]);
        else if constexpr (NP == 2) *(int2*)p = make_int2(o[0], o[1]);
        else *p = o[0];
    }

    // UA: nv of the lane's NP values lie inside the band (<= 0: none); SC = ints per store every band row of the job allows
    template <int SC>
    static __device__ __forceinline__ void store_ua(int* p, const int (&o)[NP], int nv) {
        if constexpr (SC >= 2) {
#pragma unroll
            for (int j = 0; j + 1 < NP; j += 2) {
                if (j + 1 < nv) *(int2*)(p + j) = make_int2(o[j], o[j + 1]);
                else if (j < nv) p[j] = o[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < NP; j++)
                if (j < nv) p[j] = o[j];
        }
    }

    // ... and of a strip that lies inside both bands with every storing lane (warp-uniform: all strips of a row but the last)
    template <int SC>
    static __device__ __forceinline__ void store_full(int* p, const int (&o)[NP]) {
        if constexpr (SC >= 2) {
#pragma unroll
            for (int j = 0; j + 1 < NP; j += 2) *(int2*)(p + j) = make_int2(o[j], o[j + 1]);
        } else {
#pragma unroll
            for (int j = 0; j < NP; j++) p[j] = o[j];
        }
    }

    // The job loop is instantiated per (load alignment, store class) of the general-alignment variant and chosen once per job.
    static __device__ __forceinline__ void run(const RingSeg& S, const RawFmt& raw, int item, int chunk, int strip, RingWarp& rw,
                                               int lane, FxPort* xp = nullptr, int xcomp = 0) {
        if constexpr (!UA) run_t<16, 4>(S, raw, item, chunk, strip, rw, lane, xp, xcomp);
        else {
            const int sc = min(min(S.st_cls[0], S.st_cls[1]), min(S.st_cls[2], S.st_cls[3]));
            if (sc >= 2) run_t<4, 2>(S, raw, item, chunk, strip, rw, lane); else run_t<4, 1>(S, raw, item, chunk, strip, rw, lane);
        }
    }

    template <int LA, int SC>
    static __device__ __forceinline__ void run_t(const RingSeg& S, const RawFmt& raw, int item, int chunk, int strip, RingWarp& rw,
                                               int lane, FxPort* xp = nullptr, int xcomp = 0) {
        (void)xp; (void)xcomp;
        const int w = S.w, h = S.h, py = S.py;
        const int hw = w - S.lw, hh = h - S.lh, lh = S.lh;
        const int kxs = strip * S.strip_pairs;
        const int kxe = min(kxs + S.strip_pairs, S.Kx);
        const int nl = (kxe - kxs + NP - 1) / NP + 2 * HLN;  // lanes that carry data
        const int kx0 = kxs - HLN * NP + lane * NP;
        const int i0w = 2 * (kxs - HLN * NP);
        const int m = (i0w * PB) & 15;
        const int vb = i0w * PB - m;  // virtual (16 B aligned) byte position of the slot start inside the row
        const int c0 = max(vb, 0);
        const int c1 = min(vb + ((m + nl * LB + 15) & ~15), S.row_bytes);
        const unsigned copy_bytes = (unsigned)(c1 - c0);
        const int dst_off = c0 - vb;
        const int lane_off = m + lane * LB;
        // warp-uniform: this strip touches a window border (UA: its right halo lane may reach past the window although another,
        // narrower strip follows)
        const bool fix_l = kxs == 0, fix_r = UA ? (2 * kxe + NS > w) : (kxe == S.Kx);
        const bool fix = fix_l || fix_r;
        // the window width is a multiple of 2 NP on the aligned path: a lane stores whole vectors or nothing; UA masks per band
        const bool st = lane >= HLN && kx0 < kxe && (UA || kx0 + NP <= hw);
        const int nv_l = S.lw - kx0, nv_h = hw - kx0;   // UA: values of this lane inside the low- / high-pass bands
        // UA, warp-uniform: every storing lane of this strip holds NP values of both bands (no masks needed)
        const bool full_w = kxs + ((kxe - kxs + NP - 1) / NP) * NP <= min(S.lw, hw);
        (void)nv_l; (void)nv_h; (void)full_w;

        const unsigned long long pol_load = S.pol_load, pol_ll = S.pol_ll, pol_band = S.pol_band;
        (void)pol_load; (void)pol_ll; (void)pol_band;
        const int dc = S.dc;
        const float fmagic = 8388608.0f + (float)dc;
        const float one = rw.one;
        // component-split first level: item = 3 * pixel item + component; exchange mode: pixel item, the consumer's component given
        const int comp = XCH == 1 ? xcomp : ((XC == 3 && XCH == 0) ? item % 3 : 0);
        if constexpr (XC == 3 && XCH == 0) item /= 3;
        // this component's row of the ICT matrix (encoder.go:277-288); warp-uniform
        const float k0 = comp == 0 ? 0.299f : (comp == 1 ? -0.16875f : 0.5f);
        const float k1 = comp == 0 ? 0.587f : (comp == 1 ? -0.331260f : -0.41869f);
        const float k2 = comp == 0 ? 0.114f : (comp == 1 ? 0.5f : -0.08131f);
        const unsigned char* src = S.x_base + S.x_off[item] * ES + c0;
        const long long pitch = S.x_row_bytes;

        const int ky0 = chunk * S.chunk_pairs;
        const int ky1 = min(ky0 + S.chunk_pairs, S.Ky);
        const int r_begin = 2 * (ky0 - LAG) - py;  // low-type row of iteration 0
        const int n_it = ky1 - ky0 + 2 * LAG;
        const int n_st = (n_it + RPS - 1) / RPS;
        // stages s_lo <= s < s_hi stage 2 RPS in-range, adjacent rows; the others mirror the row numbers
        const int s_lo = r_begin < 0 ? (2 * RPS - 1 - r_begin) / (2 * RPS) : 0;
        const int s_hi = (h - r_begin) / (2 * RPS);

        int* p_ll = (int*)S.ll.base + S.ll.off[item] + (long long)S.ll.y_off * S.ll.row_stride + S.ll.x_off + kx0;
        int* p_hl = (int*)S.hl.base + S.hl.off[item] + (long long)S.hl.y_off * S.hl.row_stride + S.hl.x_off + kx0;
        int* p_lh = (int*)S.lh_.base + S.lh_.off[item] + (long long)S.lh_.y_off * S.lh_.row_stride + S.lh_.x_off + kx0;
        int* p_hh = (int*)S.hh.base + S.hh.off[item] + (long long)S.hh.y_off * S.hh.row_stride + S.hh.x_off + kx0;
        const int rs_ll = S.ll.row_stride, rs_b = S.hl.row_stride;
        const long long cs_ll = S.ll.comp_stride, cs_b = S.hl.comp_stride;
        if constexpr (XC == 3 || XCH == 1) { p_ll += comp * cs_ll; p_hl += comp * cs_b; p_lh += comp * cs_b; p_hh += comp * cs_b; }
        // rows of the first storing iteration (low-type row ky0 - py, high-type row ky0); they advance one row per iteration
        p_ll += (long long)(ky0 - py) * rs_ll; p_hl += (long long)(ky0 - py) * rs_b;
        p_lh += (long long)ky0 * rs_b; p_hh += (long long)ky0 * rs_b;

        // producer cursor (warp-uniform): next stage to fill, its slot and the source pointer of its first row
        int pj = 0, pslot = 0;
        const unsigned char* psrc = src + (long long)r_begin * pitch;
        const smem_t dst_s = rw.ring + dst_off;
        // lane 0 stages the 2 RPS rows of stage pj (TMA bulk copies, completion on the stage barrier)
        auto issue = [&]() {
            if (elect_one()) {
                const smem_t bar = rw.bars + 8 * pslot;
                const smem_t dst = dst_s + pslot * STAGEB;
                if constexpr (UA) {
                    // every row starts at its own byte: copy from the 16-byte boundary below it, remember the phase in the slot
                    const unsigned char* rp[2 * RPS];
                    unsigned ph[2 * RPS], nb[2 * RPS], total = 0;
                    const bool inr = pj >= s_lo && pj < s_hi;
                    const int r0 = r_begin + 2 * RPS * pj;
#pragma unroll
                    for (int k = 0; k < 2 * RPS; k++) {
                        rp[k] = inr ? psrc + k * pitch : src + (long long)mirror_fast(r0 + k, h) * pitch;
                        ph[k] = (unsigned)((size_t)rp[k] & 15u);
                        nb[k] = (ph[k] + copy_bytes + 15u) & ~15u;
                        total += nb[k];
                        sts32(rw.ring + pslot * STAGEB + k * ROWB + ROWB - 4, ph[k]);
                    }
                    mbar_expect_tx(bar, total);
#pragma unroll
                    for (int k = 0; k < 2 * RPS; k++) bulk_g2s(dst + k * ROWB, rp[k] - ph[k], nb[k], bar J2K_POL_ARG(pol_load));
                } else {
                mbar_expect_tx(bar, 2 * RPS * copy_bytes);
                if (pj >= s_lo && pj < s_hi) {
#pragma unroll
                    for (int k = 0; k < 2 * RPS; k++) bulk_g2s(dst + k * ROWB, psrc + k * pitch, copy_bytes, bar J2K_POL_ARG(pol_load));
                } else {
                    const int r0 = r_begin + 2 * RPS * pj;
#pragma unroll
                    for (int k = 0; k < 2 * RPS; k++) bulk_g2s(dst + k * ROWB, src + (long long)mirror_fast(r0 + k, h) * pitch, copy_bytes, bar J2K_POL_ARG(pol_load));
                }
                }
            }
            pj++;
            psrc += 2 * RPS * pitch;
            pslot = (pslot + 1 == D) ? 0 : pslot + 1;
        };
        __syncwarp();
        if constexpr (XCH != 1) {
#pragma unroll 1
        for (int j = 0; j < D - 1 && j < n_st; j++) issue();
        }

        int cslot = 0;
        int xheld = -1;  // exchange consumer: the stage slot this warp is reading
        (void)xheld;
        smem_t stage = rw.ring;
        // start of a stage: the previous stage's slot is free once every lane is past its arithmetic -> refill, then wait
        auto next_stage = [&]() {
            __syncwarp();
            if constexpr (XCH == 1) {
                // hand the stage just read back to the producer, wait for the next one
                if (xheld >= 0 && lane == 0) mbar_arrive(xp->empty + 8 * xheld);
                const int slot = (int)(xp->cnt % (unsigned)xp->dx);
                mbar_wait(xp->full + 8 * slot, (xp->cnt / (unsigned)xp->dx) & 1u);
                stage = xp->data + slot * J2K_F3_STAGEB + comp * J2K_F3_COMPB;
                xheld = slot;
                xp->cnt++;
                return;
            }
            if (pj < n_st) issue();
            mbar_wait(rw.bars + 8 * cslot, (rw.phase >> cslot) & 1u);
            rw.phase ^= 1u << cslot;
            stage = rw.ring + cslot * STAGEB;
            cslot = (cslot + 1 == D) ? 0 : cslot + 1;
            if (fix) fix_halo(stage, lane, fix_l, fix_r, w, vb);
        };
        const smem_t lane_s = rw.ring + lane_off;
        if constexpr (XCH == 2) {
            // Producer of the exchange ring: every raw stage is converted ONCE for all three components (unpack, - DC, the three
            // rows of the float32 ICT, encoder.go:277-288, same operation order as the component-split jobs) and handed to the
            // three single-component consumers of the CTA.
            const float2 nm = splat2(-fmagic);
#pragma unroll 1
            for (int sj = 0; sj < n_st; sj++) {
                next_stage();
                const int slot = (int)(xp->cnt % (unsigned)xp->dx);
                mbar_wait(xp->empty + 8 * slot, ((xp->cnt / (unsigned)xp->dx) & 1u) ^ 1u);
                const smem_t xb = xp->data + slot * J2K_F3_STAGEB + lane * 16;
#pragma unroll
                for (int r = 0; r < 2 * RPS; r++) {
                    unsigned wv[NW];
                    fetch(stage + r * ROWB + lane_off, wv);
                    float2 y[NP], cb[NP], cr[NP];
#pragma unroll
                    for (int j = 0; j < NP; j++) {
                        float2 f[3];
#pragma unroll
                        for (int c = 0; c < 3; c++) {
                            unsigned p[2];
#pragma unroll
                            for (int k = 0; k < 2; k++) {
                                const int e = 3 * (2 * j + k) + c;
                                if constexpr (IN == IN_U8) p[k] = __byte_perm(wv[e >> 2], 0x4B000000u, 0x7440u | (e & 3));
                                else p[k] = __byte_perm(wv[e >> 1], 0x4B000000u, (e & 1) ? 0x7432u : 0x7410u);
                            }
                            f[c] = add2(make_float2(__uint_as_float(p[0]), __uint_as_float(p[1])), nm);
                        }
                        y[j] = addp2(mul2(f[2], splat2(0.114f)), addp2(mul2(f[0], splat2(0.299f)), mul2(f[1], splat2(0.587f)), one), one);
                        cb[j] = addp2(mul2(f[2], splat2(0.5f)), addp2(mul2(f[0], splat2(-0.16875f)), mul2(f[1], splat2(-0.331260f)), one), one);
                        cr[j] = addp2(mul2(f[2], splat2(-0.08131f)), addp2(mul2(f[0], splat2(0.5f)), mul2(f[1], splat2(-0.41869f)), one), one);
                    }
                    const smem_t xr = xb + r * J2K_F3_ROWB;
                    sts128f(xr, y[0], y[1]); sts128f(xr + 512, y[2], y[3]);
                    sts128f(xr + J2K_F3_COMPB, cb[0], cb[1]); sts128f(xr + J2K_F3_COMPB + 512, cb[2], cb[3]);
                    sts128f(xr + 2 * J2K_F3_COMPB, cr[0], cr[1]); sts128f(xr + 2 * J2K_F3_COMPB + 512, cr[2], cr[3]);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(xp->full + 8 * slot);
                xp->cnt++;
            }
        } else if constexpr (WT == 97) {
            const float2 A2 = splat2(J2K_ALPHA), B2 = splat2(J2K_BETA), G2 = splat2(J2K_GAMMA), D2 = splat2(J2K_DELTA);
            // quantizer constants per row pair: E positions hold (LL, LH), O positions hold (HL, HH)
            const float2 rcpE = S.rcpE, nstE = S.nstE, rcpO = S.rcpO, nstO = S.nstO;
            const bool raw_ll = S.q[0].mode != Q_QUANT, raw_hl = S.q[1].mode != Q_QUANT, raw_lh = S.q[2].mode != Q_QUANT,
                       raw_hh = S.q[3].mode != Q_QUANT;
            // the shape of every level but the last: LL stays float32 (next level's input), the details are quantized
            const bool std_modes = raw_ll && !raw_hl && !raw_lh && !raw_hh;
            struct VState { float2 pe[NC][NP], po[NC][NP], s1p[NC][NP], d1p[NC][NP], d2p[NC][NP]; };
            VState sa, sb;
#pragma unroll
            for (int c = 0; c < NC; c++)
#pragma unroll
                for (int j = 0; j < NP; j++) { sa.pe[c][j] = sa.po[c][j] = sa.s1p[c][j] = sa.d1p[c][j] = sa.d2p[c][j] = make_float2(0.f, 0.f); }
            // one iteration: consumes the vertical window state `in`, leaves the advanced state in `out`
            auto body = [&](int it, const int half, const VState& in, VState& out) {
                if (half == 0) next_stage();
                const smem_t row_e = stage + half * 2 * SROWB;
                const smem_t row_o = row_e + SROWB;

                if constexpr (XCH == 1) {
                    load_xch(row_e + lane * 16, out.pe);
                    load_xch(row_o + lane * 16, out.po);
                } else {
                load_pairs<LA>(row_e, lane_off, raw, dc, fmagic, one, out.pe, k0, k1, k2);
                load_pairs<LA>(row_o, lane_off, raw, dc, fmagic, one, out.po, k0, k1, k2);
                }
                // vertical lifting on column pairs; the finished (low, high) row values of column s land in Q[c][s]
                float2 Q[NC][NS];
#pragma unroll
                for (int c = 0; c < NC; c++)
#pragma unroll
                    for (int j = 0; j < NP; j++) {
                        const float2 d1 = lift97x2(in.po[c][j], in.pe[c][j], out.pe[c][j], A2, one);
                        const float2 s1 = lift97x2(in.pe[c][j], in.d1p[c][j], d1, B2, one);
                        const float2 d2 = lift97x2(in.d1p[c][j], in.s1p[c][j], s1, G2, one);
                        const float2 s2 = lift97x2(in.s1p[c][j], in.d2p[c][j], d2, D2, one);
                        Q[c][2 * j] = make_float2(__fmul_rn(s2.x, J2K_INVK), __fmul_rn(d2.x, J2K_K));
                        Q[c][2 * j + 1] = make_float2(__fmul_rn(s2.y, J2K_INVK), __fmul_rn(d2.y, J2K_K));
                        out.d1p[c][j] = d1; out.s1p[c][j] = s1; out.d2p[c][j] = d2;
                    }
                if (it < 2 * LAG) return;  // warm-up of the vertical window (warp-uniform)
                const int ky = ky0 + it - 2 * LAG;
                const int yl = ky - py, yh = ky;
                const bool row_l = st && yl >= 0 && yl < lh, row_h = st && yh < hh;
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    // horizontal lifting of both rows at once: E_j = Q[2j], O_j = Q[2j+1]
                    float2 E[NP + 1], O[NP + 1];  // O[0] = previous lane's last, E[NP] = next lane's first
#pragma unroll
                    for (int j = 0; j < NP; j++) { E[j] = Q[c][2 * j]; O[j + 1] = Q[c][2 * j + 1]; }
                    // each step as three passes over the NP independent pairs (sum, scale, accumulate) so that the
                    // dependent triples of different pairs interleave in the instruction stream
                    float2 T[NP];
                    E[NP] = shfl_down2(E[0]);
#pragma unroll
                    for (int j = 0; j < NP; j++) T[j] = add2(E[j], E[j + 1]);
#pragma unroll
                    for (int j = 0; j < NP; j++) T[j] = mul2(T[j], A2);
#pragma unroll
                    for (int j = 0; j < NP; j++) O[j + 1] = addp2(T[j], O[j + 1], one);
                    O[0] = shfl_up2(O[NP]);
#pragma unroll
                    for (int j = 0; j < NP; j++) T[j] = add2(O[j], O[j + 1]);
#pragma unroll
                    for (int j = 0; j < NP; j++) T[j] = mul2(T[j], B2);
#pragma unroll
                    for (int j = 0; j < NP; j++) E[j] = addp2(T[j], E[j], one);
                    E[NP] = shfl_down2(E[0]);
#pragma unroll
                    for (int j = 0; j < NP; j++) T[j] = add2(E[j], E[j + 1]);
#pragma unroll
                    for (int j = 0; j < NP; j++) T[j] = mul2(T[j], G2);
#pragma unroll
                    for (int j = 0; j < NP; j++) O[j + 1] = addp2(T[j], O[j + 1], one);
                    O[0] = shfl_up2(O[NP]);
#pragma unroll
                    for (int j = 0; j < NP; j++) T[j] = add2(O[j], O[j + 1]);
#pragma unroll
                    for (int j = 0; j < NP; j++) T[j] = mul2(T[j], D2);
#pragma unroll
                    for (int j = 0; j < NP; j++) E[j] = addp2(T[j], E[j], one);
                    int q_ll[NP], q_hl[NP], q_lh[NP], q_hh[NP];
                    // q = rint(c / step') from the correctly rounded reciprocal (Markstein), both rows at once
                    if (std_modes) {
#pragma unroll
                        for (int j = 0; j < NP; j++) {
                            const float2 ve = mul2(E[j], splat2(J2K_INVK)), vo = mul2(O[j + 1], splat2(J2K_K));
                            const float2 e0 = mul2(ve, rcpE), o0 = mul2(vo, rcpO);
                            const float2 re = fma2(fma2(e0, nstE, ve), rcpE, e0), ro = fma2(fma2(o0, nstO, vo), rcpO, o0);
                            q_ll[j] = __float_as_int(ve.x);
                            q_lh[j] = __float2int_rn(re.y);
                            q_hl[j] = __float2int_rn(ro.x);
                            q_hh[j] = __float2int_rn(ro.y);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < NP; j++) {
                            const float2 ve = mul2(E[j], splat2(J2K_INVK)), vo = mul2(O[j + 1], splat2(J2K_K));
                            const float2 e0 = mul2(ve, rcpE), o0 = mul2(vo, rcpO);
                            const float2 re = fma2(fma2(e0, nstE, ve), rcpE, e0), ro = fma2(fma2(o0, nstO, vo), rcpO, o0);
                            q_ll[j] = raw_ll ? __float_as_int(ve.x) : __float2int_rn(re.x);
                            q_lh[j] = raw_lh ? __float_as_int(ve.y) : __float2int_rn(re.y);
                            q_hl[j] = raw_hl ? __float_as_int(vo.x) : __float2int_rn(ro.x);
                            q_hh[j] = raw_hh ? __float_as_int(vo.y) : __float2int_rn(ro.y);
                        }
                    }
                    if constexpr (UA) {
                        if (full_w) {
                            if (row_l) { store_full<SC>(p_ll, q_ll); store_full<SC>(p_hl, q_hl); }
                            if (row_h) { store_full<SC>(p_lh, q_lh); store_full<SC>(p_hh, q_hh); }
                        } else {
                        if (row_l) { store_ua<SC>(p_ll, q_ll, nv_l); store_ua<SC>(p_hl, q_hl, nv_h); }
                        if (row_h) { store_ua<SC>(p_lh, q_lh, nv_l); store_ua<SC>(p_hh, q_hh, nv_h); }
                        }
                    } else {
                    if (row_l) {
                        store_vec(p_ll + c * cs_ll, q_ll, pol_ll);
                        store_vec(p_hl + c * cs_b, q_hl, pol_band);
                    }
                    if (row_h) {
                        store_vec(p_lh + c * cs_b, q_lh, pol_band);
                        store_vec(p_hh + c * cs_b, q_hh, pol_band);
                    }
                    }
                }
                p_ll += rs_ll; p_hl += rs_b; p_lh += rs_b; p_hh += rs_b;
            };
            // two iterations per trip with the window state ping-ponging between sa and sb: no register shuffling
            if constexpr (NC == 3 && WT == 97 && J2K_RGB_SINGLE_BODY) {
                // three 9/7 components: ONE copy of the body and a state move per iteration -- the two-copy form of these
                // variants is ~60 KB of code and misses the instruction cache (stall_no_inst 11 % -> +15 % on C3 / C5)
#pragma unroll 1
                for (int it = 0; it < n_it; it++) {
                    body(it, RPS == 1 ? 0 : (it & 1), sa, sb);
                    sa = sb;
                }
            } else
#pragma unroll 1
            for (int it = 0; it < n_it; it += 2) {
                body(it, 0, sa, sb);
                if (it + 1 >= n_it) break;
                body(it + 1, 1, sb, sa);
            }
            if constexpr (XCH == 1) {  // the last stage goes back to the producer
                __syncwarp();
                if (xheld >= 0 && lane == 0) mbar_arrive(xp->empty + 8 * xheld);
            }
        } else {
            const int sh_ll = S.q[0].shift, sh_hl = S.q[1].shift, sh_lh = S.q[2].shift, sh_hh = S.q[3].shift;
            struct VState { int pe[NC][NS], po[NC][NS], d1p[NC][NS]; };
            VState sa, sb;
#pragma unroll
            for (int c = 0; c < NC; c++)
#pragma unroll
                for (int s = 0; s < NS; s++) { sa.pe[c][s] = 0; sa.po[c][s] = 0; sa.d1p[c][s] = 0; }
            // two iterations per trip with the window state ping-ponging between sa and sb (as in the 9/7 loop)
            auto body = [&](int it, const int half, const VState& in, VState& out) {
                if (half == 0) next_stage();
                const smem_t row_e = stage + half * 2 * ROWB;
                const smem_t row_o = row_e + ROWB;

                int lo[NC][NS], hi[NC][NS];
                load_scalar<LA>(row_e, lane_off, raw, dc, out.pe);
                load_scalar<LA>(row_o, lane_off, raw, dc, out.po);
#pragma unroll
                for (int c = 0; c < NC; c++)
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        const int d = in.po[c][s] - ((in.pe[c][s] + out.pe[c][s]) >> 1);
                        const int sv = in.pe[c][s] + ((in.d1p[c][s] + d + 2) >> 2);
                        lo[c][s] = sv; hi[c][s] = d;
                        out.d1p[c][s] = d;
                    }
                if (it < 2 * LAG) return;
                const int ky = ky0 + it - 2 * LAG;
                const int yl = ky - py, yh = ky;
                const bool row_l = st && yl >= 0 && yl < lh, row_h = st && yh < hh;
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    Slow::hlift(lo[c], false);
                    Slow::hlift(hi[c], false);
                    int q_a[NP], q_b[NP];
                    if (row_l) {
#pragma unroll
                        for (int j = 0; j < NP; j++) { q_a[j] = (int)((unsigned)lo[c][2 * j] << sh_ll); q_b[j] = (int)((unsigned)lo[c][2 * j + 1] << sh_hl); }
                        if constexpr (UA) { if (full_w) { store_full<SC>(p_ll, q_a); store_full<SC>(p_hl, q_b); } else { store_ua<SC>(p_ll, q_a, nv_l); store_ua<SC>(p_hl, q_b, nv_h); } } else {
                        store_vec(p_ll + c * cs_ll, q_a, pol_ll);
                        store_vec(p_hl + c * cs_b, q_b, pol_band);
                        }
                    }
                    if (row_h) {
#pragma unroll
                        for (int j = 0; j < NP; j++) { q_a[j] = (int)((unsigned)hi[c][2 * j] << sh_lh); q_b[j] = (int)((unsigned)hi[c][2 * j + 1] << sh_hh); }
                        if constexpr (UA) { if (full_w) { store_full<SC>(p_lh, q_a); store_full<SC>(p_hh, q_b); } else { store_ua<SC>(p_lh, q_a, nv_l); store_ua<SC>(p_hh, q_b, nv_h); } } else {
                        store_vec(p_lh + c * cs_b, q_a, pol_band);
                        store_vec(p_hh + c * cs_b, q_b, pol_band);
                        }
                    }
                }
                p_ll += rs_ll; p_hl += rs_b; p_lh += rs_b; p_hh += rs_b;
            };
            if constexpr (NC == 3 && WT == 97 && J2K_RGB_SINGLE_BODY) {
                // three 9/7 components: ONE copy of the body and a state move per iteration -- the two-copy form of these
                // variants is ~60 KB of code and misses the instruction cache (stall_no_inst 11 % -> +15 % on C3 / C5)
#pragma unroll 1
                for (int it = 0; it < n_it; it++) {
                    body(it, RPS == 1 ? 0 : (it & 1), sa, sb);
                    sa = sb;
                }
            } else
#pragma unroll 1
            for (int it = 0; it < n_it; it += 2) {
                body(it, 0, sa, sb);
                if (it + 1 >= n_it) break;
                body(it + 1, 1, sb, sa);
            }
        }
        (void)lane_s;
    }
};

// ------------------------------------------------------------------ job loop shared by both directions

struct RingJob { int seg, item, chunk, strip; };

// Claims the next job (warp-uniform result); returns false when the list is exhausted.
// A CTA claims J2K_RING_WARPS consecutive jobs at a time: its warps then walk ADJACENT strips of the same rows together,
// so the band rows (and the staged source rows) are touched in runs of 4 x 480 B instead of isolated 480 B pieces.  The
// store side of this kernel is what bounds it (with the loads removed it runs at the same speed), and
// the longer runs are worth +4 % on C2.  Per-warp claiming (0) is kept for the emulator, whose warps run one by one.
#ifndef J2K_RING_CTA_CLAIM
#define J2K_RING_CTA_CLAIM 1
#endif
// Slice starts of the group-pipelined order, staged in shared memory once per CTA (the emulator reads them in place).
#ifdef J2K_EMU
#define J2K_RING_SCHED_DECL(A, name) const int* name = (A).slice_begin
#else
#define J2K_RING_SCHED_DECL(A, name)                                                              \
    __shared__ int name##_s[J2K_RING_MAXSLICE];                                                   \
    for (int i_ = (int)threadIdx.x; i_ < (A).nslice; i_ += (int)blockDim.x) name##_s[i_] = (A).slice_begin[i_]; \
    __syncthreads();                                                                              \
    const int* name = name##_s
#endif
struct ClaimQ { unsigned cnt; int base[16]; unsigned tag[16]; };  // J2K_RING_CTA_CLAIM == 2: ticket counter, posted group bases
#ifndef J2K_EMU
__device__ __forceinline__ void claimq_init(ClaimQ& q) {
    if (threadIdx.x < 16) { q.base[threadIdx.x] = 0; q.tag[threadIdx.x] = 0; }
    if (threadIdx.x == 0) q.cnt = 0;
    __syncthreads();
}
#endif
template <bool CTA>
__device__ __forceinline__ bool ring_claim(const RingArgs& A, const int* sb, int lane, RingJob& J, ClaimQ* cq = nullptr) {
    int job = 0;
#if J2K_RING_CTA_CLAIM == 2 && !defined(J2K_EMU)
    if constexpr (CTA) {
        // Group claiming without a CTA barrier: the warps of a CTA take tickets from a shared counter; the warp that draws the
        // first ticket of a group of NW claims NW consecutive jobs with ONE global atomic and posts the base, the others of the
        // group pick it up (they are at most one atomic round trip behind).  The CTA still walks adjacent strips of the same
        // rows, but a warp that finishes early no longer waits for the slowest one (stall_barrier 4-7 % of the forward kernels).
        if (lane == 0) {
            const unsigned nw = blockDim.x >> 5;
            const unsigned idx = atomicAdd(&cq->cnt, 1u), grp = idx / nw, slot = idx - grp * nw;
            volatile int* rb = cq->base + (grp & 15u);
            volatile unsigned* rt = cq->tag + (grp & 15u);
            if (slot == 0) {
                *rb = (int)atomicAdd(A.ctl, nw);
                __threadfence_block();
                *rt = grp + 1u;
            } else {
                while (*rt != grp + 1u) {}
                __threadfence_block();
            }
            job = *rb + (int)slot;
        }
        job = __shfl_sync(0xffffffffu, job, 0);
    } else
#elif J2K_RING_CTA_CLAIM && !defined(J2K_EMU)
    if constexpr (CTA) {
    __shared__ int s_job;
    __syncthreads();  // every warp has read the previous group's base
    if (threadIdx.x == 0) s_job = (int)atomicAdd(A.ctl, (unsigned)(blockDim.x >> 5));
    __syncthreads();
    job = __shfl_sync(0xffffffffu, s_job + (int)(threadIdx.x >> 5), 0);  // (the shuffle keeps it warp-uniform for the compiler)
    } else
#endif
    {
        if (lane == 0) job = (int)atomicAdd(A.ctl, 1u);
        job = __shfl_sync(0xffffffffu, job, 0);
    }
    if (job >= A.total_jobs) return false;
    int k = 0, local, item0 = 0;
    if (A.nslice > 0) {
        int lo = 0, hi = A.nslice - 1;  // last slice that starts at or before this job
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (sb[mid] <= job) lo = mid; else hi = mid - 1;
        }
        const int2 info = A.slice_info[lo];
        k = info.x; item0 = info.y;
        local = job - sb[lo];
    } else {
        while (k + 1 < A.nseg && job >= A.seg[k].job_end) k++;
        local = job - A.seg[k].job_begin;
    }
    const int ns = A.seg[k].nstrips, nc = A.seg[k].nchunks;
    J.seg = k;
    J.strip = local % ns;
    J.chunk = (local / ns) % nc;
    J.item = item0 + local / (ns * nc);
    return true;
}

// Acquire: the producers of this job's input have published their stores.
__device__ __forceinline__ void ring_wait_dep(const RingArgs& A, const RingSeg& S, int item, int lane) {
    if (S.dep_seg < 0) return;
    if (lane == 0) {
        const int mul = S.dep_mul > 1 ? S.dep_mul : 1;
        const unsigned* d = A.ctl + A.seg[S.dep_seg].done_base + (item / S.dep_div) * mul;
        // relaxed polling (no L1 invalidation per probe), then one acquire load per counter: it synchronizes with the
        // producers' release increments without the two MEMBAR.ALL.GPU + ERRBAR a fence.acq_rel.gpu + fence.proxy.async cost
        for (int k = 0; k < mul; k++) {
            while (ld_relaxed(d + k) < (unsigned)S.dep_target) backoff();
            (void)ld_acquire(d + k);
        }
        fence_proxy_async_global();  // the staged reads that follow go through the async proxy
    }
    __syncwarp();
}

// Release: every lane's stores of this job happen-before the counter increment.  Segments nobody waits on
// (has_waiters == 0: the last forward level, the final inverse level) publish nothing.
__device__ __forceinline__ void ring_signal(const RingArgs& A, const RingSeg& S, int item, int lane) {
    if (!S.has_waiters) return;
    __syncwarp();
    if (lane == 0) red_release_add(A.ctl + S.done_base + item, 1u);
}

// The last warp to retire puts the control block back to zero for the next launch.
__device__ __forceinline__ void ring_retire(const RingArgs& A, int lane) {
    const unsigned total_warps = gridDim.x * (blockDim.x >> 5);
    unsigned last = 0;
    if (lane == 0) {
        __threadfence();
        last = atomicAdd(A.ctl + 1, 1u) == total_warps - 1 ? 1u : 0u;
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
        for (int i = lane; i < A.n_ctl; i += 32) A.ctl[i] = 0;
        __threadfence();
    }
}

__device__ __forceinline__ void ring_warp_init(unsigned char* smem, RingWarp& rw, int lane, int ring_bytes) {
    const int wib = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler too: the staging addresses stay in uniform registers
    rw.ring = smem_handle(smem + wib * (ring_bytes + J2K_RING_MAXD * 8));
    rw.bars = rw.ring + ring_bytes;
    rw.phase = 0;
    if (lane == 0) {
        for (int s = 0; s < J2K_RING_MAXD; s++) mbar_init(rw.bars + 8 * s, 1);
        mbar_fence_init();
    }
    __syncwarp();
}

// WT: 53 / 97.  IN1 / NC1 / MCT1 / SG1: the image-side variant of segments with first == 1
// (raw interleaved words, or planar int32 / float32 for the wavelet-package API); deeper levels always
// read planar working-type LL planes.
#ifdef J2K_RING_MAXNREG
#define J2K_RING_BOUNDS __maxnreg__(J2K_RING_MAXNREG)
#else
#ifndef J2K_RING_MINB_RGB97
#define J2K_RING_MINB_RGB97 2  // the 3-component 9/7 level-1 variant carries 3x the window state: 2 CTAs/SM, no spills
#endif
#define J2K_RING_BOUNDS __launch_bounds__(J2K_RING_WARPS * 32, (WT == 97 && NC1 == 3 && NP1 == 2) ? J2K_RING_MINB_RGB97 : J2K_RING_MINB)
#endif
template <int WT, int NP1, int NC1, int IN1, int MCT1, int SG1, int UA = 0>
__global__ void J2K_RING_BOUNDS fwd_ring_kernel(const __grid_constant__ RingArgs A) {
    J2K_SMEM_DECL(smem);
    const int lane = threadIdx.x & 31;
    RingWarp rw;
    ring_warp_init(smem, rw, lane, fwd_ring_bytes<UA>());
    rw.one = A.one;
    RingJob J;
    J2K_RING_SCHED_DECL(A, sched);
#if J2K_RING_CTA_CLAIM == 2 && !defined(J2K_EMU)
    __shared__ ClaimQ cq;
    claimq_init(cq);
    ClaimQ* const cqp = &cq;
#else
    ClaimQ* const cqp = nullptr;
#endif
    while (ring_claim<true>(A, sched, lane, J, cqp)) {
        const RingSeg& S = A.seg[J.seg];
        ring_wait_dep(A, S, J.item, lane);
        if (S.first) {
            // NC1 == 3 with NP1 == 4 names the component-split variant (one component per job)
            if constexpr (NC1 == 3 && NP1 == 4) FwdRing<WT, 4, 1, IN1, MCT1, SG1, 3>::run(S, A.raw, J.item, J.chunk, J.strip, rw, lane);
            else FwdRing<WT, NP1, NC1, IN1, MCT1, SG1, 1, UA>::run(S, A.raw, J.item, J.chunk, J.strip, rw, lane);
        } else FwdRing<WT, 4, 1, (WT == 53 ? IN_I32 : IN_F32), MCTK_NONE, 0, 1, UA>::run(S, A.raw, J.item, J.chunk, J.strip, rw, lane);
        ring_signal(A, S, J.item, lane);
    }
    ring_retire(A, lane);
}

}  // namespace j2k

namespace j2k {

// ------------------------------------------------------------------ inverse job
//
// Mirror image of FwdRing.  One iteration stages, per component, the four band rows that one pair of
// interleaved rows needs (LL/HL row yl, LH/HH row yh), dequantizes, runs the horizontal synthesis of both rows
// at once (packed f32x2 for 9/7: the pair is (low-type row, high-type row)), then the vertical synthesis as a
// register sliding window over column pairs, and stores the two finished rows of pair t - LAG: planar working
// type for an intermediate level, or rounded / inverse-MCT / DC-shifted / clamped / packed pixels for level 1.
// Reference order per level: rows, then columns (dwt53.go:318-354, dwt97.go:360-384).

// staging bytes per warp of the inverse kernel whose level-1 variant is (WT, NC1)
template <int WT, int NC1, int UA = 0> __host__ __device__ constexpr int inv_ring_bytes() {
    return UA ? J2K_INV_RING_BYTES_UA : ((WT == 53 && NC1 == 1) ? J2K_INV_RING_BYTES_53 : J2K_INV_RING_BYTES);
}
template <int WT, int NC1, int UA = 0> __host__ __device__ constexpr int inv_cta_smem() { return J2K_RING_WARPS * (inv_ring_bytes<WT, NC1, UA>() + J2K_RING_MAXD * 8); }

// Exchange ring of the three-producer level-1 inverse (inv3w_kernel): DX stages, a stage = one row pair of all three
// components as half-even-rounded float32 samples, [component][row e / o][half][lane] x 16 bytes (conflict-free 128-bit
// accesses).  Producers (one warp per component) fill a stage and arrive on full[stage]; the consumer warp drains it and
// arrives on empty[stage].
#ifndef J2K_X3_DX
#define J2K_X3_DX 4
#endif
#define J2K_X3_STAGEB (3 * 2 * 2 * 32 * 16)
struct XchPort {
    smem_t data, full, empty;
    unsigned cnt;   // stages produced / consumed so far by this warp (all warps of the CTA count the same stages)
    int dx;         // stages in the ring
    __device__ __forceinline__ void put(int comp, int lane, const float2 (&xe)[4], const float2 (&xo)[4]) {
        const int slot = (int)(cnt % (unsigned)dx);
        mbar_wait(empty + 8 * slot, ((cnt / (unsigned)dx) & 1u) ^ 1u);
        const float2 MG = splat2(12582912.0f), NMG = splat2(-12582912.0f);
        const smem_t b = data + slot * J2K_X3_STAGEB + comp * 2048 + lane * 16;
        sts128f(b, add2(add2(xe[0], MG), NMG), add2(add2(xe[1], MG), NMG));
        sts128f(b + 512, add2(add2(xe[2], MG), NMG), add2(add2(xe[3], MG), NMG));
        sts128f(b + 1024, add2(add2(xo[0], MG), NMG), add2(add2(xo[1], MG), NMG));
        sts128f(b + 1536, add2(add2(xo[2], MG), NMG), add2(add2(xo[3], MG), NMG));
        __syncwarp();
        if (lane == 0) mbar_arrive(full + 8 * slot);
        cnt++;
    }
};

// UA = 1: general-alignment variant of the inverse (see FwdRing): band rows staged from any byte position with their own
// phase, lanes fetch with the widest load the phases allow, pixel / LL rows stored with the widest access their alignment
// allows and masked at the window's right edge; odd window widths (low band one sample wider than the high band) included.
// XCH = 1: the job is one COMPONENT of a three-component level 1 (inv3w_kernel): the finished rows go to the CTA's exchange
// ring instead of memory.
template <int WT, int NP, int NC, int OUT, int MCT, int RB = J2K_INV_RING_BYTES, int UA = 0, int XCH = 0>
struct InvRing {
    typedef typename Wt<WT>::T T;
    static constexpr int LAG = Wt<WT>::LAG;
    // Halo-free strips (5/3): the inverse runs the horizontal synthesis FIRST, on staged band rows, so the one band sample
    // a lane needs from each neighbour is read straight from the staged row (for the strip's outer lanes it belongs to the
    // next strip, or is the mirrored sample fix_halo wrote) instead of coming from a halo lane through a shuffle: all 32
    // lanes store, a 512-wide frame is two full strips instead of three of 22 lanes, and strips are 128 pairs = whole
    // 512-byte band-row pieces.  The staged slot still carries one lane width of extra samples per side (SHL).
    static constexpr bool HF = J2K_INV_HALO_FREE && WT == 53;
    static constexpr int HLN = HF ? 0 : (Wt<WT>::HALO + NP - 1) / NP;  // halo lanes per side
    static constexpr int SHL = HF ? 1 : 0;                               // staged-only lane widths per side
    static constexpr int NS = 2 * NP;
    static constexpr bool FINAL = (OUT == IN_U8 || OUT == IN_U16);
    static constexpr int LB = NP * 4;          // bytes per lane per band row
    static constexpr int ROWB = (32 + 2 * SHL) * LB + 32 + (UA ? 32 : 0);  // staged band row slot (UA: + the row's phase, its value in the last word)
    static constexpr int NROWS = 4 * NC;       // LL, HL, LH, HH per component
    static_assert(!UA || (NC == 1 && NP == 4), "general alignment: single-component jobs");
    static constexpr int RPS = (NC == 1) ? 2 : 1;  // row pairs (= loop iterations) per stage
    static constexpr int PAIRB = NROWS * ROWB;     // staged bytes of one row pair
    static constexpr int STAGEB = RPS * PAIRB;
    static constexpr int D = (RB / STAGEB) > J2K_RING_MAXD ? J2K_RING_MAXD : (RB / STAGEB);
    static_assert(D >= 2, "ring too small for this stage size");
    static_assert(NC == 1 || FINAL, "3-component jobs write interleaved pixels");

    static __device__ __forceinline__ void fetch(smem_t p, int (&q)[NP]) {
        if constexpr (NP == 4) { uint4 v = lds128(p); q[0] = (int)v.x; q[1] = (int)v.y; q[2] = (int)v.z; q[3] = (int)v.w; }
        else if constexpr (NP == 2) { uint2 v = lds64(p); q[0] = (int)v.x; q[1] = (int)v.y; }
        else q[0] = (int)lds32(p);
    }

    // the lane's NP band samples at any 4-byte phase (UA: band samples are ints).  Four 32-bit loads: two aligned 128-bit
    // loads plus a warp-uniform word select (the forward kernel's form) were measured 10 - 16 % slower here.
    static __device__ __forceinline__ void fetch_ua(smem_t row, int off, int (&q)[NP]) {
#pragma unroll
        for (int k = 0; k < NP; k++) q[k] = (int)lds32(row + off + 4 * k);
    }
    static __device__ __forceinline__ int row_phase(smem_t row) { return UA ? (int)lds32(row + ROWB - 4) : 0; }

    // Border strips: band samples just outside the window, mirrored in the interleaved domain (low stays low, high
    // stays high): lanes 0..NP-1 write the left span, lanes NP..2NP-1 the right one, in every staged band row.
    // An odd window has one high-pass sample less than low-pass samples: the high rows' outside starts one index earlier.
    static __device__ __forceinline__ void fix_halo(smem_t stage, int lane, bool fix_l, bool fix_r, int w, int bw, int vb) {
        if (lane < 2 * NP) {
            const bool left = lane < NP;
            const int k = left ? lane : lane - NP;
            if (left ? fix_l : fix_r) {
                const int kx = left ? -(k + 1) : bw + k;  // band index outside [0, bw)
                const int kxh = (UA && !left) ? kx - (2 * bw - w) : kx;  // the same for the high-pass rows (w - bw samples)
                const int src_lo = mirror_fast(2 * kx, w) >> 1, src_hi = (mirror_fast(2 * kxh + 1, w) - 1) >> 1;
                unsigned char* st = smem_ptr(stage);
#pragma unroll
                for (int r = 0; r < RPS * NROWS; r++) {  // every band row of the stage (PAIRB == NROWS * ROWB)
                    const int src = (r & 1) ? src_hi : src_lo;  // rows 1, 3 (HL, HH) are horizontally high-pass
                    const int dst = (r & 1) ? kxh : kx;
                    unsigned char* row = st + r * ROWB;
                    if constexpr (UA) row += *(const int*)(row + ROWB - 4);  // this row's phase
                    *(unsigned*)(row + dst * 4 - vb) = *(const unsigned*)(row + src * 4 - vb);
                }
            }
        }
        __syncwarp();
    }

    // UA stores: the variant (XA = bytes every destination row is aligned to: 8, 4 or 1 = element-wise) is a template
    // parameter of the job loop, chosen once per job.  n valid ints of q at p (planar destinations: LL planes, GetImageData)
    template <int XA>
    static __device__ __forceinline__ void store_ints_ua(int* p, const int (&q)[NS], int n) {
        if (n >= NS) {   // every lane of a strip that does not touch the window's right edge: no masks
#pragma unroll
            for (int j = 0; j < NS; j += 2) {
                if constexpr (XA >= 8) *(int2*)(p + j) = make_int2(q[j], q[(j + 1) % NS]);
                else { p[j] = q[j]; p[j + 1] = q[(j + 1) % NS]; }
            }
            return;
        }
#pragma unroll
        for (int j = 0; j < NS; j += 2) {
            if (XA >= 8 && j + 1 < n) *(int2*)(p + j) = make_int2(q[j], q[(j + 1) % NS]);
            else { if (j < n) p[j] = q[j]; if (j + 1 < n) p[j + 1] = q[(j + 1) % NS]; }
        }
    }
    // nb valid bytes of the packed words wv at xrow (packed pixel rows; ES = bytes per sample)
    template <int XA, int NWO, int ES>
    static __device__ __forceinline__ void store_bytes_ua(unsigned char* xrow, const unsigned (&wv)[NWO], int nb) {
        if (nb >= 4 * NWO) {   // every lane of a strip that does not touch the window's right edge: no masks
#pragma unroll
            for (int k = 0; k < NWO; k++) {
                if constexpr (XA >= 4) *(unsigned*)(xrow + 4 * k) = wv[k];
                else if constexpr (ES == 2) {
                    *(unsigned short*)(xrow + 4 * k) = (unsigned short)(wv[k] & 0xFFFFu);
                    *(unsigned short*)(xrow + 4 * k + 2) = (unsigned short)(wv[k] >> 16);
                } else {
#pragma unroll
                    for (int b = 0; b < 4; b++) xrow[4 * k + b] = (unsigned char)((wv[k] >> (8 * b)) & 0xFFu);
                }
            }
            return;
        }
#pragma unroll
        for (int k = 0; k < NWO; k++) {
            const int left = nb - 4 * k;  // valid bytes of this word
            if (XA >= 4 && left >= 4) { *(unsigned*)(xrow + 4 * k) = wv[k]; continue; }
            if constexpr (ES == 2) {
                if (left >= 2) *(unsigned short*)(xrow + 4 * k) = (unsigned short)(wv[k] & 0xFFFFu);
                if (left >= 4) *(unsigned short*)(xrow + 4 * k + 2) = (unsigned short)(wv[k] >> 16);
            } else {
#pragma unroll
                for (int b = 0; b < 4; b++)
                    if (b < left) xrow[4 * k + b] = (unsigned char)((wv[k] >> (8 * b)) & 0xFFu);
            }
        }
    }

    // final stage of level 1: inverse MCT -> (+DC, optional planes) -> clamp -> pack -> one vector store per row
    template <int XA = 16, int PA = 16>
    static __device__ __forceinline__ void store_final(const RingSeg& S, const RawFmt& raw, const RawPack& rp, unsigned char* xrow, int* prow,
                                                       int (&iv)[NC][NS], int nvs = NS) {
        if constexpr (NC == 3 && MCT != MCTK_NONE) {
#pragma unroll
            for (int s = 0; s < NS; s++) {
                int r, g, b;
                mct_inverse<MCT>(iv[0][s], iv[1][s], iv[2][s], r, g, b);
                iv[0][s] = r; iv[1][s] = g; iv[2][s] = b;
            }
        }
        if constexpr (UA) {
            if (prow) {
                int t[NS];
#pragma unroll
                for (int s = 0; s < NS; s++) t[s] = iv[0][s] + raw.dc;
                store_ints_ua<PA>(prow, t, nvs);
            }
        } else
        if (prow) {
#pragma unroll
            for (int c = 0; c < NC; c++) {
                int* p = prow + c * S.planes_comp_stride;
                if constexpr (NS == 8) {
                    *(int4*)p = make_int4(iv[c][0] + raw.dc, iv[c][1] + raw.dc, iv[c][2] + raw.dc, iv[c][3] + raw.dc);
                    *(int4*)(p + 4) = make_int4(iv[c][NS - 4] + raw.dc, iv[c][NS - 3] + raw.dc, iv[c][NS - 2] + raw.dc, iv[c][NS - 1] + raw.dc);
                } else {
                    *(int4*)p = make_int4(iv[c][0] + raw.dc, iv[c][1] + raw.dc, iv[c][2 % NS] + raw.dc, iv[c][3 % NS] + raw.dc);
                }
            }
        }
        constexpr int ES = (OUT == IN_U8) ? 1 : 2;
        constexpr int NWO = NS * NC * ES / 4;  // 32-bit words per lane per row
        unsigned wv[NWO];
        // element e of the lane's interleaved span is component e % NC of sample e / NC
#pragma unroll
        for (int k = 0; k < NWO; k++) {
            if constexpr (OUT == IN_U8) {
                const unsigned t0 = pack_clamp2(iv[(4 * k) % NC][(4 * k) / NC], iv[(4 * k + 1) % NC][(4 * k + 1) / NC], rp);
                const unsigned t1 = pack_clamp2(iv[(4 * k + 2) % NC][(4 * k + 2) / NC], iv[(4 * k + 3) % NC][(4 * k + 3) / NC], rp);
                wv[k] = __byte_perm(t0, t1, 0x6420);
            } else {
                wv[k] = pack_clamp2(iv[(2 * k) % NC][(2 * k) / NC], iv[(2 * k + 1) % NC][(2 * k + 1) / NC], rp);
            }
        }
        if constexpr (UA) { store_bytes_ua<XA, NWO, ES>(xrow, wv, nvs * ES); return; }
        if constexpr (NWO % 4 == 0) {
#pragma unroll
            for (int k = 0; k < NWO / 4; k++) *((uint4*)xrow + k) = make_uint4(wv[4 * k], wv[4 * k + 1], wv[(4 * k + 2) % NWO], wv[(4 * k + 3) % NWO]);
        } else if constexpr (NWO % 2 == 0) {
#pragma unroll
            for (int k = 0; k < NWO / 2; k++) *((uint2*)xrow + k) = make_uint2(wv[2 * k], wv[(2 * k + 1) % NWO]);
        } else {
#pragma unroll
            for (int k = 0; k < NWO; k++) *((unsigned*)xrow + k) = wv[k];
        }
    }

    // Fast form of "round half-even, then the float64 inverse ICT with math.Round" (t2/tile_decoder.go:913, ict.go:16-21) for one
    // row of 8-bit pixels, valid while M = max(|y|, |cb|, |cr|) < 512 (y, cb, cr: the rounded samples, small integers).
    //   R = Round(y + 1.402 cr), B = Round(y + 1.772 cb).  The real values have fractions that are multiples of 1/500 and 1/250:
    //     either an exact tie (cr = 250 mod 500, cb = 125 mod 250) or at least 0.002 / 0.004 away from one.  The constants are
    //     split so that the float32 result is exact AT the ties: 1.402 = 1.375 + 0.027, 1.772 = 1.75 + 0.022;
    //     a = fma(1.375, cr, y) is exact (three fraction bits), t = fma(0.027f, cr, a) has one rounding and an error of
    //     |cr| 2^-30 before it, far below half a float32 grid step at |t| < 2^11, so a real tie k + 1/2 comes out as exactly
    //     k + 1/2.  Go's float64 product lands on the tie too (|cr| 2^-53 relative error of the constant against a 2^-44
    //     grid; tests/test_ict_fast.py checks every value), so math.Round rounds it half AWAY from zero; here t' = t (1 + 2^-20)
    //     pushes a tie eight grid steps away from zero before rint, and moves a non-tie by < |t| 2^-20 + 2^-14 < 0.0015:
    //     never across the tie that is >= 0.002 away.  No guard is needed for R and B.
    //   G = Round(y - 0.34413 cb - 0.71414 cr): fractions are multiples of 1e-5, float32 cannot separate them, so G keeps a
    //     distance-to-tie guard: t = fma(-0.71414f, cr, fma(-0.34413f, cb, y)) differs from Go's float64 value by less than
    //     4.2 M 2^-24 (constant errors M 2^-26 + M 2^-25, two roundings of <= 1.35 M 2^-24 and 2.06 M 2^-24); a lane whose t comes
    //     closer than E = M 2^-21 to a half-integer in any sample of the row returns false and takes the float64 path
    //     (2 E of the samples: 0.02 % at M = 200).
    // rint(t) is t + 1.5 * 2^23 - 1.5 * 2^23 (exact for |t| < 2^22, round half-even); the biased sum's low mantissa bits are the
    // integer itself, which pack_clamp2_magic picks up without an F2I.
    template <bool PRE = false>  // PRE: the samples arrive rounded already (inv3w_kernel's exchange ring)
    static __device__ __forceinline__ bool ict_fast_row(const RawPack& rp, unsigned char* xrow, const float2 (&x)[NC][NP]) {
        static_assert(NC == 3 && OUT == IN_U8, "8-bit interleaved RGB");
        const float2 MG = splat2(12582912.0f), NMG = splat2(-12582912.0f), NEG1 = splat2(-1.0f), BIAS = splat2(9.5367431640625e-7f);
        const float2 MGD = splat2(rp.magic_dc), NMGD = splat2(-rp.magic_dc);  // output bias with the DC shift folded in (exact)
        const float2 cRh = splat2(1.375f), cRl = splat2(0.027f), cBh = splat2(1.75f), cBl = splat2(0.022f);
        const float2 cG1 = splat2(-0.34413f), cG2 = splat2(-0.71414f);
        unsigned u[3][NS];  // biased bit patterns of R, G, B per sample
        float mag = 0.f, dist = 0.f;
#pragma unroll
        for (int j = 0; j < NP; j++) {
            const float2 y = PRE ? x[0][j] : add2(add2(x[0][j], MG), NMG), cb = PRE ? x[1][j] : add2(add2(x[1][j], MG), NMG),
                         cr = PRE ? x[2][j] : add2(add2(x[2][j], MG), NMG);
            mag = fmaxf(mag, fmaxf(fabsf(y.x), fabsf(y.y)));
            mag = fmaxf(mag, fmaxf(fabsf(cb.x), fabsf(cb.y)));
            mag = fmaxf(mag, fmaxf(fabsf(cr.x), fabsf(cr.y)));
            const float2 tr = fma2(cRl, cr, fma2(cRh, cr, y)), tb = fma2(cBl, cb, fma2(cBh, cb, y));
            const float2 tg = fma2(cG2, cr, fma2(cG1, cb, y));
            const float2 br = add2(fma2(tr, BIAS, tr), MGD), bb = add2(fma2(tb, BIAS, tb), MGD), bg = add2(tg, MGD);
            const float2 d = fma2(add2(bg, NMGD), NEG1, tg);  // tg - rint(tg), exact (the integer dc moves no fraction)
            dist = fmaxf(dist, fmaxf(fabsf(d.x), fabsf(d.y)));
            u[0][2 * j] = __float_as_uint(br.x); u[0][2 * j + 1] = __float_as_uint(br.y);
            u[1][2 * j] = __float_as_uint(bg.x); u[1][2 * j + 1] = __float_as_uint(bg.y);
            u[2][2 * j] = __float_as_uint(bb.x); u[2][2 * j + 1] = __float_as_uint(bb.y);
        }
        if (!(mag < 512.0f) || !(dist < 0.5f - mag * 4.76837158203125e-7f)) return false;
        constexpr int NWO = NS * 3 / 4;
        unsigned wv[NWO];
#pragma unroll
        for (int k = 0; k < NWO; k++) {
            const unsigned t0 = pack_clamp2_magic_dc(u[(4 * k) % 3][(4 * k) / 3], u[(4 * k + 1) % 3][(4 * k + 1) / 3], rp);
            const unsigned t1 = pack_clamp2_magic_dc(u[(4 * k + 2) % 3][(4 * k + 2) / 3], u[(4 * k + 3) % 3][(4 * k + 3) / 3], rp);
            wv[k] = __byte_perm(t0, t1, 0x6420);
        }
        if constexpr (NWO % 4 == 0) {
#pragma unroll
            for (int k = 0; k < NWO / 4; k++) *((uint4*)xrow + k) = make_uint4(wv[4 * k], wv[4 * k + 1], wv[(4 * k + 2) % NWO], wv[(4 * k + 3) % NWO]);
        } else if constexpr (NWO % 2 == 0) {
#pragma unroll
            for (int k = 0; k < NWO / 2; k++) *((uint2*)xrow + k) = make_uint2(wv[2 * k], wv[(2 * k + 1) % NWO]);
        } else {
#pragma unroll
            for (int k = 0; k < NWO; k++) *((unsigned*)xrow + k) = wv[k];
        }
        return true;
    }

    // planar destination (intermediate LL, wavelet-package API, generic path): working-type bits, or rounded samples
    static __device__ __forceinline__ void store_planar(int* p, const int (&q)[NS]) {
        if constexpr (NS == 8) {
            *(int4*)p = make_int4(q[0], q[1], q[2], q[3]);
            *(int4*)(p + 4) = make_int4(q[NS - 4], q[NS - 3], q[NS - 2], q[NS - 1]);
        } else if constexpr (NS == 4) {
            *(int4*)p = make_int4(q[0], q[1], q[2 % NS], q[3 % NS]);
        } else {
            *(int2*)p = make_int2(q[0], q[1 % NS]);
        }
    }

    // The job loop is instantiated per destination alignment of the general-alignment variant and chosen once per job
    // (XA: pixel / LL rows, PA: GetImageData plane rows; in bytes, 1 = element-wise stores).
    static __device__ __forceinline__ void run(const RingSeg& S, const RawFmt& raw, int item, int chunk, int strip, RingWarp& rw,
                                               int lane, XchPort* xp = nullptr, int comp = 0) {
        if constexpr (!UA) run_t<16, 16>(S, raw, item, chunk, strip, rw, lane, xp, comp);
        else {
            const bool pa = S.st_cls[0] >= 2;  // plane rows allow 64-bit stores
            if (S.x_align >= 8) { if (pa) run_t<8, 8>(S, raw, item, chunk, strip, rw, lane); else run_t<8, 4>(S, raw, item, chunk, strip, rw, lane); }
            else if (S.x_align >= 4) { if (pa) run_t<4, 8>(S, raw, item, chunk, strip, rw, lane); else run_t<4, 4>(S, raw, item, chunk, strip, rw, lane); }
            else { if (pa) run_t<1, 8>(S, raw, item, chunk, strip, rw, lane); else run_t<1, 4>(S, raw, item, chunk, strip, rw, lane); }
        }
    }

    template <int XA, int PA>
    static __device__ __forceinline__ void run_t(const RingSeg& S, const RawFmt& raw, int item, int chunk, int strip, RingWarp& rw,
                                               int lane, XchPort* xp = nullptr, int comp = 0) {
        static_assert(!XCH || (WT == 97 && NC == 1 && NP == 4 && !FINAL && !UA), "exchange mode: one 9/7 component per job");
        const int w = S.w, h = S.h, py = S.py;
        const int bw = S.lw;  // == w - lw on this path (w is a multiple of 2 NP)
        const int kxs = strip * S.strip_pairs;
        const int kxe = min(kxs + S.strip_pairs, S.Kx);
        const int nl = (kxe - kxs + NP - 1) / NP + 2 * HLN;
        const int kx0 = kxs - HLN * NP + lane * NP;
        const int k0w = kxs - (HLN + SHL) * NP;   // band index of the first staged element
        const int m = (k0w * 4) & 15;
        const int vb = k0w * 4 - m;               // virtual (16 B aligned) byte position of the slot start inside a band row
        const int c0 = max(vb, 0);
        const int c1 = min(vb + ((m + (nl + 2 * SHL) * LB + 15) & ~15), bw * 4);
        const unsigned copy_bytes = (unsigned)(c1 - c0);
        const int dst_off = c0 - vb;
        const int lane_off = m + (lane + SHL) * LB;
        // (UA: the lanes right of a strip may reach past the high-pass band although another, narrower strip follows)
        const bool fix_l = kxs == 0, fix_r = UA ? (kxe + NP > w - bw) : (kxe == S.Kx);
        const bool fix = fix_l || fix_r;
        const bool st = lane >= HLN && kx0 < kxe && (UA || kx0 + NP <= bw);
        const int nvs = w - 2 * kx0;                 // UA: samples of this lane inside the window (<= 0: none, >= NS: all)
        (void)nvs;
        const float one = rw.one;

        const int ky0 = chunk * S.chunk_pairs;
        const int ky1 = min(ky0 + S.chunk_pairs, S.Ky);
        const int t_begin = ky0 - LAG;
        const int n_it = ky1 - ky0 + 2 * LAG;

        // band row 0 of this item, at the copy start (bytes)
        const unsigned char* b_ll = (const unsigned char*)((const int*)S.ll.base + S.ll.off[item] + (long long)S.ll.y_off * S.ll.row_stride + S.ll.x_off) + c0;
        const unsigned char* b_hl = (const unsigned char*)((const int*)S.hl.base + S.hl.off[item] + (long long)S.hl.y_off * S.hl.row_stride + S.hl.x_off) + c0;
        const unsigned char* b_lh = (const unsigned char*)((const int*)S.lh_.base + S.lh_.off[item] + (long long)S.lh_.y_off * S.lh_.row_stride + S.lh_.x_off) + c0;
        const unsigned char* b_hh = (const unsigned char*)((const int*)S.hh.base + S.hh.off[item] + (long long)S.hh.y_off * S.hh.row_stride + S.hh.x_off) + c0;
        const long long rp_ll = (long long)S.ll.row_stride * 4, rp_b = (long long)S.hl.row_stride * 4;
        const long long cp_ll = S.ll.comp_stride * 4, cp_b = S.hl.comp_stride * 4;
        if constexpr (XCH) { b_ll += comp * cp_ll; b_hl += comp * cp_b; b_lh += comp * cp_b; b_hh += comp * cp_b; }

        const int n_st = (n_it + RPS - 1) / RPS;
        int pj = 0, pslot = 0;
        const smem_t dst_s = rw.ring + dst_off;
        // stages s_lo <= s < s_hi hold only in-range row pairs t (0 <= 2t - py, 2t + 1 - py < h): their band rows are
        // yl = t - py, yh = t, reached with running pointers; the others mirror the interleaved row numbers
        const int t_min = (py + 1) >> 1, t_max = (h - 2 + py) >> 1;  // first / last pair with both rows inside the window
        const int s_lo = t_begin < t_min ? (t_min - t_begin + RPS - 1) / RPS : 0;
        const int s_hi = t_max >= t_begin ? (t_max - t_begin + 1) / RPS : 0;
        const unsigned char* q_ll = b_ll + (long long)(t_begin - py) * rp_ll;  // band rows of pair t_begin
        const unsigned char* q_hl = b_hl + (long long)(t_begin - py) * rp_b;
        const unsigned char* q_lh = b_lh + (long long)t_begin * rp_b;
        const unsigned char* q_hh = b_hh + (long long)t_begin * rp_b;
        // one elected lane stages the band rows of the RPS row pairs of stage pj
        auto issue = [&]() {
            if (elect_one()) {
                const smem_t bar = rw.bars + 8 * pslot;
                const smem_t dst = dst_s + pslot * STAGEB;
                if constexpr (UA) {
                    // every band row starts at its own byte: copy from the 16-byte boundary below, remember the phase in the slot
                    const unsigned char* rp[RPS * 4];
                    unsigned ph[RPS * 4], nb[RPS * 4], total = 0;
                    const bool inr = pj >= s_lo && pj < s_hi;
#pragma unroll
                    for (int k = 0; k < RPS; k++) {
                        if (inr) {
                            rp[4 * k + 0] = q_ll + k * rp_ll; rp[4 * k + 1] = q_hl + k * rp_b; rp[4 * k + 2] = q_lh + k * rp_b; rp[4 * k + 3] = q_hh + k * rp_b;
                        } else {
                            const int t = t_begin + RPS * pj + k;
                            const int pl = mirror_fast(2 * t - py, h), phh = mirror_fast(2 * t + 1 - py, h);
                            const int yl = (pl - py) >> 1, yh = (phh - (1 - py)) >> 1;
                            rp[4 * k + 0] = b_ll + yl * rp_ll; rp[4 * k + 1] = b_hl + yl * rp_b; rp[4 * k + 2] = b_lh + yh * rp_b; rp[4 * k + 3] = b_hh + yh * rp_b;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RPS * 4; r++) {
                        ph[r] = (unsigned)((size_t)rp[r] & 15u);
                        nb[r] = (ph[r] + copy_bytes + 15u) & ~15u;
                        total += nb[r];
                        sts32(rw.ring + pslot * STAGEB + r * ROWB + ROWB - 4, ph[r]);
                    }
                    mbar_expect_tx(bar, total);
#pragma unroll
                    for (int r = 0; r < RPS * 4; r++) bulk_g2s(dst + r * ROWB, rp[r] - ph[r], nb[r], bar);
                } else {
                mbar_expect_tx(bar, RPS * NROWS * copy_bytes);
                if (pj >= s_lo && pj < s_hi) {
#pragma unroll
                    for (int k = 0; k < RPS; k++)
#pragma unroll
                        for (int c = 0; c < NC; c++) {
                            bulk_g2s(dst + k * PAIRB + (4 * c + 0) * ROWB, q_ll + c * cp_ll + k * rp_ll, copy_bytes, bar);
                            bulk_g2s(dst + k * PAIRB + (4 * c + 1) * ROWB, q_hl + c * cp_b + k * rp_b, copy_bytes, bar);
                            bulk_g2s(dst + k * PAIRB + (4 * c + 2) * ROWB, q_lh + c * cp_b + k * rp_b, copy_bytes, bar);
                            bulk_g2s(dst + k * PAIRB + (4 * c + 3) * ROWB, q_hh + c * cp_b + k * rp_b, copy_bytes, bar);
                        }
                } else {
#pragma unroll
                    for (int k = 0; k < RPS; k++) {
                        // band rows of the pair: low-type interleaved row 2t - py -> yl, high-type row 2t + 1 - py -> yh (mirrored)
                        const int t = t_begin + RPS * pj + k;
                        const int pl = mirror_fast(2 * t - py, h), ph = mirror_fast(2 * t + 1 - py, h);
                        const int yl = (pl - py) >> 1, yh = (ph - (1 - py)) >> 1;
#pragma unroll
                        for (int c = 0; c < NC; c++) {
                            bulk_g2s(dst + k * PAIRB + (4 * c + 0) * ROWB, b_ll + c * cp_ll + yl * rp_ll, copy_bytes, bar);
                            bulk_g2s(dst + k * PAIRB + (4 * c + 1) * ROWB, b_hl + c * cp_b + yl * rp_b, copy_bytes, bar);
                            bulk_g2s(dst + k * PAIRB + (4 * c + 2) * ROWB, b_lh + c * cp_b + yh * rp_b, copy_bytes, bar);
                            bulk_g2s(dst + k * PAIRB + (4 * c + 3) * ROWB, b_hh + c * cp_b + yh * rp_b, copy_bytes, bar);
                        }
                    }
                }
                }
            }
            pj++;
            q_ll += RPS * rp_ll; q_hl += RPS * rp_b; q_lh += RPS * rp_b; q_hh += RPS * rp_b;
            pslot = (pslot + 1 == D) ? 0 : pslot + 1;
        };
        __syncwarp();
#pragma unroll 1
        for (int j = 0; j < D - 1 && j < n_st; j++) issue();

        // destination rows
        constexpr int XES = FINAL ? ((OUT == IN_U8) ? 1 : 2) : 4;
        constexpr int XPB = XES * (FINAL ? NC : 1);
        unsigned char* xlane = (unsigned char*)S.x_base + S.x_off[item] * XES + (long long)(2 * kx0) * XPB;
        const long long xpitch = S.x_row_bytes;
        int* planes = (FINAL && S.planes_out) ? S.planes_out + S.planes_off[item] + 2 * kx0 : nullptr;
        const int planes_rs = S.planes_row_stride;
        const RawPack rp = make_raw_pack(raw);
        // row 2 ky0 - py of the first storing iteration; both advance two rows per iteration
        xlane += (long long)(2 * ky0 - py) * xpitch;
        if (planes) planes += (long long)(2 * ky0 - py) * planes_rs;
        const int x_mode = S.x_mode;

        int cslot = 0;
        smem_t stage_base = rw.ring;
        auto next_stage = [&]() {
            __syncwarp();
            if (pj < n_st) issue();
            mbar_wait(rw.bars + 8 * cslot, (rw.phase >> cslot) & 1u);
            rw.phase ^= 1u << cslot;
            stage_base = rw.ring + cslot * STAGEB;
            cslot = (cslot + 1 == D) ? 0 : cslot + 1;
            if (fix) fix_halo(stage_base, lane, fix_l, fix_r, w, bw, vb);
        };
        if constexpr (WT == 97) {
            const float2 sclE = S.rcpE, sclO = S.rcpO;
            const bool raw_ll = S.ll.mode == DQ_RAW, raw_hl = S.hl.mode == DQ_RAW, raw_lh = S.lh_.mode == DQ_RAW, raw_hh = S.hh.mode == DQ_RAW;
            const bool std_modes = raw_ll && !raw_hl && !raw_lh && !raw_hh;
            const float2 nD = splat2(-J2K_DELTA), nG = splat2(-J2K_GAMMA), nB = splat2(-J2K_BETA), nA = splat2(-J2K_ALPHA);
            struct VState { float2 dp[NC][NP], s1p[NC][NP], d1p[NC][NP], s2p[NC][NP]; };
            VState sa, sb;
#pragma unroll
            for (int c = 0; c < NC; c++)
#pragma unroll
                for (int j = 0; j < NP; j++) { sa.dp[c][j] = sa.s1p[c][j] = sa.d1p[c][j] = sa.s2p[c][j] = make_float2(0.f, 0.f); }
            auto body = [&](int it, const int half, const VState& in, VState& out) {
                if (RPS == 1 || half == 0) next_stage();
                const smem_t stage = stage_base + (RPS == 1 ? 0 : half) * PAIRB;

                float2 xe[NC][NP], xo[NC][NP];  // finished rows of pair t - LAG as column pairs
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    // dequantize: S[j] = (LL, LH)[j], Dd[j] = (HL, HH)[j] as (low-type row, high-type row) pairs
                    float2 Sx[NP + 1], Dx[NP + 1];  // Dx[0] = previous lane's last high, Sx[NP] = next lane's first low
                    int q0[NP], q1[NP], q2[NP], q3[NP];
                    if constexpr (UA) {
                        const smem_t r0 = stage + (4 * c) * ROWB;
                        fetch_ua(r0, lane_off + row_phase(r0), q0);
                        fetch_ua(r0 + ROWB, lane_off + row_phase(r0 + ROWB), q1);
                        fetch_ua(r0 + 2 * ROWB, lane_off + row_phase(r0 + 2 * ROWB), q2);
                        fetch_ua(r0 + 3 * ROWB, lane_off + row_phase(r0 + 3 * ROWB), q3);
                    } else {
                    fetch(stage + (4 * c + 0) * ROWB + lane_off, q0);  // every lane, also past the strip (see FwdRing::load_scalar)
                    fetch(stage + (4 * c + 1) * ROWB + lane_off, q1);
                    fetch(stage + (4 * c + 2) * ROWB + lane_off, q2);
                    fetch(stage + (4 * c + 3) * ROWB + lane_off, q3);
                    }
                    // float32(q) * float32(scale) (t2/tile_decoder.go:970-987); DQ_CVT is scale == 1 (exact); then the
                    // horizontal synthesis scaling: low * K, high * two_invK (dwt97.go:207-212)
                    if (std_modes) {  // every level but the coarsest: LL is the float32 plane of the level below
#pragma unroll
                        for (int j = 0; j < NP; j++) {
                            const float2 fs = make_float2(0.f, (float)q2[j]), fd = make_float2((float)q1[j], (float)q3[j]);
                            float2 vs = mul2(fs, sclE);
                            const float2 vd = mul2(fd, sclO);
                            vs.x = __int_as_float(q0[j]);
                            Sx[j] = mul2(vs, splat2(J2K_K));
                            Dx[j + 1] = mul2(vd, splat2(J2K_TWOINVK));
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < NP; j++) {
                            const float2 fs = make_float2((float)q0[j], (float)q2[j]), fd = make_float2((float)q1[j], (float)q3[j]);
                            float2 vs = mul2(fs, sclE), vd = mul2(fd, sclO);
                            if (raw_ll) vs.x = __int_as_float(q0[j]);
                            if (raw_lh) vs.y = __int_as_float(q2[j]);
                            if (raw_hl) vd.x = __int_as_float(q1[j]);
                            if (raw_hh) vd.y = __int_as_float(q3[j]);
                            Sx[j] = mul2(vs, splat2(J2K_K));
                            Dx[j + 1] = mul2(vd, splat2(J2K_TWOINVK));
                        }
                    }
                    float2 Tq[NP];
                    Dx[0] = shfl_up2(Dx[NP]);
#pragma unroll
                    for (int j = 0; j < NP; j++) Tq[j] = add2(Dx[j], Dx[j + 1]);
#pragma unroll
                    for (int j = 0; j < NP; j++) Tq[j] = mul2(Tq[j], nD);
#pragma unroll
                    for (int j = 0; j < NP; j++) Sx[j] = addp2(Tq[j], Sx[j], one);
                    Sx[NP] = shfl_down2(Sx[0]);
#pragma unroll
                    for (int j = 0; j < NP; j++) Tq[j] = add2(Sx[j], Sx[j + 1]);
#pragma unroll
                    for (int j = 0; j < NP; j++) Tq[j] = mul2(Tq[j], nG);
#pragma unroll
                    for (int j = 0; j < NP; j++) Dx[j + 1] = addp2(Tq[j], Dx[j + 1], one);
                    Dx[0] = shfl_up2(Dx[NP]);
#pragma unroll
                    for (int j = 0; j < NP; j++) Tq[j] = add2(Dx[j], Dx[j + 1]);
#pragma unroll
                    for (int j = 0; j < NP; j++) Tq[j] = mul2(Tq[j], nB);
#pragma unroll
                    for (int j = 0; j < NP; j++) Sx[j] = addp2(Tq[j], Sx[j], one);
                    Sx[NP] = shfl_down2(Sx[0]);
#pragma unroll
                    for (int j = 0; j < NP; j++) Tq[j] = add2(Sx[j], Sx[j + 1]);
#pragma unroll
                    for (int j = 0; j < NP; j++) Tq[j] = mul2(Tq[j], nA);
#pragma unroll
                    for (int j = 0; j < NP; j++) Dx[j + 1] = addp2(Tq[j], Dx[j + 1], one);
                    // vertical synthesis on column pairs (2j, 2j+1) = (Sx[j], Dx[j+1]); .x = low-type row e, .y = high-type row o
#pragma unroll
                    for (int j = 0; j < NP; j++) {
                        const float2 sv = make_float2(__fmul_rn(Sx[j].x, J2K_K), __fmul_rn(Dx[j + 1].x, J2K_K));
                        const float2 dv = make_float2(__fmul_rn(Sx[j].y, J2K_TWOINVK), __fmul_rn(Dx[j + 1].y, J2K_TWOINVK));
                        const float2 s1 = lift97x2(sv, in.dp[c][j], dv, nD, one);               // s'[t]
                        const float2 d1 = lift97x2(in.dp[c][j], in.s1p[c][j], s1, nG, one);     // d'[t-1]
                        const float2 s2 = lift97x2(in.s1p[c][j], in.d1p[c][j], d1, nB, one);    // s''[t-1]
                        const float2 d2 = lift97x2(in.d1p[c][j], in.s2p[c][j], s2, nA, one);    // d''[t-2]
                        xe[c][j] = in.s2p[c][j];                                                 // s''[t-2]
                        xo[c][j] = d2;
                        out.dp[c][j] = dv; out.s1p[c][j] = s1; out.d1p[c][j] = d1; out.s2p[c][j] = s2;
                    }
                }
                if (it < 2 * LAG) return;
                if constexpr (XCH) { xp->put(comp, lane, xe[0], xo[0]); return; }  // every lane, every emitting iteration: the consumer masks
                const int ky = ky0 + it - 2 * LAG;
                const int re = 2 * ky - py, ro = re + 1;
                unsigned char* const xrow = xlane;
                int* const prow = planes;
                xlane += 2 * xpitch;
                if (planes) planes += 2 * planes_rs;
                if (!st) return;
                if constexpr (FINAL) {
                    constexpr bool ICTF = J2K_ICT_FAST && NC == 3 && MCT == MCTK_ICT && OUT == IN_U8;
                    int iv[NC][NS];
                    if (re >= 0 && re < h) {
                        bool done = false;
                        if constexpr (ICTF) { if (!prow) done = ict_fast_row(rp, xrow, xe); }
                        if (!done) {
#pragma unroll
                            for (int c = 0; c < NC; c++)
#pragma unroll
                                for (int j = 0; j < NP; j++) { iv[c][2 * j] = __float2int_rn(xe[c][j].x); iv[c][2 * j + 1] = __float2int_rn(xe[c][j].y); }
                            store_final<XA, PA>(S, raw, rp, xrow, prow, iv, nvs);
                        }
                    }
                    if (ro < h) {
                        bool done = false;
                        if constexpr (ICTF) { if (!prow) done = ict_fast_row(rp, xrow + xpitch, xo); }
                        if (!done) {
#pragma unroll
                            for (int c = 0; c < NC; c++)
#pragma unroll
                                for (int j = 0; j < NP; j++) { iv[c][2 * j] = __float2int_rn(xo[c][j].x); iv[c][2 * j + 1] = __float2int_rn(xo[c][j].y); }
                            store_final<XA, PA>(S, raw, rp, xrow + xpitch, prow ? prow + planes_rs : nullptr, iv, nvs);
                        }
                    }
                } else {
                    int q[NS];
                    if (re >= 0 && re < h) {
#pragma unroll
                        for (int j = 0; j < NP; j++) {
                            q[2 * j] = x_mode == 1 ? __float2int_rn(xe[0][j].x) : __float_as_int(xe[0][j].x);
                            q[2 * j + 1] = x_mode == 1 ? __float2int_rn(xe[0][j].y) : __float_as_int(xe[0][j].y);
                        }
                        if constexpr (UA) store_ints_ua<XA>((int*)xrow, q, nvs); else
                        store_planar((int*)xrow, q);
                    }
                    if (ro < h) {
#pragma unroll
                        for (int j = 0; j < NP; j++) {
                            q[2 * j] = x_mode == 1 ? __float2int_rn(xo[0][j].x) : __float_as_int(xo[0][j].x);
                            q[2 * j + 1] = x_mode == 1 ? __float2int_rn(xo[0][j].y) : __float_as_int(xo[0][j].y);
                        }
                        if constexpr (UA) store_ints_ua<XA>((int*)(xrow + xpitch), q, nvs); else
                        store_planar((int*)(xrow + xpitch), q);
                    }
                }
            };
            if constexpr (NC == 3 && WT == 97 && J2K_RGB_SINGLE_BODY) {
                // three 9/7 components: ONE copy of the body and a state move per iteration -- the two-copy form of these
                // variants is ~60 KB of code and misses the instruction cache (stall_no_inst 11 % -> +15 % on C3 / C5)
#pragma unroll 1
                for (int it = 0; it < n_it; it++) {
                    body(it, RPS == 1 ? 0 : (it & 1), sa, sb);
                    sa = sb;
                }
            } else
#pragma unroll 1
            for (int it = 0; it < n_it; it += 2) {
                body(it, 0, sa, sb);
                if (it + 1 >= n_it) break;
                body(it + 1, 1, sb, sa);
            }
        } else {
            const bool halve_ll = S.ll.mode == DQ_HALVE, halve_hl = S.hl.mode == DQ_HALVE, halve_lh = S.lh_.mode == DQ_HALVE,
                       halve_hh = S.hh.mode == DQ_HALVE;
            const bool any_halve = halve_ll || halve_hl || halve_lh || halve_hh;
            struct VState { int dp[NC][NS], s1p[NC][NS]; };
            VState sa, sb;
#pragma unroll
            for (int c = 0; c < NC; c++)
#pragma unroll
                for (int s = 0; s < NS; s++) { sa.dp[c][s] = 0; sa.s1p[c][s] = 0; }
            // one iteration: consumes the vertical window state `in`, leaves the advanced state in `out` (two per trip,
            // ping-ponging, so the window never moves between registers)
            auto body = [&](int it, const int half, const VState& in, VState& out) {
                if (RPS == 1 || half == 0) next_stage();
                const smem_t stage = stage_base + (RPS == 1 ? 0 : half) * PAIRB;

                int xe[NC][NS], xo[NC][NS];
#pragma unroll
                for (int c = 0; c < NC; c++) {
                    int q0[NP], q1[NP], q2[NP], q3[NP];
                    if constexpr (UA) {
                        const smem_t r0 = stage + (4 * c) * ROWB;
                        fetch_ua(r0, lane_off + row_phase(r0), q0);
                        fetch_ua(r0 + ROWB, lane_off + row_phase(r0 + ROWB), q1);
                        fetch_ua(r0 + 2 * ROWB, lane_off + row_phase(r0 + 2 * ROWB), q2);
                        fetch_ua(r0 + 3 * ROWB, lane_off + row_phase(r0 + 3 * ROWB), q3);
                    } else {
                    fetch(stage + (4 * c + 0) * ROWB + lane_off, q0);  // every lane, also past the strip (see FwdRing::load_scalar)
                    fetch(stage + (4 * c + 1) * ROWB + lane_off, q1);
                    fetch(stage + (4 * c + 2) * ROWB + lane_off, q2);
                    fetch(stage + (4 * c + 3) * ROWB + lane_off, q3);
                    }
                    // t2/tile_decoder.go:989-993: truncating /2 of the classic T1 output (fuse_t1_halve only: warp-uniform)
                    if (any_halve) {
#pragma unroll
                        for (int j = 0; j < NP; j++) {
                            if (halve_ll) q0[j] /= 2;
                            if (halve_hl) q1[j] /= 2;
                            if (halve_lh) q2[j] /= 2;
                            if (halve_hh) q3[j] /= 2;
                        }
                    }
                    // horizontal synthesis of the low-type row (q0 | q1) and the high-type row (q2 | q3) (dwt53.go:123-234)
                    int e[NS], o[NS];
                    // neighbours: the high sample left of the lane's span, the low and high samples right of it.  Halo-free
                    // strips read them from the staged rows (every lane: no shuffle, no lane predicate); otherwise they come
                    // from the adjacent lanes.
                    auto hsyn = [&](const int (&ql)[NP], const int (&qh)[NP], int row_l, bool halve_l, bool halve_h, int (&out)[NS]) {
                        int s[NP + 1], d[NP + 1];
#pragma unroll
                        for (int j = 0; j < NP; j++) { s[j] = ql[j]; d[j + 1] = qh[j]; }
                        if constexpr (HF) {
                            // (every lane reads its neighbours' samples from the staged row; shuffles for the inner lanes plus
                            // predicated loads for the strip's outer lanes were measured 9 % slower on C1 / C4)
                            const smem_t rl = stage + row_l * ROWB;
                            const smem_t pl = rl + lane_off + row_phase(rl), ph = rl + ROWB + lane_off + row_phase(rl + ROWB);
                            int dl = (int)lds32(ph - 4), sr = (int)lds32(pl + LB), dr = (int)lds32(ph + LB);
                            if (halve_h) { dl /= 2; dr /= 2; }
                            if (halve_l) sr /= 2;
                            d[0] = dl;
#pragma unroll
                            for (int j = 0; j < NP; j++) s[j] = s[j] - ((d[j] + d[j + 1] + 2) >> 2);
                            s[NP] = sr - ((d[NP] + dr + 2) >> 2);
                        } else {
                            d[0] = __shfl_up_sync(0xffffffffu, d[NP], 1);
#pragma unroll
                            for (int j = 0; j < NP; j++) s[j] = s[j] - ((d[j] + d[j + 1] + 2) >> 2);
                            s[NP] = __shfl_down_sync(0xffffffffu, s[0], 1);
                        }
#pragma unroll
                        for (int j = 0; j < NP; j++) d[j + 1] = d[j + 1] + ((s[j] + s[j + 1]) >> 1);
#pragma unroll
                        for (int j = 0; j < NP; j++) { out[2 * j] = s[j]; out[2 * j + 1] = d[j + 1]; }
                    };
                    hsyn(q0, q1, 4 * c + 0, halve_ll, halve_hl, e);
                    hsyn(q2, q3, 4 * c + 2, halve_lh, halve_hh, o);
                    // vertical synthesis (dwt53.go:318-354 columns after rows)
#pragma unroll
                    for (int s = 0; s < NS; s++) {
                        const int sv = e[s] - ((in.dp[c][s] + o[s] + 2) >> 2);        // s[t]
                        const int xodd = in.dp[c][s] + ((in.s1p[c][s] + sv) >> 1);    // x[2(t-1)+1]
                        xe[c][s] = in.s1p[c][s];
                        xo[c][s] = xodd;
                        out.dp[c][s] = o[s]; out.s1p[c][s] = sv;
                    }
                }
                if (it < 2 * LAG) return;
                const int ky = ky0 + it - 2 * LAG;
                const int re = 2 * ky - py, ro = re + 1;
                unsigned char* const xrow = xlane;
                int* const prow = planes;
                xlane += 2 * xpitch;
                if (planes) planes += 2 * planes_rs;
                if (!st) return;
                if constexpr (FINAL) {
                    if (re >= 0 && re < h) store_final<XA, PA>(S, raw, rp, xrow, prow, xe, nvs);
                    if (ro < h) store_final<XA, PA>(S, raw, rp, xrow + xpitch, prow ? prow + planes_rs : nullptr, xo, nvs);
                } else if constexpr (UA) {
                    if (re >= 0 && re < h) store_ints_ua<XA>((int*)xrow, xe[0], nvs);
                    if (ro < h) store_ints_ua<XA>((int*)(xrow + xpitch), xo[0], nvs);
                } else {
                    if (re >= 0 && re < h) store_planar((int*)xrow, xe[0]);
                    if (ro < h) store_planar((int*)(xrow + xpitch), xo[0]);
                }
            };
            if constexpr (NC == 3 && WT == 97 && J2K_RGB_SINGLE_BODY) {
                // three 9/7 components: ONE copy of the body and a state move per iteration -- the two-copy form of these
                // variants is ~60 KB of code and misses the instruction cache (stall_no_inst 11 % -> +15 % on C3 / C5)
#pragma unroll 1
                for (int it = 0; it < n_it; it++) {
                    body(it, RPS == 1 ? 0 : (it & 1), sa, sb);
                    sa = sb;
                }
            } else
#pragma unroll 1
            for (int it = 0; it < n_it; it += 2) {
                body(it, 0, sa, sb);
                if (it + 1 >= n_it) break;
                body(it + 1, 1, sb, sa);
            }
        }
    }
};

// WT: 53 / 97.  OUT1 / NC1 / MCT1: the variant of segments with first == 1 (level 1: packed pixels, or planar for the
// wavelet-package API / generic path); coarser levels always write planar working-type LL planes.
#ifndef J2K_INV_MINB_53
#define J2K_INV_MINB_53 4    // resident CTAs per SM targeted by the single-component 5/3 inverse (128 registers, 16 warps: +2 % on C1 / C4)
#endif
#ifndef J2K_INV3W_DEFAULT
#define J2K_INV3W_DEFAULT 1  // ICT + 9/7 8-bit RGB inverse: three-producer level 1 (inv3w_kernel) instead of one warp per pixel strip
#endif
#ifndef J2K_INV_RGB_NP
#define J2K_INV_RGB_NP 2     // sample pairs per lane of the 3-component level-1 inverse (4: 2 CTAs/SM, the window state of three components in 255 registers)
#endif
template <int WT, int NP1, int NC1, int OUT1, int MCT1, int UA = 0>
__global__ void __launch_bounds__(J2K_RING_WARPS * 32, (NC1 == 3 && NP1 == 4) ? 2 : ((WT == 53 && NC1 == 1 && !UA) ? J2K_INV_MINB_53 : J2K_RING_MINB)) inv_ring_kernel(const __grid_constant__ RingArgs A) {
    J2K_SMEM_DECL(smem);
    const int lane = threadIdx.x & 31;
    RingWarp rw;
    constexpr int RB = inv_ring_bytes<WT, NC1, UA>();
    ring_warp_init(smem, rw, lane, RB);
    rw.one = A.one;
    RingJob J;
    #ifndef J2K_INV_CTA_CLAIM
#define J2K_INV_CTA_CLAIM 0
#endif
    J2K_RING_SCHED_DECL(A, sched);
    while (ring_claim<(J2K_INV_CTA_CLAIM != 0)>(A, sched, lane, J)) {
        const RingSeg& S = A.seg[J.seg];
        ring_wait_dep(A, S, J.item, lane);
        if (S.first) InvRing<WT, NP1, NC1, OUT1, MCT1, RB, UA>::run(S, A.raw, J.item, J.chunk, J.strip, rw, lane);
        else InvRing<WT, 4, 1, (WT == 53 ? IN_I32 : IN_F32), MCTK_NONE, RB, UA>::run(S, A.raw, J.item, J.chunk, J.strip, rw, lane);
        ring_signal(A, S, J.item, lane);
    }
    ring_retire(A, lane);
}


// ------------------------------------------------------------------ three-producer level-1 inverse (ICT + 9/7, 8-bit RGB)
//
// The level-1 inverse of a three-component 9/7 frame needs all three components of a pixel at its very end (the inverse
// ICT), which forced the one-warp-per-pixel-strip job to carry three vertical window states (NP = 2: half the columns per
// lane, twice the per-lane overhead, 1.0 warp instruction per sample against 0.66 for a single component).  Here a CTA of
// four warps shares the job instead: warps 0..2 each run the SINGLE-component pipeline (InvRing<97, NP = 4> in exchange
// mode, own TMA ring) for Y, Cb, Cr of the same pixel strip and hand their finished, rounded rows to warp 3 through a
// DX-stage shared-memory ring (mbarrier full / empty per stage: the warps are decoupled by DX row pairs, no CTA barrier in
// the loop); warp 3 runs the inverse ICT (float32 fast path, float64 fallback per lane), +DC, clamp, pack and the pixel
// stores.  Coarser levels: warps 0..2 take ordinary single-component jobs, warp 3 idles.  Jobs are taken three at a time
// per CTA; a level-1 pixel job occupies three consecutive job numbers (one per component), every segment's job range is
// padded to a multiple of three.
#ifndef J2K_X3_RING_BYTES
#define J2K_X3_RING_BYTES 13056   // producer staging per warp: 3 stages of 2 x four 544 B band rows
#endif
#define J2K_X3_WARP_SMEM (J2K_X3_RING_BYTES + J2K_RING_MAXD * 8)
#define J2K_X3_CTA_SMEM (3 * J2K_X3_WARP_SMEM + J2K_X3_DX * J2K_X3_STAGEB + 2 * J2K_X3_DX * 8)

template <int OUT>
struct Inv3WConsumer {
    typedef InvRing<97, 4, 3, OUT, MCTK_ICT> C3;   // its pixel tail: ict_fast_row / store_final (inverse ICT, pack, stores)
    static constexpr int NP = 4, NS = 8;
    static __device__ __forceinline__ void run(const RingSeg& S, const RawFmt& raw, int item, int chunk, int strip, XchPort& xp, int lane) {
        const int h = S.h, py = S.py, bw = S.lw;
        const int kxs = strip * S.strip_pairs;
        const int kxe = min(kxs + S.strip_pairs, S.Kx);
        const int kx0 = kxs - NP + lane * NP;     // the producers' lane mapping (one halo lane per side)
        const bool st = lane >= 1 && kx0 < kxe && kx0 + NP <= bw;
        const int ky0 = chunk * S.chunk_pairs;
        const int ky1 = min(ky0 + S.chunk_pairs, S.Ky);
        constexpr int XES = (OUT == IN_U8) ? 1 : 2;
        unsigned char* xlane = (unsigned char*)S.x_base + S.x_off[item] * XES + (long long)(2 * kx0) * (3 * XES);
        const long long xpitch = S.x_row_bytes;
        int* planes = S.planes_out ? S.planes_out + S.planes_off[item] + 2 * kx0 : nullptr;
        const int planes_rs = S.planes_row_stride;
        const RawPack rp = make_raw_pack(raw);
        xlane += (long long)(2 * ky0 - py) * xpitch;
        if (planes) planes += (long long)(2 * ky0 - py) * planes_rs;
        auto emit = [&](const float2 (&x)[3][NP], unsigned char* xrow, int* prow) {
            bool done = false;
            if constexpr (J2K_ICT_FAST && OUT == IN_U8) { if (!prow) done = C3::template ict_fast_row<true>(rp, xrow, x); }
            if (!done) {
                int iv[3][NS];
#pragma unroll
                for (int c = 0; c < 3; c++)
#pragma unroll
                    for (int j = 0; j < NP; j++) { iv[c][2 * j] = __float2int_rn(x[c][j].x); iv[c][2 * j + 1] = __float2int_rn(x[c][j].y); }
                C3::store_final(S, raw, rp, xrow, prow, iv);
            }
        };
#pragma unroll 1
        for (int ky = ky0; ky < ky1; ky++) {
            const int slot = (int)(xp.cnt % (unsigned)xp.dx);
            mbar_wait(xp.full + 8 * slot, (xp.cnt / (unsigned)xp.dx) & 1u);
            const smem_t b = xp.data + slot * J2K_X3_STAGEB + lane * 16;
            float2 xe[3][NP], xo[3][NP];
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const uint4 e0 = lds128(b + c * 2048), e1 = lds128(b + c * 2048 + 512), o0 = lds128(b + c * 2048 + 1024), o1 = lds128(b + c * 2048 + 1536);
                xe[c][0] = make_float2(__uint_as_float(e0.x), __uint_as_float(e0.y)); xe[c][1] = make_float2(__uint_as_float(e0.z), __uint_as_float(e0.w));
                xe[c][2] = make_float2(__uint_as_float(e1.x), __uint_as_float(e1.y)); xe[c][3] = make_float2(__uint_as_float(e1.z), __uint_as_float(e1.w));
                xo[c][0] = make_float2(__uint_as_float(o0.x), __uint_as_float(o0.y)); xo[c][1] = make_float2(__uint_as_float(o0.z), __uint_as_float(o0.w));
                xo[c][2] = make_float2(__uint_as_float(o1.x), __uint_as_float(o1.y)); xo[c][3] = make_float2(__uint_as_float(o1.z), __uint_as_float(o1.w));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(xp.empty + 8 * slot);  // the values are in registers: the stage may be refilled
            xp.cnt++;
            const int re = 2 * ky - py, ro = re + 1;
            if (st) {
                if (re >= 0 && re < h) emit(xe, xlane, planes);
                if (ro < h) emit(xo, xlane + xpitch, planes ? planes + planes_rs : nullptr);
            }
            xlane += 2 * xpitch;
            if (planes) planes += 2 * planes_rs;
        }
    }
};

// job number -> (segment, item, chunk, strip, component); level-1 pixel jobs hold three numbers, ranges are padded to
// multiples of three (J.item < 0: a padding number, nothing to do)
// (fdiv = 3: the forward's first segment counts its items per component, S.n_items = 3 x pixel items)
__device__ __forceinline__ void x3_decode(const RingArgs& A, int job, RingJob& J, int& comp, int fdiv = 1) {
    int k = 0;
    while (k + 1 < A.nseg && job >= A.seg[k].job_end) k++;
    const RingSeg& S = A.seg[k];
    int local = job - S.job_begin;
    comp = 0;
    int n_items = S.n_items;
    if (S.first) { comp = local % 3; local /= 3; n_items /= fdiv; }
    const int ns = S.nstrips, nc = S.nchunks;
    J.seg = k;
    J.strip = local % ns;
    J.chunk = (local / ns) % nc;
    J.item = local / (ns * nc);
    if (J.item >= n_items) J.item = -1;
}

template <int OUT>
__global__ void __launch_bounds__(128, J2K_RING_MINB) inv3w_kernel(const __grid_constant__ RingArgs A) {
    J2K_SMEM_DECL(smem);
    const int lane = threadIdx.x & 31;
#ifdef J2K_EMU
    // the emulator runs warps one after the other: ONE warp plays all four roles of a pixel job in turn, with an exchange
    // ring deep enough for a whole chunk (the arithmetic and the index logic are the same; the barriers are no-ops)
    static thread_local unsigned char* xbig = nullptr;
    if (!xbig) xbig = (unsigned char*)malloc((size_t)J2K_X3_STAGEB * 4096);
    RingWarp rw;
    rw.ring = smem_handle(smem);
    rw.bars = rw.ring + J2K_X3_RING_BYTES;
    rw.phase = 0;
    rw.one = A.one;
    for (;;) {
        int job = 0;
        if (lane == 0) job = (int)atomicAdd(A.ctl, 3u);
        job = __shfl_sync(0xffffffffu, job, 0);
        if (job >= A.total_jobs) break;
        RingJob J; int comp;
        x3_decode(A, job, J, comp);
        const RingSeg& S = A.seg[J.seg];
        if (S.first) {
            if (J.item < 0) continue;
            XchPort xp{smem_handle(xbig), smem_handle(xbig), smem_handle(xbig), 0u, 4096};
            for (int c = 0; c < 3; c++) {
                xp.cnt = 0;
                InvRing<97, 4, 1, IN_F32, MCTK_NONE, J2K_X3_RING_BYTES, 0, 1>::run(S, A.raw, J.item, J.chunk, J.strip, rw, lane, &xp, c);
            }
            xp.cnt = 0;
            Inv3WConsumer<OUT>::run(S, A.raw, J.item, J.chunk, J.strip, xp, lane);
        } else {
            for (int q = 0; q < 3; q++) {
                x3_decode(A, job + q, J, comp);
                if (J.item < 0 || job + q >= A.seg[J.seg].job_end) continue;
                const RingSeg& T = A.seg[J.seg];
                InvRing<97, 4, 1, IN_F32, MCTK_NONE, J2K_X3_RING_BYTES>::run(T, A.raw, J.item, J.chunk, J.strip, rw, lane);
                if (T.has_waiters && lane == 0) red_release_add(A.ctl + T.done_base + J.item, 1u);
            }
        }
    }
    ring_retire(A, lane);
#else
    const int wib = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    __shared__ int s_job;
    RingWarp rw;
    rw.phase = 0;
    rw.one = A.one;
    if (wib < 3) {
        rw.ring = smem_handle(smem + wib * J2K_X3_WARP_SMEM);
        rw.bars = rw.ring + J2K_X3_RING_BYTES;
        if (lane == 0) {
            for (int q = 0; q < J2K_RING_MAXD; q++) mbar_init(rw.bars + 8 * q, 1);
        }
    } else {
        rw.ring = rw.bars = 0;
    }
    XchPort xp;
    xp.data = smem_handle(smem + 3 * J2K_X3_WARP_SMEM);
    xp.full = xp.data + J2K_X3_DX * J2K_X3_STAGEB;
    xp.empty = xp.full + 8 * J2K_X3_DX;
    xp.cnt = 0;
    xp.dx = J2K_X3_DX;
    if (threadIdx.x == 96) {
        for (int q = 0; q < J2K_X3_DX; q++) { mbar_init(xp.full + 8 * q, 3); mbar_init(xp.empty + 8 * q, 1); }
    }
    if (lane == 0) mbar_fence_init();
    __syncthreads();
    // Job triples are claimed per CTA through a barrier (one atomic per triple).  Measured alternative, rejected: dealing the
    // triples round-robin to the CTAs, so that the producers start the next pixel job while the consumer drains the last one
    // (no barrier, no pipeline drain at job boundaries) - C3 0.48 and C5 0.46 of the roofline against 0.54 / 0.53 with the
    // barrier and four-times-taller jobs: the dynamic claim keeps the CTAs on adjacent strips and balances the tail.
    for (;;) {
        __syncthreads();  // every warp has read the previous base
        if (threadIdx.x == 0) s_job = (int)atomicAdd(A.ctl, 3u);
        __syncthreads();
        const int base = s_job;
        if (base >= A.total_jobs) break;
        RingJob J; int comp;
        x3_decode(A, base + (wib < 3 ? wib : 0), J, comp);
        const RingSeg& S = A.seg[J.seg];
        if (S.first) {
            // the three numbers of a pixel job never straddle a segment: ranges are padded to multiples of three
            if (J.item < 0) continue;
            if (wib < 3) {
                if (S.dep_seg >= 0) {  // this component's LL plane of the level below is complete
                    if (lane == 0) {
                        const unsigned* d = A.ctl + A.seg[S.dep_seg].done_base + 3 * J.item + comp;
                        while (ld_relaxed(d) < (unsigned)S.dep_target) backoff();
                        (void)ld_acquire(d);
                        fence_proxy_async_global();
                    }
                    __syncwarp();
                }
                InvRing<97, 4, 1, IN_F32, MCTK_NONE, J2K_X3_RING_BYTES, 0, 1>::run(S, A.raw, J.item, J.chunk, J.strip, rw, lane, &xp, comp);
            } else {
                Inv3WConsumer<OUT>::run(S, A.raw, J.item, J.chunk, J.strip, xp, lane);
            }
        } else if (wib < 3 && J.item >= 0 && base + wib < S.job_end) {
            ring_wait_dep(A, S, J.item, lane);
            InvRing<97, 4, 1, IN_F32, MCTK_NONE, J2K_X3_RING_BYTES>::run(S, A.raw, J.item, J.chunk, J.strip, rw, lane);
            ring_signal(A, S, J.item, lane);
        }
    }
    ring_retire(A, lane);
#endif
}

// Measured and rejected (round 2, profiles/exp_r02_inv3q_rejected.log): this job split with four quads per CTA and setmaxnreg, as
// fwd3w_kernel below does for the forward direction (12 wavelet warps per SM instead of 9).  The inverse wavelet warp needs
// 166 registers and the pixel warp about 100; 12 x 144 + 4 x 80 and 12 x 152 + 4 x 56 both spill (232 / 416 bytes of stack)
// and run at 0.28 / 0.17 of the HBM peak on C5 against 0.51 for this kernel (bit-exact either way).

// ------------------------------------------------------------------ one-producer level-1 forward (ICT + 9/7, raw RGB)
//
// Mirror image of inv3w_kernel.  The component-split level 1 (FwdRing XC = 3) lets every component job unpack all three
// samples of a pixel and keep one ICT row: the raw bytes are converted three times (35.9 thread instructions per sample
// against 25.9 for a single-component frame).  Here a CTA of four warps shares a pixel job: warp 3 stages the raw rows
// (own TMA ring), converts them ONCE and writes the float32 Y / Cb / Cr rows into an exchange ring in shared memory;
// warps 0..2 each run the single-component pipeline (FwdRing<97, NP = 4, IN_F32> in exchange mode) for one component,
// reading their rows from that ring (mbarrier full / empty per stage, no CTA barrier in the loop).  Coarser levels: warps
// 0..2 take ordinary single-component jobs with private TMA rings that alias the (then idle) exchange ring, warp 3 idles.
// Jobs are taken three at a time per CTA; a level-1 pixel job occupies three consecutive job numbers, every segment's
// job range is padded to a multiple of three.
#ifndef J2K_FWD3W_DEFAULT
#define J2K_FWD3W_DEFAULT 1   // ICT + 9/7 raw RGB forward: one-producer level 1 (fwd3w_kernel) instead of three converting component jobs
#endif
// A CTA holds FOUR such quads (16 warps, one CTA per SM): warps 0..3 (warpgroup 0) are the producers of quads 0..3, warps
// 4 + 3q .. 6 + 3q the consumers of quad q.  The producers give registers back (setmaxnreg.dec) and the consumer
// warpgroups take them (setmaxnreg.inc), so the SM runs 12 wavelet warps with the full window state in registers PLUS the
// four converting warps - with uniform registers a fourth CTA of four warps does not fit (168 x 16 warps).  There is no
// CTA barrier in the job loop: a quad's producer claims the job triples (one atomic each), announces them to its
// consumers through a two-slot queue in shared memory and runs ahead into the next job while they finish the last one.
#define J2K_F3_QUADS 4
#define J2K_F3_REGS_PRODUCER 56
#define J2K_F3_REGS_CONSUMER 152   // 12 x 32 x 152 + 4 x 32 x 56 = 65536 = 512 threads x 128 registers at launch
#define J2K_F3_XBYTES (3 * J2K_RING_BYTES)       // exchange ring / the three private rings of the coarser levels
// per quad: exchange / private rings, the producer's raw ring, 4 x MAXD staging barriers, full / empty of the exchange ring,
// full / empty of the job queue, the queue's two job numbers
#define J2K_F3_QUAD_SMEM (J2K_F3_XBYTES + J2K_RING_BYTES + 4 * J2K_RING_MAXD * 8 + 2 * J2K_F3_DX * 8 + 4 * 8 + 16 + 160)
#define J2K_F3_CTA_SMEM (J2K_F3_QUADS * J2K_F3_QUAD_SMEM)
static_assert(J2K_F3_DX * J2K_F3_STAGEB <= J2K_F3_XBYTES, "exchange ring must fit the aliased staging");
static_assert(J2K_F3_QUAD_SMEM % 128 == 0, "quads keep the 128-byte alignment of the staging");

template <int IN>
__global__ void __launch_bounds__(J2K_F3_QUADS * 128, 1) fwd3w_kernel(const __grid_constant__ RingArgs A) {
    J2K_SMEM_DECL(smem);
    const int lane = threadIdx.x & 31;
    typedef FwdRing<97, 4, 1, IN, MCTK_ICT, 0, 3, 0, 2> Producer;
    typedef FwdRing<97, 4, 1, IN_F32, MCTK_NONE, 0, 1, 0, 1> Consumer;
    typedef FwdRing<97, 4, 1, IN_F32, MCTK_NONE, 0, 1, 0> Deep;
#ifdef J2K_EMU
    // the emulator runs warps one after the other: ONE warp plays all four roles of a pixel job in turn, with an exchange
    // ring deep enough for a whole chunk (the arithmetic and the index logic are the same; the barriers are no-ops)
    static thread_local unsigned char* xbig = nullptr;
    if (!xbig) xbig = (unsigned char*)malloc((size_t)J2K_F3_STAGEB * 256);
    RingWarp rw;
    rw.ring = smem_handle(smem);
    rw.bars = rw.ring + J2K_RING_BYTES;
    rw.phase = 0;
    rw.one = A.one;
    for (;;) {
        int job = 0;
        if (lane == 0) job = (int)atomicAdd(A.ctl, 3u);
        job = __shfl_sync(0xffffffffu, job, 0);
        if (job >= A.total_jobs) break;
        RingJob J; int comp;
        x3_decode(A, job, J, comp, 3);
        const RingSeg& S = A.seg[J.seg];
        if (S.first) {
            if (J.item < 0) continue;
            FxPort xp{smem_handle(xbig), smem_handle(xbig), smem_handle(xbig), 0u, 256};
            Producer::run(S, A.raw, J.item, J.chunk, J.strip, rw, lane, &xp, 0);
            for (int c = 0; c < 3; c++) {
                xp.cnt = 0;
                Consumer::run(S, A.raw, J.item, J.chunk, J.strip, rw, lane, &xp, c);
                if (S.has_waiters && lane == 0) red_release_add(A.ctl + S.done_base + 3 * J.item + c, 1u);
            }
        } else {
            for (int q = 0; q < 3; q++) {
                x3_decode(A, job + q, J, comp, 3);
                if (J.item < 0 || job + q >= A.seg[J.seg].job_end) continue;
                const RingSeg& T = A.seg[J.seg];
                Deep::run(T, A.raw, J.item, J.chunk, J.strip, rw, lane);
                if (T.has_waiters && lane == 0) red_release_add(A.ctl + T.done_base + J.item, 1u);
            }
        }
    }
    ring_retire(A, lane);
#else
    const int wib = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const bool producer = wib < J2K_F3_QUADS;
    const int quad = producer ? wib : (wib - J2K_F3_QUADS) / 3;
    const int role = producer ? 3 : (wib - J2K_F3_QUADS) % 3;   // consumers: their component
    unsigned char* qs = smem + quad * J2K_F3_QUAD_SMEM;
    RingWarp rw;
    rw.phase = 0;
    rw.one = A.one;
    // consumers: private staging inside the exchange region (coarser levels only); producer: the raw-row staging behind it
    rw.ring = smem_handle(qs + role * J2K_RING_BYTES);
    rw.bars = smem_handle(qs + 4 * J2K_RING_BYTES + role * J2K_RING_MAXD * 8);
    if (lane == 0) {
        for (int q = 0; q < J2K_RING_MAXD; q++) mbar_init(rw.bars + 8 * q, 1);
    }
    FxPort xp;
    xp.data = smem_handle(qs);
    xp.full = smem_handle(qs + 4 * J2K_RING_BYTES + 4 * J2K_RING_MAXD * 8);
    xp.empty = xp.full + 8 * J2K_F3_DX;
    xp.cnt = 0;
    xp.dx = J2K_F3_DX;
    const smem_t jq_full = xp.empty + 8 * J2K_F3_DX, jq_empty = jq_full + 16, jq_job = jq_empty + 16;
    if (producer && lane == 0) {
        for (int q = 0; q < J2K_F3_DX; q++) { mbar_init(xp.full + 8 * q, 1); mbar_init(xp.empty + 8 * q, 3); }
        for (int q = 0; q < 2; q++) { mbar_init(jq_full + 8 * q, 1); mbar_init(jq_empty + 8 * q, 3); }
    }
    if (lane == 0) mbar_fence_init();
    __syncthreads();
    unsigned jcnt = 0;  // job triples announced / taken so far
    if (producer) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(J2K_F3_REGS_PRODUCER));
        for (;;) {
            int base = 0;
            if (lane == 0) base = (int)atomicAdd(A.ctl, 3u);
            base = __shfl_sync(0xffffffffu, base, 0);
            const int slot = (int)(jcnt & 1u);
            mbar_wait(jq_empty + 8 * slot, ((jcnt >> 1) & 1u) ^ 1u);
            if (lane == 0) { sts32(jq_job + 4 * slot, (unsigned)base); mbar_arrive(jq_full + 8 * slot); }
            jcnt++;
            if (base >= A.total_jobs) break;
            RingJob J; int comp;
            x3_decode(A, base, J, comp, 3);
            const RingSeg& S = A.seg[J.seg];
            if (S.first && J.item >= 0) Producer::run(S, A.raw, J.item, J.chunk, J.strip, rw, lane, &xp, 0);
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(J2K_F3_REGS_CONSUMER));
        bool deep_seen = false;
        for (;;) {
            const int slot = (int)(jcnt & 1u);
            mbar_wait(jq_full + 8 * slot, (jcnt >> 1) & 1u);
            const int base = __shfl_sync(0xffffffffu, (int)lds32(jq_job + 4 * slot), 0);
            __syncwarp();
            if (lane == 0) mbar_arrive(jq_empty + 8 * slot);
            jcnt++;
            if (base >= A.total_jobs) break;
            RingJob J; int comp;
            x3_decode(A, base + role, J, comp, 3);
            const RingSeg& S = A.seg[J.seg];
            if (S.first) {
                // the three numbers of a pixel job never straddle a segment: ranges are padded to multiples of three
                if (J.item < 0) continue;
                Consumer::run(S, A.raw, J.item, J.chunk, J.strip, rw, lane, &xp, comp);
                if (S.has_waiters) {
                    __syncwarp();
                    if (lane == 0) red_release_add(A.ctl + S.done_base + 3 * J.item + comp, 1u);
                }
            } else {
                if (!deep_seen) {
                    // level 1 is over for this quad (the list is level-major): the exchange region becomes the three private
                    // staging rings once all three consumers have read its last stage (generic-proxy accesses before, bulk copies after)
                    fence_proxy_async();
                    asm volatile("bar.sync %0, 96;" ::"r"(1 + quad) : "memory");
                    deep_seen = true;
                }
                if (J.item >= 0 && base + role < S.job_end) {
                    ring_wait_dep(A, S, J.item, lane);
                    Deep::run(S, A.raw, J.item, J.chunk, J.strip, rw, lane);
                    ring_signal(A, S, J.item, lane);
                }
            }
        }
    }
    ring_retire(A, lane);
#endif
}

}  // namespace j2k
