// tests/emu/cuda_runtime.h — TEST INFRASTRUCTURE ONLY.
//
// A single-threaded CPU stand-in for the slice of the CUDA runtime and device intrinsics that
// go-dicom-codec_b200/csrc uses, so that the REAL kernel and host sources can be compiled with
// g++ and exercised by the `-m "not gpu"` tests in a container without a GPU (host logic, plan
// building, kernel index arithmetic and lifting order against the oracle).  Each warp is run as
// 32 ucontext fibers that switch at every __shfl_*_sync.  It is never part of the product, is not
// a fallback (libj2kb200.so does not contain it), and nothing outside tests/ may load it.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>

#include <algorithm>
#include <functional>

#define J2K_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __grid_constant__
#define __launch_bounds__(...)
#define __restrict__

struct emu_dim3 { unsigned x = 1, y = 1, z = 1; };
extern emu_dim3 threadIdx, blockIdx, blockDim, gridDim;

struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
static inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }

using std::max;
using std::min;

template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline int __float2int_rn(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (-2147483647 - 1);
    return (int)lrintf(f);
}
static inline long long __float2ll_rz(float f) {  // saturating, NaN -> 0 (the device intrinsic's behaviour)
    if (f != f) return 0;
    if (f >= 9223372036854775808.0f) return 9223372036854775807LL;
    if (f <= -9223372036854775808.0f) return (-9223372036854775807LL - 1);
    return (long long)f;
}

namespace emu {
uint32_t shfl_exchange(uint32_t v, int src_lane);  // yields the calling fiber
int lane_id();
void launch(unsigned grid, unsigned block, const std::function<void()>& body);
unsigned char* smem();  // per-CTA dynamic shared memory (warps of a CTA run one after the other)
}  // namespace emu

static inline void __syncwarp(unsigned = 0xffffffffu) { emu::shfl_exchange(0, emu::lane_id()); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src) {
    static_assert(sizeof(T) == 4, "emu shuffles move 32-bit values");
    uint32_t b; memcpy(&b, &v, 4);
    b = emu::shfl_exchange(b, src & 31);
    T r; memcpy(&r, &b, 4); return r;
}
static inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline unsigned atomicOr(unsigned* p, unsigned v) { unsigned o = *p; *p = o | v; return o; }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline void __threadfence() {}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
    sh &= 31;
    return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
}
static inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {
    unsigned long long src = ((unsigned long long)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; i++) {
        unsigned sel = (s >> (4 * i)) & 0xF;
        unsigned b = (unsigned)(src >> (8 * (sel & 7))) & 0xFF;
        if (sel & 8) b = (b & 0x80) ? 0xFF : 0x00;
        r |= b << (8 * i);
    }
    return r;
}

template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) {
    static_assert(sizeof(T) == 4, "emu shuffles move 32-bit values");
    int lane = emu::lane_id(), src = lane + d > 31 ? lane : lane + d;
    uint32_t b; memcpy(&b, &v, 4);
    b = emu::shfl_exchange(b, src);
    T r; memcpy(&r, &b, 4); return r;
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) {
    static_assert(sizeof(T) == 4, "emu shuffles move 32-bit values");
    int lane = emu::lane_id(), src = lane - d < 0 ? lane : lane - d;
    uint32_t b; memcpy(&b, &v, 4);
    b = emu::shfl_exchange(b, src);
    T r; memcpy(&r, &b, 4); return r;
}
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int d) {
    static_assert(sizeof(T) == 4, "emu shuffles move 32-bit values");
    uint32_t b; memcpy(&b, &v, 4);
    b = emu::shfl_exchange(b, (emu::lane_id() ^ d) & 31);
    T r; memcpy(&r, &b, 4); return r;
}

// ---- runtime API subset (everything is synchronous host memory)
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidDeviceFunction = 98, cudaErrorMemoryAllocation = 2 };
typedef struct emu_stream* cudaStream_t;
typedef struct emu_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocPortable = 1 };

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = getenv("J2K_EMU_NO_DEVICE") ? 0 : 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { return posix_memalign(p, 256, n ? n : 1) ? cudaErrorMemoryAllocation : cudaSuccess; }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t) {
    for (size_t r = 0; r < h; r++) memmove((char*)d + r * dp, (const char*)s + r * sp, w);
    return cudaSuccess;
}
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)malloc(1); return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t s) { free(s); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)malloc(1); return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
static inline cudaError_t cudaEventDestroy(cudaEvent_t e) { free(e); return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }

static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }

#define J2K_LAUNCH(kernel, grid, block, stream, ...) emu::launch((grid), (block), [&]() { kernel(__VA_ARGS__); })
#define J2K_LAUNCH_SMEM(kernel, grid, block, smem, stream, ...) emu::launch((grid), (block), [&]() { kernel(__VA_ARGS__); })
