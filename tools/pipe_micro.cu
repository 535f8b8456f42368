// Micro-benchmark: issue rate of packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2) against scalar FFMA and against an
// ALU-pipe instruction (PRMT) on sm_100a, 12 warps per SM as in the ring kernels.  Prints warp instructions per cycle per
// SM sub-partition.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_ub/pipe_micro tools/pipe_micro.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N_IT 4096
template <int MODE> __global__ void __launch_bounds__(128, 3) k(float* out, float a, float b, unsigned sel, long long* cyc) {
    float2 x[8];
    float s[16];
    unsigned u[8];
    for (int i = 0; i < 8; i++) { x[i] = make_float2(threadIdx.x + i, i); u[i] = threadIdx.x * 77u + i; }
    for (int i = 0; i < 16; i++) s[i] = threadIdx.x + i;
    const float2 A = make_float2(a, a), B = make_float2(b, b);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < N_IT; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (MODE == 0) {          // 8 independent FFMA2
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = __ffma2_rn(x[i], A, B);
            } else if (MODE == 1) {   // 16 independent scalar FFMA (register operands)
#pragma unroll
                for (int i = 0; i < 16; i++) s[i] = __fmaf_rn(s[i], a, b);
            } else if (MODE == 2) {   // 8 FMUL2
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = __fmul2_rn(x[i], A);
            } else if (MODE == 3) {   // 8 FADD2
#pragma unroll
                for (int i = 0; i < 8; i++) x[i] = __fadd2_rn(x[i], B);
            } else if (MODE == 4) {   // 8 PRMT (ALU pipe)
#pragma unroll
                for (int i = 0; i < 8; i++) u[i] = __byte_perm(u[i], 0x4B000000u, sel);
            } else if (MODE == 5) {   // 8 FFMA2 + 8 PRMT interleaved
#pragma unroll
                for (int i = 0; i < 8; i++) { x[i] = __ffma2_rn(x[i], A, B); u[i] = __byte_perm(u[i], 0x4B000000u, sel); }
            } else if (MODE == 6) {   // 8 FFMA2 + 16 PRMT
#pragma unroll
                for (int i = 0; i < 8; i++) { x[i] = __ffma2_rn(x[i], A, B); u[i] = __byte_perm(u[i], 0x4B000000u, sel); u[i] = __byte_perm(u[i], 0x4C000000u, sel); }
            } else if (MODE == 7) {   // 16 scalar FFMA + 16 PRMT
#pragma unroll
                for (int i = 0; i < 16; i++) { s[i] = __fmaf_rn(s[i], a, b); u[i & 7] = __byte_perm(u[i & 7], 0x4B000000u + i, sel); }
            }
        }
    }
    long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < 8; i++) acc += x[i].x + x[i].y + __uint_as_float(u[i]);
    for (int i = 0; i < 16; i++) acc += s[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE> void run(const char* name, double per_it) {
    float* out; long long* cyc;
    const int grid = 148 * 3;
    cudaMalloc(&out, grid * 128 * 4); cudaMalloc(&cyc, grid * 8);
    for (int rep = 0; rep < 2; rep++) k<MODE><<<grid, 128>>>(out, 1.0001f, 0.5f, 0x7440u, cyc);
    cudaDeviceSynchronize();
    long long h[148 * 3]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double m = 0; for (int i = 0; i < grid; i++) m += h[i]; m /= grid;
    // 12 warps per SM = 3 per sub-partition, each issuing per_it * N_IT instructions in m cycles
    printf("%-28s %.3f warp-instr / cycle / SMSP  (%.0f cycles)\n", name, 3.0 * per_it * N_IT / m, m);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("FFMA2", 32); run<1>("FFMA scalar", 64); run<2>("FMUL2", 32); run<3>("FADD2", 32); run<4>("PRMT", 32);
    run<5>("FFMA2 + PRMT 1:1", 64); run<6>("FFMA2 + PRMT 1:2", 96); run<7>("FFMA + PRMT 1:1", 128);
    return 0;
}
