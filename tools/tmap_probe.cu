// Probe: 1-D tiled tensor-map copies (UTMALDG.1D) from arbitrary element coordinates, descriptor inside a large __grid_constant__ struct.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_ub/tmap_probe tools/tmap_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
struct Big { alignas(64) CUtensorMap tm[4]; int pad[PADN]; int which; };
__global__ void k(const __grid_constant__ Big A, int coord, unsigned short* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ int dummy[257];
    dummy[threadIdx.x] = 1;
    unsigned bar = (unsigned)__cvta_generic_to_shared(smem + 1024);
    unsigned dst = (unsigned)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("fence.proxy.async.shared::cta;");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(512));
        const CUtensorMap* tm = A.tm + A.which;
        asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];" ::"r"(dst),
                     "l"((unsigned long long)tm), "r"(coord), "r"(bar) : "memory");
    }
    unsigned ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(bar));
    out[threadIdx.x] = ((unsigned short*)smem)[threadIdx.x] + (unsigned short)(dummy[threadIdx.x] - 1);
}
int main() {
    const int N = 100000;
    std::vector<unsigned short> h(N);
    for (int i = 0; i < N; i++) h[i] = (unsigned short)(i % 60000 + 1);
    unsigned short *d, *o;
    cudaMalloc(&d, N * 2); cudaMalloc(&o, 512);
    cudaMemcpy(d, h.data(), N * 2, cudaMemcpyHostToDevice);
    void* f = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
    typedef CUresult (*Enc)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    Big A; memset(&A, 0, sizeof A);
    A.which = 2;
    cuuint64_t dims[1] = {(cuuint64_t)N}, strides[1] = {0}; cuuint32_t box[1] = {256}, es[1] = {1};
    CUresult r = ((Enc)f)(&A.tm[2], CU_TENSOR_MAP_DATA_TYPE_UINT16, 1, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("sizeof(Big)=%zu encode=%d q=%d\n", sizeof(Big), (int)r, (int)q);
    int coords[] = {0, 8, 3, 1071, -5, N - 100};
    for (int c : coords) {
        k<<<1, 256, 2048>>>(A, c, o);
        cudaError_t e = cudaDeviceSynchronize();
        unsigned short got[256]; cudaMemcpy(got, o, 512, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < 256; i++) { long long idx = (long long)c + i; unsigned short want = (idx >= 0 && idx < N) ? h[idx] : 0; if (got[i] != want) bad++; }
        printf("coord %d: %s, %d mismatches (first %u %u %u)\n", c, cudaGetErrorString(e), bad, got[0], got[1], got[2]);
        if (e != cudaSuccess) break;
    }
    return 0;
}
