// Kernel-only timing harness for ht_decode_kernel (tools/ht_micro_data.py writes the input): compiles in seconds because it
// instantiates nothing but the HT kernel.  Checks the decoded planes against the coefficients in the file, then times launches
// of `frames` copies of the frame with CUDA events.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -o ht_micro ht_micro.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../include/j2k_b200.h"
#include "../go-dicom-codec_b200/csrc/j2k_kernels.cuh"
#include "../go-dicom-codec_b200/csrc/j2k_pointwise.cuh"
#include "../go-dicom-codec_b200/csrc/j2k_ring.cuh"
#include "../go-dicom-codec_b200/csrc/j2k_ht.cuh"
using namespace j2k;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
int main(int argc, char** argv) {
    const char* path = argc > 1 ? argv[1] : "tools/_ht/ht_2048_12.bin";
    const int frames = argc > 2 ? atoi(argv[2]) : 8, steps = argc > 3 ? atoi(argv[3]) : 20, warps = argc > 4 ? atoi(argv[4]) : 4;
    FILE* f = fopen(path, "rb");
    if (!f) { printf("no %s\n", path); return 1; }
    long long hd[6];
    if (fread(hd, 8, 6, f) != 6) return 1;
    const long long nb = hd[0], W = hd[1], H = hd[2], cbw = hd[3], cbh = hd[4], nbytes = hd[5];
    std::vector<BlockEntry> tab(nb);
    std::vector<HtBlock> rec(nb);
    std::vector<unsigned char> st(nbytes);
    std::vector<int> co(W * H);
    if (fread(tab.data(), sizeof(BlockEntry), nb, f) != (size_t)nb || fread(rec.data(), sizeof(HtBlock), nb, f) != (size_t)nb ||
        fread(st.data(), 1, nbytes, f) != (size_t)nbytes || fread(co.data(), 4, W * H, f) != (size_t)(W * H)) return 1;
    fclose(f);
    std::vector<HtBlock> recs(nb * frames);
    for (int fr = 0; fr < frames; fr++) memcpy(&recs[fr * nb], rec.data(), nb * sizeof(HtBlock));
    unsigned char* d_st; HtBlock* d_rec; BlockEntry* d_tab; int* d_out; int* d_status;
    CK(cudaMalloc(&d_st, nbytes)); CK(cudaMalloc(&d_rec, recs.size() * sizeof(HtBlock))); CK(cudaMalloc(&d_tab, nb * sizeof(BlockEntry)));
    CK(cudaMalloc(&d_out, (size_t)frames * W * H * 4)); CK(cudaMalloc(&d_status, nb * frames * 4));
    CK(cudaMemcpy(d_st, st.data(), nbytes, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_rec, recs.data(), recs.size() * sizeof(HtBlock), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_tab, tab.data(), nb * sizeof(BlockEntry), cudaMemcpyHostToDevice));
    const long long total = nb * frames;
    const int wsm = ht_warp_smem((int)cbw), smem = warps * wsm, scw = ht_scratch_words((int)cbw, (int)cbh), qs = (int)(cbw + 1) / 2;
    const int only = argc > 5 ? atoi(argv[5]) : 0;   // 1: VLC kernel only, 2: MagSgn kernel only (after one full run)
    unsigned* d_sc; CK(cudaMalloc(&d_sc, (size_t)total * scw * 4));
    const unsigned grid = (unsigned)((total + warps - 1) / warps);
    cudaStream_t s; CK(cudaStreamCreate(&s));
    int phase = 0;
    auto launch = [&]() {
        if (phase != 2) ht_vlc_kernel<<<(unsigned)((total + 31) / 32), 32, 0, s>>>(d_st, d_rec, d_tab, (int)nb, total, d_sc, scw, qs);
        if (phase != 1) ht_magsgn_kernel<<<grid, warps * 32, smem, s>>>(d_st, d_rec, d_tab, (int)nb, total, W * H, d_sc, scw, qs, d_out, 1, d_status, wsm);
    };
    CK(cudaMemsetAsync(d_out, 0xEE, (size_t)frames * W * H * 4, s));
    launch();
    CK(cudaStreamSynchronize(s));
    std::vector<int> back(W * H);
    long long bad = 0;
    for (int fr = 0; fr < frames; fr += frames - 1 > 0 ? frames - 1 : 1) {
        CK(cudaMemcpy(back.data(), d_out + (size_t)fr * W * H, W * H * 4, cudaMemcpyDeviceToHost));
        for (long long i = 0; i < W * H; i++) bad += back[i] != co[i];
    }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ht_magsgn_kernel, warps * 32, smem);
    phase = only;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; i++) launch();
    cudaEventRecord(e0, s);
    for (int i = 0; i < steps; i++) launch();
    cudaEventRecord(e1, s);
    CK(cudaStreamSynchronize(s));
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= steps;
    printf("{\"file\": \"%s\", \"frames\": %d, \"blocks\": %lld, \"mismatches\": %lld, \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"smem_per_cta\": %d, "
           "\"ms\": %.4f, \"Mpixel_s\": %.0f, \"compressed_GBps\": %.1f}\n", path, frames, total, bad, warps, occ, smem, ms,
           frames * W * H / 1e3 / ms, frames * (double)nbytes / 1e6 / ms);
    return bad != 0;
}
