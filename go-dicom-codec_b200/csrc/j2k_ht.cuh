// HTJ2K cleanup-pass block decoder on the device (decode side).
//
// Replaces, per code-block, HTDecoder.Decode -> decodeOpenJPHCleanup
// (/root/reference/jpeg2000/htj2k/decoder.go:43-58, openjph_cleanup_decoder.go:115-161) and, in the same pass,
// TileDecoder.assembleSubbands (jpeg2000/t2/tile_decoder.go:840-883): the decoded samples go straight into the Mallat
// coefficient plane the inverse ring kernel reads (or into the block-major plane of the code-block interface).  What crosses
// PCIe on the decode side is then the compressed cleanup segments + 16 bytes per block instead of 4 bytes per sample.
//
// Two kernels per batch of code-blocks:
//   1. ht_vlc_kernel -- VLC / MEL / UVLC (openjph_cleanup_decoder.go:180-276).  Inherently serial inside a block (every codeword
//      length depends on the context of the quads before it), so the parallelism is ACROSS blocks: one THREAD per code-block,
//      32 blocks per warp walking their quad pairs in lock-step (same loop structure, data-dependent lengths).  A first version
//      ran this phase on lane 0 of a warp-per-block kernel; ncu showed it issue-bound at 123 k warp instructions per 64 x 64
//      block with 4.7 of 32 lanes active -- a warp instruction costs the same issue slot for 1 lane as for 32.  Each thread
//      leaves (inf, u_q) per quad -- the `scratch` array of the reference -- in a global buffer (4 B per quad, ~1 B per
//      sample); the previous quad row's significance pattern, which the context needs, lives in a per-thread local array.
//      The reverse VLC reader and the MEL reader follow vlc_reverse_decoder.go:18-100 and openjph_cleanup_decoder.go:25-101
//      (byte-wise instead of 4-byte chunks; the bit values are the same, see the notes at the readers).
//   2. ht_magsgn_kernel -- MagSgn (openjph_cleanup_decoder.go:278-372, :432-447), one WARP per code-block: the number of bits
//      of every sample is known once U_q is, and U_q of a quad row depends only on the row above (the exponent predictor), so
//      a quad row is decoded by 32 lanes at once: one lane per quad, bit offsets by a warp prefix sum.  Random access into
//      the MagSgn stream needs the byte-stuffing removed first: the warp un-stuffs 512 source bytes at a time (a byte after
//      0xFF carries 7 bits, magsgn.go:160-204; beyond the end the stream reads as 0xFF) into a 16 Kbit ring in shared memory,
//      each lane depositing its 16 bytes at the bit position a prefix sum over the lanes' bit counts gives it.
// Errors (U_q beyond missing_msbs + 2, bad Scup, Kmax = 0 ...) zero the block, as TileDecoder.decodeCodeBlock does
// (t2/tile_decoder.go:718-721), and are reported per block in `status`.
#pragma once

namespace j2k {

#define J2K_HT_TABLE static __device__ const
#include "j2k_ht_tables.inc"
#undef J2K_HT_TABLE

struct HtBlock {                 // == j2k_ht_cblk (include/j2k_b200.h)
    unsigned long long offset;   // first byte of the cleanup segment in the byte stream
    unsigned length;             // Lcup; 0 = no data: zero coefficients
    unsigned char kmax, mmsb;    // bandNumbps (t2/bitplane.go:22-61), zero bit-planes (t2/tile_decoder.go:691-699)
    unsigned short reserved;
};

#define HT_RING_WORDS 512        // un-stuffed MagSgn bits: 16 Kbit ring per warp
#define HT_RING_MASK (HT_RING_WORDS - 1)

#define HT_MAX_QW 512            // quads per row of the widest code-block (1024 samples)

// global scratch per code-block: one 32-bit word (inf | u_q << 16) per quad, rows of cbw / 2 quads
__host__ __device__ inline int ht_scratch_words(int cbw, int cbh) { return ((cbw + 1) / 2) * ((cbh + 1) / 2); }
// shared memory per warp of the MagSgn kernel (bytes): two rows of exponent-predictor state + the bit ring
__host__ __device__ inline int ht_warp_smem(int cbw) { return 2 * ((cbw + 1) / 2 + 2) * 4 + HT_RING_WORDS * 4; }

__device__ __forceinline__ int ht_warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// ---- phase 1 readers (lane 0 only)

// vlc_reverse_decoder.go: bytes from the end of the cleanup segment towards its start, LSB first; a byte whose low 7 bits are
// all ones carries 7 bits when the byte read before it was > 0x8F.  The reference reads 4 bytes per chunk and counts missing
// bytes of the last chunk as zero bytes; read byte-wise, the bits beyond the last real byte are zeros as well (peek returns
// tmp with zeros above `num`, advance beyond `num` clears it), so every peek returns the same 32 bits.
struct HtRev {
    const unsigned char* data; int pos; unsigned long long tmp; int num; bool unstuff;
    __device__ __forceinline__ void init(const unsigned char* d, int len) {
        data = d;
        pos = len - 2;
        const unsigned b = data[pos--];
        tmp = b >> 4;
        num = 4 - (((tmp & 7) == 7) ? 1 : 0);
        unstuff = (b | 0x0F) > 0x8F;
    }
    // (measured: four bytes per step, loaded together as the reference's readChunk does, is slower here -- 0.78 -> 1.02 ms for
    // 65536 blocks: the extra predicated work costs more than the exposed load latency it saves)
    __device__ __forceinline__ void fill() {
        while (num < 32 && pos >= 0) {
            const unsigned b = data[pos--];
            const int bits = (unstuff && (b & 0x7F) == 0x7F) ? 7 : 8;
            tmp |= (unsigned long long)b << num;
            num += bits;
            unstuff = b > 0x8F;
        }
    }
    __device__ __forceinline__ unsigned peek() { fill(); return (unsigned)tmp; }
    __device__ __forceinline__ void advance(int n) {
        if (n <= 0) return;
        fill();
        if (n > num) { tmp = 0; num = 0; return; }
        tmp >>= n;
        num -= n;
    }
};

// openjph_cleanup_decoder.go:8-101: forward, MSB first, 7 bits after a 0xFF byte, the last counted byte ORed with 0x0F, ones
// once `size` is used up.  The reference decodes eight runs ahead; the runs do not depend on anything but the MEL bytes, so
// decoding them one at a time gives the same sequence.
struct HtMel {
    const unsigned char* data; int len, pos, size, k, nbits; unsigned bits; bool unstuff;
    __device__ __forceinline__ void init(const unsigned char* d, int n) { data = d; len = n; pos = 0; size = n - 1; k = 0; nbits = 0; bits = 0; unstuff = false; }
    __device__ __forceinline__ int bit() {
        while (nbits == 0) {
            if (size <= 0) return 1;
            unsigned d = 0xFF;
            if (pos < len) {
                d = data[pos++];
                if (size == 1) d |= 0x0F;
                size--;
            }
            const int valid = unstuff ? 7 : 8;
            bits = d & ((1u << valid) - 1);
            nbits = valid;
            unstuff = d == 0xFF;
        }
        nbits--;
        return (int)((bits >> nbits) & 1);
    }
    __device__ __forceinline__ int run() {
        const int eval = (int)((0x5433222111000ULL >> (4 * k)) & 0xF);   // MelE, mel_spec.go:8-22
        int r;
        if (bit()) {
            r = ((1 << eval) - 1) << 1;
            if (k < 12) k++;
        } else {
            r = 0;
            for (int i = 0; i < eval; i++) r = (r << 1) | bit();
            if (k > 0) k--;
            r = (r << 1) + 1;
        }
        return r;
    }
};

struct HtCleanup {
    HtRev vlc; HtMel mel; int run;
    // openjph_cleanup_decoder.go:169-178
    __device__ __forceinline__ unsigned zero_run(unsigned entry) {
        run -= 2;
        if (run != -1) entry = 0;
        if (run < 0) run = mel.run();
        return entry;
    }
    // :258-276
    __device__ __forceinline__ void uvlc(const unsigned short* tbl, int mode, int& u0, int& u1) {
        unsigned v = vlc.peek();
        const unsigned e = tbl[mode + (int)(v & 0x3F)];
        vlc.advance((int)(e & 7));
        v = vlc.peek();
        const int total_suffix = (int)((e >> 3) & 0xF);
        const int t = (int)(v & ((1u << total_suffix) - 1));
        vlc.advance(total_suffix);
        const int u0suf = (int)((e >> 7) & 7);
        u0 = (int)((e >> 10) & 7) + (t & ((1 << u0suf) - 1));
        u1 = (int)((e >> 13) & 7) + (t >> u0suf);
    }
};

// decoder.go:44-50, openjph_cleanup_decoder.go:116-133 + parseStandardSegments (decoder.go:60-70): 0 and the suffix length, or
// the error code of the block.  Both kernels evaluate it.
__device__ __forceinline__ int ht_validate(const unsigned char* __restrict__ bytes, const HtBlock& d, int& scup) {
    scup = 0;
    const int lcup = (int)d.length;
    if (lcup == 0) return 0;
    if (d.kmax == 0) return -1;
    if (d.mmsb >= 30 || lcup < 2) return -2;
    const unsigned char* cb = bytes + d.offset;
    scup = ((int)cb[lcup - 1] << 4) | (cb[lcup - 2] & 0x0F);
    if (scup < 2 || scup > lcup || scup > 4079) return -2;
    return 0;
}

// Kernel 1: openjph_cleanup_decoder.go:180-256 (decodeOpenJPHInitialRow, decodeOpenJPHRemainingRows), one thread per block.
// `sc`: ht_scratch_words(cbw, cbh) words per block, row stride qs = cbw / 2 quads.  The reference keeps (inf, u_q) of the
// row above in its scratch array and reads three entries of it per quad pair (sp - sstr, + 2, + 4); here those are
// prow[q0], prow[q0 + 1], prow[q0 + 2], overwritten with the current row's patterns once the pair is done (entries past the
// last quad stay zero: the reference's sentinels).
__global__ void __launch_bounds__(32) ht_vlc_kernel(const unsigned char* __restrict__ bytes, const HtBlock* __restrict__ descs,
                                                    const BlockEntry* __restrict__ tab, int nblocks, long long total,
                                                    unsigned* __restrict__ sc_all, int sc_words, int qs) {
    const long long wid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (wid >= total) return;
    const int bi = (int)(wid % nblocks);
    const int width = tab[bi].w, height = tab[bi].h;
    const HtBlock d = descs[wid];
    int scup;
    if (d.length == 0 || ht_validate(bytes, d, scup) != 0) return;
    unsigned* sc = sc_all + wid * sc_words;
    const unsigned char* cleanup = bytes + d.offset + (d.length - scup);
    const int qw = (width + 1) >> 1;
    unsigned char prow[HT_MAX_QW + 4];
    for (int i = 0; i < qw + 3; i++) prow[i] = 0;
    HtCleanup st;
    st.mel.init(cleanup, scup);
    st.vlc.init(cleanup, scup);
    st.run = st.mel.run();
    int cq = 0;
    for (int x = 0, q0 = 0; x < width; q0 += 2) {
        unsigned t0 = HT_VLC_TBL0[cq + (int)(st.vlc.peek() & 0x7F)];
        if (cq == 0) t0 = st.zero_run(t0);
        x += 2;
        cq = (int)(((t0 & 0x10) << 3) | ((t0 & 0xE0) << 2));
        st.vlc.advance((int)(t0 & 7));

        unsigned t1 = HT_VLC_TBL0[cq + (int)(st.vlc.peek() & 0x7F)];
        if (cq == 0 && x < width) t1 = st.zero_run(t1);
        if (x >= width) t1 = 0;
        x += 2;
        cq = (int)(((t1 & 0x10) << 3) | ((t1 & 0xE0) << 2));
        st.vlc.advance((int)(t1 & 7));

        int mode = (int)(((t0 & 0x8) << 3) | ((t1 & 0x8) << 4));
        if (mode == 0xC0) {
            st.run -= 2;
            if (st.run == -1) mode += 0x40;
            if (st.run < 0) st.run = st.mel.run();
        }
        int u0, u1;
        st.uvlc(HT_UVLC_TBL0, mode, u0, u1);
        sc[q0] = (t0 & 0xFFFF) | ((unsigned)(1 + u0) << 16);
        if (q0 + 1 < qw) sc[q0 + 1] = (t1 & 0xFFFF) | ((unsigned)(1 + u1) << 16);
        prow[q0] = (unsigned char)(t0 & 0xF0);
        prow[q0 + 1] = (unsigned char)(t1 & 0xF0);
    }
    for (int y = 2; y < height; y += 2) {
        cq = 0;
        unsigned* row = sc + (y >> 1) * qs;
        for (int x = 0, q0 = 0; x < width; q0 += 2) {
            const unsigned p0 = prow[q0], p1 = prow[q0 + 1], p2 = prow[q0 + 2];
            cq |= (int)(((p0 & 0xA0) << 2) | ((p1 & 0x20) << 4));
            unsigned t0 = HT_VLC_TBL1[cq + (int)(st.vlc.peek() & 0x7F)];
            if (cq == 0) t0 = st.zero_run(t0);
            x += 2;
            cq = (int)(((t0 & 0x40) << 2) | ((t0 & 0x80) << 1));
            cq |= (int)(p0 & 0x80);
            cq |= (int)(((p1 & 0xA0) << 2) | ((p2 & 0x20) << 4));
            st.vlc.advance((int)(t0 & 7));

            unsigned t1 = HT_VLC_TBL1[cq + (int)(st.vlc.peek() & 0x7F)];
            if (cq == 0 && x < width) t1 = st.zero_run(t1);
            if (x >= width) t1 = 0;
            x += 2;
            cq = (int)(((t1 & 0x40) << 2) | ((t1 & 0x80) << 1));
            cq |= (int)(p1 & 0x80);
            st.vlc.advance((int)(t1 & 7));

            int u0, u1;
            st.uvlc(HT_UVLC_TBL1, (int)(((t0 & 0x8) << 3) | ((t1 & 0x8) << 4)), u0, u1);
            row[q0] = (t0 & 0xFFFF) | ((unsigned)u0 << 16);
            if (q0 + 1 < qw) row[q0 + 1] = (t1 & 0xFFFF) | ((unsigned)u1 << 16);
            prow[q0] = (unsigned char)(t0 & 0xF0);
            prow[q0 + 1] = (unsigned char)(t1 & 0xF0);
        }
    }
}

// ---- phase 2: the un-stuffed MagSgn ring

struct HtRing {
    unsigned* w;                   // HT_RING_WORDS words of shared memory
    const unsigned char* src; int len, pos;   // MagSgn bytes (magsgn.go: forward), next source byte
    unsigned wbits, rbits;         // un-stuffed bits written / consumed so far (warp-uniform)
    unsigned last;                 // the byte before `pos` (0 at the start): a byte after 0xFF carries 7 bits

    // appends the next 512 source bytes (0xFF beyond the end, as MagSgnDecoder.readBits pads): all lanes
    __device__ __forceinline__ void refill(int lane) {
        unsigned b[16];
        const int p0 = pos + 16 * lane;
#pragma unroll
        for (int j = 0; j < 16; j++) b[j] = (p0 + j < len) ? (unsigned)src[p0 + j] : 0xFFu;
        unsigned prev = __shfl_up_sync(0xffffffffu, b[15], 1);
        if (lane == 0) prev = last;
        int nb = 0;
        unsigned seven = 0;                         // bit j: byte j carries 7 bits
        unsigned pv = prev;
#pragma unroll
        for (int j = 0; j < 16; j++) {
            if (pv == 0xFF) seven |= 1u << j;
            nb += (pv == 0xFF) ? 7 : 8;
            pv = b[j];
        }
        const int incl = ht_warp_incl_scan(nb, lane);
        const int added = __shfl_sync(0xffffffffu, incl, 31);
        // words that start at or after wbits are stale from the previous lap: clear the ones this refill touches
        const unsigned c0 = (wbits + 31) >> 5, c1 = (wbits + (unsigned)added + 31) >> 5;
        for (unsigned i = c0 + lane; i < c1; i += 32) w[i & HT_RING_MASK] = 0;
        __syncwarp();
        unsigned start = wbits + (unsigned)(incl - nb);
        unsigned wi = start >> 5;
        unsigned long long acc = 0;
        int nacc = (int)(start & 31);
#pragma unroll
        for (int j = 0; j < 16; j++) {
            const bool s7 = (seven >> j) & 1;
            acc |= (unsigned long long)(b[j] & (s7 ? 0x7Fu : 0xFFu)) << nacc;
            nacc += s7 ? 7 : 8;
            if (nacc >= 32) {
                atomicOr(&w[wi & HT_RING_MASK], (unsigned)acc);
                acc >>= 32;
                nacc -= 32;
                wi++;
            }
        }
        if (nacc > 0) atomicOr(&w[wi & HT_RING_MASK], (unsigned)acc);
        last = __shfl_sync(0xffffffffu, b[15], 31);
        pos += 512;
        wbits += (unsigned)added;
        __syncwarp();
    }
    // n bits (0..31) at absolute un-stuffed bit offset `off`
    __device__ __forceinline__ unsigned get(unsigned off, int n) const {
        const unsigned i = off >> 5;
        const unsigned lo = w[i & HT_RING_MASK], hi = w[(i + 1) & HT_RING_MASK];
        return __funnelshift_r(lo, hi, off & 31) & ((1u << n) - 1);
    }
};

// decodeOJPHSampleMS, openjph_cleanup_decoder.go:432-447 (the bits come from the ring at `off`)
__device__ __forceinline__ void ht_sample(const HtRing& ring, unsigned& off, unsigned inf, int uq, int bit, int pm1, bool exists, unsigned& val,
                                          unsigned& vn) {
    val = 0; vn = 0;
    if (!exists || (inf & (1u << (4 + bit))) == 0) return;
    const int mn = uq - (int)((inf >> (12 + bit)) & 1);
    const unsigned ms = ring.get(off, mn);
    off += (unsigned)mn;
    unsigned n = ms;                                 // already masked to mn bits
    n |= ((inf >> (8 + bit)) & 1) << mn;
    n |= 1;
    val = (ms << 31) | ((n + 2) << pm1);
    vn = n;
}

__device__ __forceinline__ int ht_final(unsigned v, unsigned shift) {   // openjph_cleanup_decoder.go:152-160
    const int mag = shift >= 32 ? 0 : (int)((v & 0x7FFFFFFFu) >> shift);
    return (v & 0x80000000u) ? -mag : mag;
}

// Kernel 2: one warp per code-block.  `tab`: the block table of the code-block interface (plane / block-major offsets, sizes).
// to_planes: 1 = write the Mallat coefficient planes (assembleSubbands fused), 0 = block-major planes.
__global__ void __launch_bounds__(128) ht_magsgn_kernel(const unsigned char* __restrict__ bytes, const HtBlock* __restrict__ descs,
                                                        const BlockEntry* __restrict__ tab, int nblocks, long long total,
                                                        long long coeffs_per_frame, const unsigned* __restrict__ sc_all, int sc_words,
                                                        int qs, int* __restrict__ out, int to_planes, int* __restrict__ status,
                                                        int warp_smem) {
    J2K_SMEM_DECL(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (wid >= total) return;
    const long long frame = wid / nblocks;
    const int bi = (int)(wid - frame * nblocks);
    const BlockEntry e = tab[bi];
    const HtBlock d = descs[wid];
    const int width = e.w, height = e.h;
    int* dst = out + frame * coeffs_per_frame + (to_planes ? e.plane_off : e.block_off);
    const long long dstride = to_planes ? e.stride : e.w;
    const int lcup = (int)d.length;
    int scup;
    int rc = ht_validate(bytes, d, scup);
    if (lcup > 0 && rc == 0) {
        unsigned char* my = smem + (size_t)warp * warp_smem;
        const int qw = (width + 1) >> 1, qh = (height + 1) >> 1;
        unsigned* vnbuf = (unsigned*)my;
        const int vnlen = qw + 2;
        HtRing ring;
        ring.w = (unsigned*)(my + warp_smem - HT_RING_WORDS * 4);
        for (int i = lane; i < 2 * vnlen; i += 32) vnbuf[i] = 0;
        for (int i = lane; i < HT_RING_WORDS; i += 32) ring.w[i] = 0;
        __syncwarp();
        const unsigned* sc = sc_all + wid * sc_words;
        ring.src = bytes + d.offset; ring.len = lcup - scup; ring.pos = 0; ring.wbits = 0; ring.rbits = 0; ring.last = 0;
        const int mmsbp2 = d.mmsb + 2, pm1 = 30 - d.mmsb - 1;
        const unsigned shift = (unsigned)(31 - (int)d.kmax);
        unsigned err = 0;
        unsigned ent_pf = lane < qw ? sc[lane] : 0;
        for (int qy = 0; qy < qh && !err; qy++) {
            const int y = 2 * qy;
            const unsigned* vold = vnbuf + (qy & 1) * vnlen;
            unsigned* vnew = vnbuf + ((qy & 1) ^ 1) * vnlen;
            unsigned carry3 = 0;                        // vn of sample 3 of the quad left of this group
            for (int g0 = 0; g0 < qw; g0 += 32) {
                const int q = g0 + lane;
                const bool act = q < qw;
                unsigned ent;
                if (g0 == 0) {   // the first group's entries were loaded a row ahead
                    ent = ent_pf;
                    ent_pf = (qy + 1 < qh && lane < qw) ? sc[(qy + 1) * qs + lane] : 0;
                } else ent = act ? sc[qy * qs + q] : 0;
                const unsigned inf = ent & 0xFFFF;
                int uq = (int)(ent >> 16);
                if (qy > 0) {
                    unsigned gamma = inf & 0xF0;
                    gamma &= gamma - 0x10;
                    const unsigned ev = act ? (vold[q] | vold[q + 1]) : 0;
                    const int emax = 31 - __clz((int)(ev | 2));
                    uq += gamma ? emax : 1;
                }
                const bool right = act && (2 * q + 1 < width);   // samples 2 and 3 exist
                int nbits = (act && uq > mmsbp2) ? (1 << 16) : 0;   // bits 16..: quads whose U_q is out of range
#pragma unroll
                for (int i = 0; i < 4; i++)
                    if (act && (i < 2 || right) && (inf & (1u << (4 + i)))) nbits += uq - (int)((inf >> (12 + i)) & 1);
                const int incl = ht_warp_incl_scan(nbits, lane);
                const int need = __shfl_sync(0xffffffffu, incl, 31);
                if (need >> 16) { err = 1; break; }
                while ((int)(ring.wbits - ring.rbits) < need) ring.refill(lane);
                unsigned off = ring.rbits + (unsigned)(incl - nbits);
                unsigned v0, v1, v2, v3, n0, n1, n2, n3;
                ht_sample(ring, off, inf, uq, 0, pm1, act, v0, n0);
                ht_sample(ring, off, inf, uq, 1, pm1, act, v1, n1);
                ht_sample(ring, off, inf, uq, 2, pm1, right, v2, n2);
                ht_sample(ring, off, inf, uq, 3, pm1, right, v3, n3);
                (void)n0; (void)n2;
                ring.rbits += (unsigned)need;
                __syncwarp();   // every lane has read its bits before a later refill clears ring words
                if (act) {
                    int* r0 = dst + (long long)y * dstride + 2 * q;
                    r0[0] = ht_final(v0, shift);
                    if (right) r0[1] = ht_final(v2, shift);
                    if (y + 1 < height) {
                        r0[dstride] = ht_final(v1, shift);
                        if (right) r0[dstride + 1] = ht_final(v3, shift);
                    }
                }
                // exponent-predictor state for the row below: vnew[q] = vn3(q - 1) | vn1(q), vnew[qw] = vn3(qw - 1)
                unsigned left3 = __shfl_up_sync(0xffffffffu, n3, 1);
                if (lane == 0) left3 = carry3;
                if (act) vnew[q] = left3 | n1;
                if (act && q == qw - 1) vnew[qw] = n3;
                carry3 = __shfl_sync(0xffffffffu, n3, 31);
            }
            __syncwarp();
        }
        if (err) rc = -3;
    }
    if (lcup == 0 || rc != 0) {   // no data, or an error: the block reads as zeros
        __syncwarp();
        const int n = width * height;
        for (int i = lane; i < n; i += 32) {
            const int y = i / width, x = i - y * width;
            dst[(long long)y * dstride + x] = 0;
        }
    }
    if (status && lane == 0) status[wid] = rc;
}

}  // namespace j2k
