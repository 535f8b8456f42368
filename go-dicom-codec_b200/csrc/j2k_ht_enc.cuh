// HTJ2K cleanup-pass block ENCODER on the device (encode side).
//
// Replaces, per code-block, HTEncoder.Encode -> encodeOpenJPHCleanup
// (/root/reference/jpeg2000/htj2k/encoder.go:54-68, openjph_cleanup_encoder.go:200-252) and, in the same pass, the sub-band
// extraction + partitionIntoCodeBlocks copy of buildTilePacketEncoder (jpeg2000/encoder.go:2424-2431): the quads are read
// straight from the Mallat coefficient planes the forward ring kernel wrote.  Byte-identical to the reference encoder (and,
// through it, to OpenJPH: htj2k/go_byte_parity_test.go).  What crosses PCIe on the encode side is then the compressed
// cleanup segments + 16 bytes per block instead of 4 bytes per coefficient.
//
// Unlike decoding, everything but the byte-stuffing and the adaptive MEL state is parallel over the quads of a block: the
// significance patterns, exponents, contexts and U_q of a quad depend only on its own samples and on the quad row above.
// Three kernels (+ a scan of the block lengths):
//   1. ht_enc_quads_kernel, one WARP per code-block, one lane per quad of a quad row: prepareOJPHQuad, the context / kappa /
//      U_q / EMB rules and the VLC table lookup of encodeOJPHInitialRows / encodeOJPHSubsequentRows (:254-380); the MagSgn
//      bits, the VLC + U-VLC bits and the MEL events of the row are appended to three UN-STUFFED bit streams at offsets a warp
//      prefix sum gives each lane (staged in shared memory, flushed to the block's scratch slot as whole words).
//   2. ht_enc_pack_kernel, one THREAD per code-block (32 blocks per warp in lock-step): the sequential parts -- MagSgn byte
//      stuffing (ojphMSWriter, :114-166), VLC byte stuffing (ojphVLCWriter, :64-112), MEL coding (ojphMELWriter, :8-62) and the
//      MEL / VLC termination with its byte fusion (:522-545).
//   3. ht_enc_scan_kernel: exclusive prefix sum of the block lengths -> offsets in the compact stream.
//   4. ht_enc_compact_kernel, one warp per block: MagSgn | MEL | reversed VLC copied to the compact stream, the Scup locator
//      (encoder.go:84-90) patched in, the block's j2k_ht_cblk record written.
#pragma once

namespace j2k {

#define J2K_HT_TABLE static __device__ const
#include "j2k_ht_enc_tables.inc"
#undef J2K_HT_TABLE

// Per-block scratch slot (bytes, all 16-byte aligned): un-stuffed MagSgn words | un-stuffed VLC words | MEL event words |
// final MagSgn bytes | final MEL bytes | final VLC bytes (write order).  Sized for the largest Kmax of the call.
struct HtEncLayout {
    int ms_bits_off, vlc_bits_off, mel_ev_off, ms_out_off, mel_out_off, vlc_out_off, slot_bytes;
};
__host__ __device__ inline int ht_align16(int v) { return (v + 15) & ~15; }
__host__ inline HtEncLayout ht_enc_layout(int cbw, int cbh, int kmax_max) {
    const int samples = cbw * cbh, pairs = ((cbw + 3) / 4) * ((cbh + 1) / 2);
    const int ms_bits = samples * (kmax_max + 1), vlc_bits = pairs * 40, mel_ev = pairs * 3;
    HtEncLayout L;
    int o = 0;
    L.ms_bits_off = o; o += ht_align16(ms_bits / 8 + 8);
    L.vlc_bits_off = o; o += ht_align16(vlc_bits / 8 + 8);
    L.mel_ev_off = o; o += ht_align16(mel_ev / 8 + 48);
    L.ms_out_off = o; o += ht_align16(ms_bits / 7 + 32);
    L.mel_out_off = o; o += ht_align16(mel_ev * 6 / 7 + 32);
    L.vlc_out_off = o; o += ht_align16(vlc_bits / 7 + 32);
    // (vector reads run at most two vectors past the end of a stream: the slack of every area covers that)
    L.slot_bytes = o;
    return L;
}

struct HtEncInfo {       // per block, between the kernels
    unsigned ms_nbits, vlc_nbits, mel_nev, nonempty;   // written by kernel 1
    unsigned ms_n, mel_n, vlc_n, total;                // written by kernel 2 (bytes)
};

// Appends per-lane bit strings to an un-stuffed stream: staged in `buf` (shared, words), flushed to `gw` as whole words.
struct HtBitOut {
    unsigned* buf; unsigned* gw; unsigned flushed, pend, total;
    __device__ __forceinline__ void init(unsigned* b, int words, unsigned* g, int lane) {
        buf = b; gw = g; flushed = 0; pend = 0; total = 0;
        for (int i = lane; i < words; i += 32) buf[i] = 0;
    }
    __device__ __forceinline__ void put(unsigned& pos, unsigned v, int n) {   // n <= 32 bits of v at bit position pos
        if (n <= 0) return;
        const unsigned w = pos >> 5, o = pos & 31;
        atomicOr(&buf[w], v << o);
        if (o + n > 32) atomicOr(&buf[w + 1], v >> (32 - o));
        pos += n;
    }
    // every lane has deposited its bits at pend + its exclusive offset; `sum` = bits of all lanes (warp-uniform)
    __device__ __forceinline__ void flush(int sum, int lane) {
        __syncwarp();
        const unsigned tot = pend + (unsigned)sum, nfull = tot >> 5;
        for (unsigned i = lane; i < nfull; i += 32) gw[flushed + i] = buf[i];
        __syncwarp();
        const unsigned part = buf[nfull];
        __syncwarp();
        for (unsigned i = lane; i <= nfull; i += 32) buf[i] = 0;
        __syncwarp();
        if (lane == 0) buf[0] = part;
        __syncwarp();
        flushed += nfull; pend = tot & 31; total += (unsigned)sum;
    }
    __device__ __forceinline__ void finish(int lane) {
        if (pend && lane == 0) gw[flushed] = buf[0];
    }
};

// ojphUVLC (:168-198): prefix and suffix of one u value (the `ext` field is never written by the reference, :485-520)
__device__ __forceinline__ void ht_uvlc_code(int code, int& pre, int& pre_len, int& suf, int& suf_len) {
    pre = 0; pre_len = 0; suf = 0; suf_len = 0;
    if (code <= 0) return;
    if (code == 1) { pre = 1; pre_len = 1; }
    else if (code == 2) { pre = 2; pre_len = 2; }
    else if (code <= 4) { pre = 4; pre_len = 3; suf = code - 3; suf_len = 1; }
    else if (code <= 32) { pre_len = 3; suf = code - 5; suf_len = 5; }
    else { pre_len = 3; suf = 28 + ((code - 33) % 4); suf_len = 5; }
}

#define HT_ENC_MS_WORDS 128   // per-row staging: 32 quads x 124 bits + 31 pending
#define HT_ENC_VLC_WORDS 24
#define HT_ENC_MEL_WORDS 4
__host__ __device__ inline int ht_enc_warp_smem(int cbw) {
    return (HT_ENC_MS_WORDS + HT_ENC_VLC_WORDS + HT_ENC_MEL_WORDS) * 4 + 2 * ht_align16((cbw + 1) / 2 + 3) * 2;
}

// Kernel 1.  `tab`: block table (plane offsets, sizes, component); kmax: [components][3 * levels + 1] band precisions with
// `band_of` giving each block's (component, band index) through tab[].comp and bidx[].
__global__ void __launch_bounds__(128) ht_enc_quads_kernel(const int* __restrict__ coeffs, long long coeffs_per_frame,
                                                           const BlockEntry* __restrict__ tab, const unsigned char* __restrict__ blk_kmax,
                                                           int nblocks, long long total, unsigned char* __restrict__ slots, HtEncLayout L,
                                                           HtEncInfo* __restrict__ info, int warp_smem) {
    J2K_SMEM_DECL(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (wid >= total) return;
    const long long frame = wid / nblocks;
    const int bi = (int)(wid - frame * nblocks);
    const BlockEntry e = tab[bi];
    const int width = e.w, height = e.h, kmax = blk_kmax[bi];
    const int* src = coeffs + frame * coeffs_per_frame + e.plane_off;
    const long long sstride = e.stride;
    unsigned char* slot = slots + (size_t)wid * L.slot_bytes;
    unsigned char* my = smem + (size_t)warp * warp_smem;
    HtBitOut ms, vlc, mel;
    ms.init((unsigned*)my, HT_ENC_MS_WORDS, (unsigned*)(slot + L.ms_bits_off), lane);
    vlc.init((unsigned*)my + HT_ENC_MS_WORDS, HT_ENC_VLC_WORDS, (unsigned*)(slot + L.vlc_bits_off), lane);
    mel.init((unsigned*)my + HT_ENC_MS_WORDS + HT_ENC_VLC_WORDS, HT_ENC_MEL_WORDS, (unsigned*)(slot + L.mel_ev_off), lane);
    const int qw = (width + 1) >> 1, qh = (height + 1) >> 1;
    const int alen = ht_align16(qw + 3);
    unsigned char* cxbuf = my + (HT_ENC_MS_WORDS + HT_ENC_VLC_WORDS + HT_ENC_MEL_WORDS) * 4;   // two rows: A[j] = rho3(j-1) | rho1(j)
    unsigned char* ebuf = cxbuf + 2 * alen;                                                    // two rows: E[j] = max(E3(j-1), E1(j))
    for (int i = lane; i < 2 * alen; i += 32) { cxbuf[i] = 0; ebuf[i] = 0; }
    __syncwarp();
    const unsigned shift = (unsigned)(31 - kmax);   // = p: missing_msbs = kmax - 1 (:226-227)
    unsigned any = 0;
    for (int qy = 0; qy < qh; qy++) {
        const int y = 2 * qy;
        const unsigned char* cxo = cxbuf + (qy & 1) * alen; unsigned char* cxn = cxbuf + ((qy & 1) ^ 1) * alen;
        const unsigned char* eo = ebuf + (qy & 1) * alen; unsigned char* en = ebuf + ((qy & 1) ^ 1) * alen;
        int carry_rho = 0, carry_e3 = 0;   // rho and E3 of the quad left of this group
        for (int g0 = 0; g0 < qw; g0 += 32) {
            const int q = g0 + lane;
            const bool act = q < qw;
            // prepareOJPHQuad / prepareOJPHSample (:382-413)
            int rho = 0, eqmax = 0, eq[4] = {0, 0, 0, 0};
            unsigned s[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int xx = 2 * q + (i >> 1), yy = y + (i & 1);
                if (act && xx < width && yy < height) {
                    const int v = src[(long long)yy * sstride + xx];
                    const unsigned sign = v < 0 ? 0x80000000u : 0u;
                    const unsigned mag = v < 0 ? 0u - (unsigned)v : (unsigned)v;
                    const unsigned val0 = mag << shift;
                    any |= val0;
                    const unsigned t = sign | val0;
                    unsigned val = ((t + t) >> shift) & ~1u;
                    if (val) {
                        rho |= 1 << i;
                        val--;
                        eq[i] = 32 - __clz((int)val);
                        eqmax = max(eqmax, eq[i]);
                        val--;
                        s[i] = val + (t >> 31);
                    }
                }
            }
            // neighbours: the quad to the left (this row), the context / exponent rows of the quad row above
            int rho_l = __shfl_up_sync(0xffffffffu, rho, 1), e3_l = __shfl_up_sync(0xffffffffu, eq[3], 1);
            if (lane == 0) { rho_l = carry_rho; e3_l = carry_e3; }
            int cq, kappa = 1;
            if (qy == 0) cq = (rho_l >> 1) | (rho_l & 1);
            else {
                const int a0 = act ? cxo[q] : 0, a1 = act ? cxo[q + 1] : 0;
                cq = a0 + (a1 << 2) + (((rho_l & 4) >> 1) | ((rho_l & 8) >> 2));
                if (rho & (rho - 1)) kappa = max(1, max(act ? (int)eo[q] : 0, act ? (int)eo[q + 1] : 0) - 1);
            }
            if (q == 0 && qy == 0) cq = 0;
            const int uq = max(eqmax, kappa), u = uq - kappa;
            int eps = 0;
            if (u > 0) {
#pragma unroll
                for (int i = 0; i < 4; i++) if (eq[i] == eqmax) eps |= 1 << i;
            }
            int tuple = 0;
            if (act && !(rho == 0 && cq == 0)) tuple = (qy == 0 ? HT_ENC_TBL0 : HT_ENC_TBL1)[(cq << 8) | (rho << 4) | eps];
            // rows of context / exponent state for the quad row below
            if (act) {
                cxn[q] = (unsigned char)(((rho_l & 8) >> 3) | ((rho & 2) >> 1));
                en[q] = (unsigned char)max(e3_l, eq[1]);
                if (q == qw - 1) { cxn[qw] = (unsigned char)((rho & 8) >> 3); en[qw] = (unsigned char)eq[3]; cxn[qw + 1] = 0; en[qw + 1] = 0; }
            }
            carry_rho = __shfl_sync(0xffffffffu, rho, 31);
            carry_e3 = __shfl_sync(0xffffffffu, eq[3], 31);
            // ---- MagSgn bits of the quad (ojphEncodeMagSgn, :472-483)
            int mlen[4], nms = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                mlen[i] = (rho & (1 << i)) ? max(uq - ((tuple >> i) & 1), 0) : 0;
                nms += mlen[i];
            }
            // ---- VLC + U-VLC bits and MEL events of the pair: the even lane owns them
            const int t1 = __shfl_down_sync(0xffffffffu, tuple, 1), u1 = __shfl_down_sync(0xffffffffu, u, 1);
            const int rho1 = __shfl_down_sync(0xffffffffu, rho, 1), cq1 = __shfl_down_sync(0xffffffffu, cq, 1);
            const bool owner = act && !(lane & 1);
            const bool second = owner && (q + 1 < qw);
            unsigned long long vbits = 0; int nv = 0;
            unsigned mbits = 0; int nm = 0;
            if (owner) {
                const int uu1 = second ? u1 : 0;
                auto add = [&](int v, int n) { vbits |= (unsigned long long)(unsigned)v << nv; nv += n; };
                add(tuple >> 8, (tuple >> 4) & 7);
                if (cq == 0) { mbits |= (unsigned)(rho != 0) << nm; nm++; }
                if (second) {
                    add(t1 >> 8, (t1 >> 4) & 7);
                    if (cq1 == 0) { mbits |= (unsigned)(rho1 != 0) << nm; nm++; }
                }
                int p0, l0, s0, sl0, p1, l1, s1, sl1;
                if (qy == 0) {   // ojphEncodeInitialUVLC (:485-511)
                    if (u > 0 && uu1 > 0) { mbits |= (unsigned)(min(u, uu1) > 2) << nm; nm++; }
                    if (u > 2 && uu1 > 2) {
                        ht_uvlc_code(u - 2, p0, l0, s0, sl0); ht_uvlc_code(uu1 - 2, p1, l1, s1, sl1);
                        add(p0, l0); add(p1, l1); add(s0, sl0); add(s1, sl1);
                    } else if (u > 2 && uu1 > 0) {
                        ht_uvlc_code(u, p0, l0, s0, sl0);
                        add(p0, l0); add(uu1 - 1, 1); add(s0, sl0);
                    } else {
                        ht_uvlc_code(u, p0, l0, s0, sl0); ht_uvlc_code(uu1, p1, l1, s1, sl1);
                        add(p0, l0); add(p1, l1); add(s0, sl0); add(s1, sl1);
                    }
                } else {         // ojphEncodeNonInitialUVLC (:513-520)
                    ht_uvlc_code(u, p0, l0, s0, sl0); ht_uvlc_code(uu1, p1, l1, s1, sl1);
                    add(p0, l0); add(p1, l1); add(s0, sl0); add(s1, sl1);
                }
            }
            // ---- append: offsets by prefix sums (MagSgn | VLC << 12 | MEL << 24 in one scan)
            const int packed = nms | (nv << 12) | (nm << 24);
            const int incl = ht_warp_incl_scan(packed, lane);
            const int sum = __shfl_sync(0xffffffffu, incl, 31);
            const int excl = incl - packed;
            unsigned pos = ms.pend + (unsigned)(excl & 0xFFF);
#pragma unroll
            for (int i = 0; i < 4; i++) ms.put(pos, mlen[i] >= 32 ? s[i] : (s[i] & ((1u << mlen[i]) - 1)), mlen[i]);
            pos = vlc.pend + (unsigned)((excl >> 12) & 0xFFF);
            vlc.put(pos, (unsigned)vbits, min(nv, 32));
            if (nv > 32) vlc.put(pos, (unsigned)(vbits >> 32), nv - 32);
            pos = mel.pend + (unsigned)(excl >> 24);
            mel.put(pos, mbits, nm);
            ms.flush(sum & 0xFFF, lane);
            vlc.flush((sum >> 12) & 0xFFF, lane);
            mel.flush(sum >> 24, lane);
        }
        __syncwarp();
    }
    ms.finish(lane); vlc.finish(lane); mel.finish(lane);
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) any |= __shfl_xor_sync(0xffffffffu, any, d);
    if (lane == 0) {
        HtEncInfo I;
        I.ms_nbits = ms.total; I.vlc_nbits = vlc.total; I.mel_nev = mel.total;
        I.nonempty = any >= (1u << shift) ? 1u : 0u;   // "if maxVal < 1 << shift: return nil, nil" (:218-220)
        I.ms_n = I.mel_n = I.vlc_n = I.total = 0;
        info[wid] = I;
    }
}

// Sequential reader of an un-stuffed bit stream (words in global memory).  A thread of the pack kernel runs alone on its block,
// so its speed is the latency of its own loads: 16 bytes per load, the next 16 requested one step ahead.  (The stream areas are
// 16-byte aligned and followed by other areas of the slot, so reading a vector or two past the last word stays inside it.)
// (Measured: the kernel takes the time of ONE thread whatever the number of blocks -- 1.3 ms for 16 k as for 33 k blocks of C2
// frames; ncu puts 51 % of the samples on the move that takes `nxt` into `cur`, i.e. on load latency.  An L1 prefetch three
// lines ahead changed nothing; what pays is more blocks per launch, which is why j2k_forward_ht uses 64-Msample sub-batches.)
struct HtBitIn {
    const uint4* w; uint4 cur, nxt; int wi, vi; unsigned long long acc; int nacc; unsigned left;
    __device__ __forceinline__ void init(const void* p, unsigned nbits) {
        w = (const uint4*)p; acc = 0; nacc = 0; left = nbits; wi = 0; vi = 0;
        cur = w[0]; nxt = w[1];
    }
    __device__ __forceinline__ unsigned word() {
        const unsigned v = cur.x;
        cur.x = cur.y; cur.y = cur.z; cur.z = cur.w;
        if (++wi == 4) { cur = nxt; vi++; nxt = w[vi + 1]; wi = 0; }
        return v;
    }
    __device__ __forceinline__ unsigned take(int n) {   // n <= 8, n <= left
        if (nacc < n) { acc |= (unsigned long long)word() << nacc; nacc += 32; }
        const unsigned v = (unsigned)acc & ((1u << n) - 1);
        acc >>= n; nacc -= n; left -= (unsigned)n;
        return v;
    }
    __device__ __forceinline__ unsigned peek32() {      // the next 32 bits (left >= 32)
        if (nacc < 32) { acc |= (unsigned long long)word() << nacc; nacc += 32; }
        return (unsigned)acc;
    }
    __device__ __forceinline__ void skip32() { acc >>= 32; nacc -= 32; left -= 32; }
};

// Byte writer: 16 bytes per store (the output areas are 16-byte aligned, with room for a last full vector)
struct HtByteOut {
    uint4* p; unsigned n; unsigned long long lo, hi;
    __device__ __forceinline__ void init(void* q) { p = (uint4*)q; n = 0; lo = 0; hi = 0; }
    __device__ __forceinline__ void put(unsigned b) {
        const unsigned k = n & 15;
        if (k < 8) lo |= (unsigned long long)(b & 0xFF) << (8 * k);
        else hi |= (unsigned long long)(b & 0xFF) << (8 * (k - 8));
        if ((++n & 15) == 0) { *p++ = make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)hi, (unsigned)(hi >> 32)); lo = 0; hi = 0; }
    }
    __device__ __forceinline__ void put4(unsigned w4) {   // four bytes, least significant first
        const unsigned k = n & 15;
        unsigned long long over = 0;
        if (k < 8) {
            lo |= (unsigned long long)w4 << (8 * k);
            if (k > 4) hi |= (unsigned long long)w4 >> (64 - 8 * k);
        } else {
            hi |= (unsigned long long)w4 << (8 * (k - 8));
            if (k > 12) over = (unsigned long long)w4 >> (8 * (16 - k));
        }
        n += 4;
        if (k >= 12) {
            *p++ = make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)hi, (unsigned)(hi >> 32));
            lo = over; hi = 0;
        }
    }
    __device__ __forceinline__ void finish() {
        if (n & 15) *p = make_uint4((unsigned)lo, (unsigned)(lo >> 32), (unsigned)hi, (unsigned)(hi >> 32));
    }
};

// Kernel 2: one thread per block.
__global__ void __launch_bounds__(32) ht_enc_pack_kernel(long long total, unsigned char* __restrict__ slots, HtEncLayout L,
                                                         HtEncInfo* __restrict__ info) {
    const long long wid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (wid >= total) return;
    HtEncInfo I = info[wid];
    if (!I.nonempty) return;
    unsigned char* slot = slots + (size_t)wid * L.slot_bytes;
    // ---- MagSgn: ojphMSWriter.encode / terminate (:114-166), fed 8 (or 7) bits at a time
    unsigned ms_n = 0;
    {
        HtBitIn in; in.init(slot + L.ms_bits_off, I.ms_nbits);
        HtByteOut out; out.init(slot + L.ms_out_off);
        int max_bits = 8;
        while (in.left >= (unsigned)max_bits) {
            // four bytes at a time while none of them is 0xFF (only a 0xFF byte changes the width of the byte after it):
            // a lone thread pays ~200 cycles of dependent instructions per byte, and 98 % of the words take this path
            if (max_bits == 8 && in.left >= 32) {
                const unsigned w4 = in.peek32(), nw = ~w4;
                if (((nw - 0x01010101u) & w4 & 0x80808080u) == 0) { out.put4(w4); in.skip32(); continue; }
            }
            const unsigned b = in.take(max_bits);
            out.put(b);
            max_bits = b == 0xFF ? 7 : 8;
        }
        const int used = (int)in.left;
        ms_n = out.n;
        if (used != 0) {
            unsigned tmp = in.take(used);
            const int t = max_bits - used;
            tmp |= (0xFFu & ((1u << t) - 1)) << used;
            if ((tmp & 0xFF) != 0xFF) { out.put(tmp); ms_n++; }
        } else if (max_bits == 7 && ms_n > 0) ms_n--;
        out.finish();
    }
    // ---- VLC: ojphVLCWriter.encode (:74-102) bit-serially (the stuffing decision only looks at the accumulated byte)
    int vtmp = 0xF, vused = 4; bool vlast = true;
    HtByteOut vout; vout.init(slot + L.vlc_out_off);
    vout.put(0xFF);
    {
        HtBitIn in; in.init(slot + L.vlc_bits_off, I.vlc_nbits);
        while (in.left > 0) {
            int avail = 8 - (vlast ? 1 : 0) - vused;
            const int t = (int)min((unsigned)avail, in.left);
            if (t > 0) { vtmp |= (int)in.take(t) << vused; vused += t; avail -= t; }
            if (avail == 0) {
                if (vlast && vtmp != 0x7F) { vlast = false; continue; }
                vout.put((unsigned)vtmp);
                vlast = vtmp > 0x8F;
                vtmp = 0; vused = 0;
            }
        }
    }
    // ---- MEL: ojphMELWriter (:8-62)
    int mtmp = 0, mrem = 8, mrun = 0, mk = 0, mthr = 1;
    HtByteOut mout; mout.init(slot + L.mel_out_off);
    auto emit = [&](int v) {
        mtmp = (mtmp << 1) | (v & 1);
        if (--mrem == 0) {
            mout.put((unsigned)mtmp);
            mrem = mtmp == 0xFF ? 7 : 8;
            mtmp = 0;
        }
    };
    {
        HtBitIn in; in.init(slot + L.mel_ev_off, I.mel_nev);
        while (in.left > 0) {
            const int bit = (int)in.take(1);
            const int ev = (int)((0x5433222111000ULL >> (4 * mk)) & 0xF);
            if (!bit) {
                if (++mrun >= mthr) {
                    emit(1);
                    mrun = 0;
                    if (mk < 12) mk++;
                    mthr = 1 << (int)((0x5433222111000ULL >> (4 * mk)) & 0xF);
                }
            } else {
                emit(0);
                for (int t = ev; t > 0;) { t--; emit((mrun >> t) & 1); }
                mrun = 0;
                if (mk > 0) mk--;
                mthr = 1 << (int)((0x5433222111000ULL >> (4 * mk)) & 0xF);
            }
        }
    }
    // ---- terminateOJPHMELVLC (:522-545)
    if (mrun > 0) emit(1);
    mtmp <<= mrem;
    const int mel_mask = (0xFF << mrem) & 0xFF;
    const int vlc_mask = vused > 0 ? 0xFF >> (8 - vused) : 0;
    if ((mel_mask | vlc_mask) != 0) {
        const int fuse = mtmp | vtmp;
        if ((((fuse ^ mtmp) & mel_mask) | ((fuse ^ vtmp) & vlc_mask)) == 0 && fuse != 0xFF && vout.n > 1) {
            mout.put((unsigned)fuse);
        } else {
            mout.put((unsigned)mtmp);
            vout.put((unsigned)vtmp);
        }
    }
    vout.finish(); mout.finish();
    I.ms_n = ms_n; I.mel_n = mout.n; I.vlc_n = vout.n; I.total = ms_n + mout.n + vout.n;
    info[wid] = I;
}

// Kernel 3: exclusive prefix sum of info[].total over the blocks of the launch (one CTA of 1024 threads); offsets[total] = sum.
__global__ void __launch_bounds__(1024) ht_enc_scan_kernel(const HtEncInfo* __restrict__ info, long long total,
                                                           unsigned long long* __restrict__ offsets) {
#ifdef J2K_EMU
    unsigned long long run = 0;
    if (threadIdx.x == 0) {
        for (long long i = 0; i < total; i++) { offsets[i] = run; run += info[i].nonempty ? info[i].total : 0; }
        offsets[total] = run;
    }
#else
    __shared__ unsigned long long part[32];
    __shared__ unsigned long long base;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    for (long long i0 = 0; i0 < total; i0 += 1024) {
        const long long i = i0 + threadIdx.x;
        const unsigned long long v = (i < total && info[i].nonempty) ? info[i].total : 0;
        unsigned long long s = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += t;
        }
        if (lane == 31) part[warp] = s;
        __syncthreads();
        if (warp == 0) {
            unsigned long long p = part[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long t = __shfl_up_sync(0xffffffffu, p, d);
                if (lane >= d) p += t;
            }
            part[lane] = p;
        }
        __syncthreads();
        const unsigned long long off = base + (warp ? part[warp - 1] : 0) + s - v;
        if (i < total) offsets[i] = off;
        __syncthreads();
        if (threadIdx.x == 1023) base = off + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[total] = base;
#endif
}

// Kernel 4: one warp per block.
__global__ void __launch_bounds__(128) ht_enc_compact_kernel(long long total, const unsigned char* __restrict__ slots, HtEncLayout L,
                                                             const HtEncInfo* __restrict__ info, const unsigned long long* __restrict__ offsets,
                                                             const unsigned char* __restrict__ blk_kmax, int nblocks,
                                                             unsigned char* __restrict__ out, unsigned long long cap,
                                                             HtBlock* __restrict__ recs) {
    const int lane = threadIdx.x & 31;
    const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (wid >= total) return;
    const HtEncInfo I = info[wid];
    const int kmax = blk_kmax[(int)(wid % nblocks)];
    const unsigned long long off = offsets[wid];
    const unsigned n = I.nonempty ? I.total : 0;
    if (lane == 0) {
        HtBlock r;
        r.offset = off; r.length = n; r.kmax = (unsigned char)kmax; r.mmsb = (unsigned char)(kmax - 1); r.reserved = 0;
        recs[wid] = r;
    }
    if (!n || off + n > cap) return;   // a stream that does not fit is reported through offsets[total], never written past `cap`
    const unsigned char* slot = slots + (size_t)wid * L.slot_bytes;
    unsigned char* dst = out + off;
    const unsigned char* a = slot + L.ms_out_off;
    for (unsigned i = lane; i < I.ms_n; i += 32) dst[i] = a[i];
    a = slot + L.mel_out_off;
    for (unsigned i = lane; i < I.mel_n; i += 32) dst[I.ms_n + i] = a[i];
    // ojphVLCWriter.bytes() (:104-112): newest byte first, the Scup placeholder (vlc[0]) last; then writeScupLocator
    a = slot + L.vlc_out_off;
    const unsigned scup = I.mel_n + I.vlc_n;
    for (unsigned i = lane; i < I.vlc_n; i += 32) {
        unsigned b = a[I.vlc_n - 1 - i];
        const unsigned at = I.ms_n + I.mel_n + i;
        if (at == n - 1) b = scup >> 4;
        dst[at] = (unsigned char)b;
    }
    __syncwarp();
    if (lane == 0 && n >= 2) dst[n - 2] = (unsigned char)((dst[n - 2] & 0xF0) | (scup & 0x0F));
}

}  // namespace j2k
