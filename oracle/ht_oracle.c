/* CPU ORACLE of the HTJ2K cleanup-pass block decoder -- TEST INFRASTRUCTURE, never shipped, never on the product path.
 *
 * A loop-for-loop C restatement of the reference's production HT block decoder
 *   /root/reference/jpeg2000/htj2k/openjph_cleanup_decoder.go  (decodeOpenJPHCleanup and its readers)
 *   /root/reference/jpeg2000/htj2k/vlc_reverse_decoder.go      (reverseBitReader)
 *   /root/reference/jpeg2000/htj2k/magsgn.go:160-204           (MagSgnDecoder.readBits)
 *   /root/reference/jpeg2000/htj2k/vlc_tables.go:876-925       (InitVLCTables)
 *   /root/reference/jpeg2000/htj2k/uvlc_tables.go:40-143       (generateUVLCTables)
 *   /root/reference/jpeg2000/htj2k/decoder.go:60-70            (parseStandardSegments)
 * each function citing the lines it follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may load it.
 *
 * PINNED: tests/test_ht_oracle.py decodes the 14 OpenJPH codestreams of test-data/htj2k/interop (copied to
 * tests/golden/htj2k_interop) through this decoder + the sample-domain oracle and compares with input.raw, which is what the
 * reference's own interop test asserts (htj2k/interop_manifest_test.go:43-74).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define EXPORT __attribute__((visibility("default")))

#include "ht_vlc_src.inc"

static uint16_t VLC0[1024], VLC1[1024], UVLC0[320], UVLC1[256];
static int tables_ready;

/* vlc_tables.go:876-925 */
static void init_vlc(uint16_t* tab, const unsigned char (*src)[7], int n) {
    for (int i = 0; i < 1024; i++) {
        uint8_t cwd = (uint8_t)(i & 0x7F), cq = (uint8_t)(i >> 7);
        tab[i] = 0;
        for (int j = 0; j < n; j++) {
            const unsigned char* e = src[j];
            if (e[0] == cq) {
                uint8_t mask = (uint8_t)((1u << e[6]) - 1);
                if (e[5] == (cwd & mask)) {
                    tab[i] = (uint16_t)((e[3] << 12) | (e[4] << 8) | (e[1] << 4) | (e[2] << 3) | e[6]);
                    break;
                }
            }
        }
    }
}

/* uvlc_tables.go:40-143 */
static void init_uvlc(void) {
    const uint8_t dec[8] = {3 | (5 << 2) | (5 << 5), 1 | (0 << 2) | (1 << 5), 2 | (0 << 2) | (2 << 5), 1 | (0 << 2) | (1 << 5),
                            3 | (1 << 2) | (3 << 5), 1 | (0 << 2) | (1 << 5), 2 | (0 << 2) | (2 << 5), 1 | (0 << 2) | (1 << 5)};
    for (int i = 0; i < 320; i++) {
        int mode = i >> 6, vlc = i & 0x3F;
        int lp, ls, u0suf, u0, u1;
        uint8_t d, d0, d1;
        switch (mode) {
        case 0: UVLC0[i] = 0; break;
        case 1: case 2:
            d = dec[vlc & 7]; lp = d & 3; ls = (d >> 2) & 7; u0suf = ls; u0 = d >> 5; u1 = 0;
            if (mode == 2) { u0suf = 0; u0 = 0; u1 = d >> 5; }
            UVLC0[i] = (uint16_t)(lp | (ls << 3) | (u0suf << 7) | (u0 << 10) | (u1 << 13));
            break;
        case 3:
            d0 = dec[vlc & 7]; vlc >>= d0 & 3; d1 = dec[vlc & 7];
            if ((d0 & 3) == 3) {
                lp = (d0 & 3) + 1; u0suf = (d0 >> 2) & 7; ls = u0suf; u0 = d0 >> 5; u1 = (vlc & 1) + 1;
            } else {
                lp = (d0 & 3) + (d1 & 3); u0suf = (d0 >> 2) & 7; ls = u0suf + ((d1 >> 2) & 7); u0 = d0 >> 5; u1 = d1 >> 5;
            }
            UVLC0[i] = (uint16_t)(lp | (ls << 3) | (u0suf << 7) | (u0 << 10) | (u1 << 13));
            break;
        case 4:
            d0 = dec[vlc & 7]; vlc >>= d0 & 3; d1 = dec[vlc & 7];
            lp = (d0 & 3) + (d1 & 3); u0suf = (d0 >> 2) & 7; ls = u0suf + ((d1 >> 2) & 7); u0 = (d0 >> 5) + 2; u1 = (d1 >> 5) + 2;
            UVLC0[i] = (uint16_t)(lp | (ls << 3) | (u0suf << 7) | (u0 << 10) | (u1 << 13));
            break;
        }
    }
    for (int i = 0; i < 256; i++) {
        int mode = i >> 6, vlc = i & 0x3F;
        int lp, ls, u0suf, u0, u1;
        uint8_t d, d0, d1;
        switch (mode) {
        case 0: UVLC1[i] = 0; break;
        case 1: case 2:
            d = dec[vlc & 7]; lp = d & 3; ls = (d >> 2) & 7; u0suf = ls; u0 = d >> 5; u1 = 0;
            if (mode == 2) { u0suf = 0; u0 = 0; u1 = d >> 5; }
            UVLC1[i] = (uint16_t)(lp | (ls << 3) | (u0suf << 7) | (u0 << 10) | (u1 << 13));
            break;
        case 3:
            d0 = dec[vlc & 7]; vlc >>= d0 & 3; d1 = dec[vlc & 7];
            lp = (d0 & 3) + (d1 & 3); u0suf = (d0 >> 2) & 7; ls = u0suf + ((d1 >> 2) & 7); u0 = d0 >> 5; u1 = d1 >> 5;
            UVLC1[i] = (uint16_t)(lp | (ls << 3) | (u0suf << 7) | (u0 << 10) | (u1 << 13));
            break;
        }
    }
}

static void init_tables(void) {
    if (tables_ready) return;
    init_vlc(VLC0, HT_VLC_SRC0, (int)(sizeof(HT_VLC_SRC0) / 7));
    init_vlc(VLC1, HT_VLC_SRC1, (int)(sizeof(HT_VLC_SRC1) / 7));
    init_uvlc();
    tables_ready = 1;
}

/* which: 0 VLCLookupTable0, 1 VLCLookupTable1, 2 UVLCTbl0, 3 UVLCTbl1; returns the entry count */
EXPORT int orc_ht_table(int which, uint16_t* out) {
    init_tables();
    const uint16_t* t = which == 0 ? VLC0 : which == 1 ? VLC1 : which == 2 ? UVLC0 : UVLC1;
    int n = which < 2 ? 1024 : which == 2 ? 320 : 256;
    memcpy(out, t, (size_t)n * 2);
    return n;
}

/* mel_spec.go:8-22 */
static const int MEL_E[13] = {0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 4, 5};

/* openjph_cleanup_decoder.go:8-101 (ojphMELReader).  bitBuf is a FIFO of single bits there; here a shift register. */
typedef struct {
    const uint8_t* data; int len; int pos; int size; int unstuff; int k; int num_runs; uint64_t runs;
    uint32_t bits; int nbits;
} mel_t;

static int mel_read_bit(mel_t* m) { /* :74-101 */
    while (m->nbits == 0) {
        if (m->size <= 0) return 1;
        uint8_t d = 0xFF;
        if (m->pos < m->len) {
            d = m->data[m->pos++];
            if (m->size == 1) d |= 0x0F;
            m->size--;
        }
        int valid = m->unstuff ? 7 : 8;
        m->bits = (uint32_t)d & ((1u << valid) - 1);
        m->nbits = valid;
        m->unstuff = d == 0xFF;
    }
    m->nbits--;
    return (int)((m->bits >> m->nbits) & 1);
}

static void mel_decode_more(mel_t* m) { /* :45-72 */
    while (m->num_runs < 8) {
        int eval = MEL_E[m->k], run = 0;
        int lead = mel_read_bit(m);
        if (lead == 1) {
            run = (1 << eval) - 1;
            if (m->k < 12) m->k++;
            run <<= 1;
        } else {
            for (int i = 0; i < eval; i++) run = (run << 1) | mel_read_bit(m);
            if (m->k > 0) m->k--;
            run = (run << 1) + 1;
        }
        unsigned shift = (unsigned)(m->num_runs * 7);
        m->runs &= ~((uint64_t)0x3F << shift);
        m->runs |= (uint64_t)run << shift;
        m->num_runs++;
    }
}

static int mel_get_run(mel_t* m) { /* :25-39 */
    if (m->num_runs == 0) mel_decode_more(m);
    if (m->num_runs == 0) return 1 << 30;
    int run = (int)(m->runs & 0x7F);
    m->runs >>= 7;
    m->num_runs--;
    return run;
}

/* vlc_reverse_decoder.go:9-100 */
typedef struct { const uint8_t* data; int len; int pos; uint64_t tmp; int num; int unstuff; int init_done; } rev_t;

static void rev_read_chunk(rev_t* r) { /* :42-87 */
    if (r->num > 32) return;
    uint32_t val = 0;
    int shift = 24;
    for (int i = 0; i < 4 && r->pos >= 0; i++) {
        val |= (uint32_t)r->data[r->pos] << shift;
        r->pos--;
        shift -= 8;
    }
    uint32_t tmp = val >> 24;
    int bits = 8;
    if (r->unstuff && ((val >> 24) & 0x7F) == 0x7F) bits = 7;
    int unstuff = (val >> 24) > 0x8F;
    tmp |= ((val >> 16) & 0xFF) << bits;
    bits += (unstuff && ((val >> 16) & 0x7F) == 0x7F) ? 7 : 8;
    unstuff = ((val >> 16) & 0xFF) > 0x8F;
    tmp |= ((val >> 8) & 0xFF) << bits;
    bits += (unstuff && ((val >> 8) & 0x7F) == 0x7F) ? 7 : 8;
    unstuff = ((val >> 8) & 0xFF) > 0x8F;
    tmp |= (val & 0xFF) << bits;
    bits += (unstuff && (val & 0x7F) == 0x7F) ? 7 : 8;
    r->unstuff = (val & 0xFF) > 0x8F;
    r->tmp |= (uint64_t)tmp << r->num;
    r->num += bits;
}

static int rev_init(rev_t* r) { /* :18-40 */
    if (r->init_done) return 1;
    r->init_done = 1;
    if (r->len < 2) return 0;
    r->pos = r->len - 2;
    uint8_t d = r->data[r->pos];
    r->pos--;
    r->tmp = (uint64_t)(d >> 4);
    r->num = 4;
    if ((r->tmp & 0x7) == 0x7) r->num--;
    r->unstuff = (d | 0x0F) > 0x8F;
    rev_read_chunk(r);
    return 1;
}

static int rev_read_more(rev_t* r, int min_bits) { /* :89-100 */
    if (!rev_init(r)) return 0;
    while (r->num < min_bits) {
        if (r->pos < 0) break;
        rev_read_chunk(r);
    }
    return r->num >= min_bits;
}

/* ht_block_decoder.go:243-264 */
static uint32_t vlc_peek(rev_t* r) {
    (void)rev_read_more(r, 32);
    return (uint32_t)r->tmp;
}
static void vlc_advance(rev_t* r, int n) {
    if (n <= 0) return;
    (void)rev_read_more(r, n);
    if (n > r->num) { r->tmp = 0; r->num = 0; return; }
    r->tmp >>= n;
    r->num -= n;
}

/* magsgn.go:113-204 (forward, LSB first, a byte after 0xFF carries 7 bits, 0xFF padding when exhausted) */
typedef struct { const uint8_t* data; int len; int pos; uint64_t buf; int cnt; uint8_t last; } ms_t;

static uint32_t ms_fetch(ms_t* m, int n) {
    if (n == 0) return 0;
    while (m->cnt < n && m->pos < m->len) {
        uint8_t b = m->data[m->pos++];
        if (m->last == 0xFF) { m->buf |= (uint64_t)(b & 0x7F) << m->cnt; m->cnt += 7; }
        else { m->buf |= (uint64_t)b << m->cnt; m->cnt += 8; }
        m->last = b;
    }
    while (m->cnt < n) {
        uint8_t b = 0xFF;
        if (m->last == 0xFF) { m->buf |= (uint64_t)(b & 0x7F) << m->cnt; m->cnt += 7; }
        else { m->buf |= (uint64_t)b << m->cnt; m->cnt += 8; }
        m->last = b;
    }
    uint64_t v = m->buf & (((uint64_t)1 << n) - 1);
    m->buf >>= n;
    m->cnt -= n;
    return (uint32_t)v;
}

typedef struct { mel_t* mel; rev_t* vlc; int run; } cstate_t;

/* openjph_cleanup_decoder.go:169-178 */
static uint16_t apply_zero_run(cstate_t* s, uint16_t entry) {
    s->run -= 2;
    if (s->run != -1) entry = 0;
    if (s->run < 0) s->run = mel_get_run(s->mel);
    return entry;
}

/* :258-276 */
static void decode_uvlc(int initial, int mode, rev_t* vlc, int* u0, int* u1) {
    uint32_t v = vlc_peek(vlc);
    int idx = mode + (int)(v & 0x3F);
    uint16_t e = initial ? UVLC0[idx] : UVLC1[idx];
    vlc_advance(vlc, e & 7);
    v = vlc_peek(vlc);
    int total_suffix = (e >> 3) & 0xF;
    int tmp = (int)(v & ((1u << total_suffix) - 1));
    vlc_advance(vlc, total_suffix);
    int u0suf = (e >> 7) & 7;
    *u0 = ((e >> 10) & 7) + (tmp & ((1 << u0suf) - 1));
    *u1 = ((e >> 13) & 7) + (tmp >> u0suf);
}

/* :180-219 */
static void initial_row(uint16_t* scratch, int width, cstate_t* st) {
    int cq = 0;
    for (int x = 0, sp = 0; x < width; sp += 4) {
        uint16_t t0 = VLC0[cq + (int)(vlc_peek(st->vlc) & 0x7F)];
        if (cq == 0) t0 = apply_zero_run(st, t0);
        scratch[sp] = t0;
        x += 2;
        cq = ((t0 & 0x10) << 3) | ((t0 & 0xE0) << 2);
        vlc_advance(st->vlc, t0 & 7);

        uint16_t t1 = VLC0[cq + (int)(vlc_peek(st->vlc) & 0x7F)];
        if (cq == 0 && x < width) t1 = apply_zero_run(st, t1);
        if (x >= width) t1 = 0;
        scratch[sp + 2] = t1;
        x += 2;
        cq = ((t1 & 0x10) << 3) | ((t1 & 0xE0) << 2);
        vlc_advance(st->vlc, t1 & 7);

        int mode = ((t0 & 0x8) << 3) | ((t1 & 0x8) << 4);
        if (mode == 0xC0) {
            st->run -= 2;
            if (st->run == -1) mode += 0x40;
            if (st->run < 0) st->run = mel_get_run(st->mel);
        }
        int u0, u1;
        decode_uvlc(1, mode, st->vlc, &u0, &u1);
        scratch[sp + 1] = (uint16_t)(1 + u0);
        scratch[sp + 3] = (uint16_t)(1 + u1);
    }
}

/* :221-256 */
static void remaining_rows(uint16_t* scratch, int width, int height, int sstr, cstate_t* st) {
    for (int y = 2; y < height; y += 2) {
        int cq = 0;
        int sp = (y >> 1) * sstr;
        for (int x = 0; x < width; sp += 4) {
            cq |= ((scratch[sp - sstr] & 0xA0) << 2) | ((scratch[sp - sstr + 2] & 0x20) << 4);
            uint16_t t0 = VLC1[cq + (int)(vlc_peek(st->vlc) & 0x7F)];
            if (cq == 0) t0 = apply_zero_run(st, t0);
            scratch[sp] = t0;
            x += 2;
            cq = ((t0 & 0x40) << 2) | ((t0 & 0x80) << 1);
            cq |= scratch[sp - sstr] & 0x80;
            cq |= ((scratch[sp - sstr + 2] & 0xA0) << 2) | ((scratch[sp - sstr + 4] & 0x20) << 4);
            vlc_advance(st->vlc, t0 & 7);

            uint16_t t1 = VLC1[cq + (int)(vlc_peek(st->vlc) & 0x7F)];
            if (cq == 0 && x < width) t1 = apply_zero_run(st, t1);
            if (x >= width) t1 = 0;
            scratch[sp + 2] = t1;
            x += 2;
            cq = ((t1 & 0x40) << 2) | ((t1 & 0x80) << 1);
            cq |= scratch[sp - sstr + 2] & 0x80;
            vlc_advance(st->vlc, t1 & 7);

            int u0, u1;
            decode_uvlc(0, ((t0 & 0x8) << 3) | ((t1 & 0x8) << 4), st->vlc, &u0, &u1);
            scratch[sp + 1] = (uint16_t)u0;
            scratch[sp + 3] = (uint16_t)u1;
        }
        scratch[sp] = 0;
        scratch[sp + 1] = 0;
    }
}

/* :432-447 */
static void sample_ms(ms_t* ms, uint32_t inf, int uq, int bit, unsigned p, uint32_t* val, uint32_t* vn) {
    *val = 0; *vn = 0;
    if ((inf & (1u << (4 + bit))) == 0) return;
    int mn = uq - (int)((inf >> (12 + bit)) & 1);
    uint32_t msv = ms_fetch(ms, mn);
    uint32_t v = msv << 31;
    uint32_t n = msv & (uint32_t)(((uint64_t)1 << mn) - 1);
    n |= ((inf >> (8 + bit)) & 1) << mn;
    n |= 1;
    v |= (n + 2) << (p - 1);
    *val = v; *vn = n;
}

static int bitlen32(uint32_t v) { return v ? 32 - __builtin_clz(v) : 0; }

/* :278-372; returns 0, or -3 when a U_q exceeds missing_msbs + 2 */
static int scratch_magsgn(const uint8_t* ms_data, int ms_len, const uint16_t* scratch, int width, int height, int sstr, unsigned p,
                          int missing_msbs, uint32_t* out) {
    int mmsbp2 = missing_msbs + 2;
    ms_t ms = {ms_data, ms_len, 0, 0, 0, 0};
    uint32_t* vns = (uint32_t*)calloc((size_t)width + 4, 4);
    uint32_t prev_vn = 0, v, n;
    int sp = 0, vp = 0, rc = 0;
    for (int x = 0; x < width; sp += 2) {
        uint32_t inf = scratch[sp];
        int uq = scratch[sp + 1];
        if (uq > mmsbp2) { rc = -3; goto done; }
        sample_ms(&ms, inf, uq, 0, p, &v, &n);
        out[x] = v;
        sample_ms(&ms, inf, uq, 1, p, &v, &n);
        if (height > 1) out[width + x] = v;
        vns[vp] = prev_vn | n;
        prev_vn = 0;
        x++; vp++;
        if (x >= width) { vp++; break; }
        sample_ms(&ms, inf, uq, 2, p, &v, &n);
        out[x] = v;
        sample_ms(&ms, inf, uq, 3, p, &v, &n);
        if (height > 1) out[width + x] = v;
        prev_vn = n;
        x++;
    }
    vns[vp] = prev_vn;
    for (int y = 2; y < height; y += 2) {
        sp = (y >> 1) * sstr;
        vp = 0;
        prev_vn = 0;
        for (int x = 0; x < width; sp += 2) {
            uint32_t inf = scratch[sp];
            uint32_t uq = scratch[sp + 1];
            uint32_t gamma = inf & 0xF0;
            gamma &= gamma - 0x10;
            uint32_t emax = (uint32_t)(bitlen32((vns[vp] | vns[vp + 1]) | 2) - 1);
            uint32_t kappa = gamma ? emax : 1;
            int Uq = (int)(uq + kappa);
            if (Uq > mmsbp2) { rc = -3; goto done; }
            sample_ms(&ms, inf, Uq, 0, p, &v, &n);
            out[y * width + x] = v;
            sample_ms(&ms, inf, Uq, 1, p, &v, &n);
            if (y + 1 < height) out[(y + 1) * width + x] = v;
            vns[vp] = prev_vn | n;
            prev_vn = 0;
            x++; vp++;
            if (x >= width) { vp++; break; }
            sample_ms(&ms, inf, Uq, 2, p, &v, &n);
            out[y * width + x] = v;
            sample_ms(&ms, inf, Uq, 3, p, &v, &n);
            if (y + 1 < height) out[(y + 1) * width + x] = v;
            prev_vn = n;
            x++;
        }
        vns[vp] = prev_vn;
    }
done:
    free(vns);
    return rc;
}

/* decodeOpenJPHCleanup, openjph_cleanup_decoder.go:115-161 behind HTDecoder.Decode (decoder.go:43-58).
 * out: width*height int32, row-major.  Returns 0, or a negative code where the Go function returns an error
 * (-1 Kmax <= 0, -2 missing MSBs out of range / invalid Scup, -3 U_q out of range); on error `out` is all zeros, which is
 * what TileDecoder.decodeCodeBlock substitutes (t2/tile_decoder.go:718-721). */
EXPORT int orc_ht_decode_block(const uint8_t* cb, int lcup, int width, int height, int kmax, int missing_msbs, int32_t* out) {
    init_tables();
    memset(out, 0, (size_t)width * height * 4);
    if (lcup == 0) return 0;
    if (kmax <= 0) return -1;
    if (missing_msbs < 0 || missing_msbs >= 30) return -2;
    /* parseStandardSegments, decoder.go:60-70 (Go panics below two bytes; every caller holds at least the Scup locator) */
    if (lcup < 2) return -2;
    int scup = ((int)cb[lcup - 1] << 4) | (cb[lcup - 2] & 0x0F);
    if (scup < 2 || scup > lcup || scup > 4079) return -2;
    int ms_len = lcup - scup;
    const uint8_t* cleanup = cb + ms_len;

    unsigned p = (unsigned)(30 - missing_msbs);
    int sstr = ((width + 2) + 7) & ~7;
    uint16_t* scratch = (uint16_t*)calloc((size_t)sstr * ((height + 1) / 2 + 1) + 8, 2);
    mel_t mel = {cleanup, scup, 0, scup - 1, 0, 0, 0, 0, 0, 0};
    rev_t rev = {cleanup, scup, 0, 0, 0, 0, 0};
    cstate_t st = {&mel, &rev, 0};
    st.run = mel_get_run(&mel);
    initial_row(scratch, width, &st);
    int sentinel = ((width + 3) / 4) * 4;
    scratch[sentinel] = 0;
    scratch[sentinel + 1] = 0;
    remaining_rows(scratch, width, height, sstr, &st);

    uint32_t* cbv = (uint32_t*)calloc((size_t)width * height, 4);
    int rc = scratch_magsgn(cb, ms_len, scratch, width, height, sstr, p, missing_msbs, cbv);
    if (rc == 0) {
        unsigned shift = (unsigned)(31 - kmax); /* uint(31 - kmax): a huge count when kmax > 31, and Go shifts >= 32 give 0 */
        for (int i = 0; i < width * height; i++) {
            uint32_t v = cbv[i];
            int32_t mag = shift >= 32 ? 0 : (int32_t)((v & 0x7FFFFFFF) >> shift);
            out[i] = (v & 0x80000000u) ? -mag : mag;
        }
    }
    free(cbv);
    free(scratch);
    return rc;
}

/* Many blocks at once (tests and the bench's CPU leg): descriptors as in include/j2k_b200.h j2k_ht_cblk, block sizes and
 * output offsets (in samples) from the caller's layout.  status[i] receives each block's return code. */
EXPORT void orc_ht_decode_blocks(const uint8_t* bytes, const uint64_t* offsets, const uint32_t* lengths, const uint8_t* kmax,
                                 const uint8_t* mmsb, const int32_t* widths, const int32_t* heights, const int64_t* out_offsets,
                                 long nblocks, int32_t* out, int32_t* status) {
    for (long i = 0; i < nblocks; i++) {
        int rc = orc_ht_decode_block(bytes + offsets[i], (int)lengths[i], widths[i], heights[i], kmax[i], mmsb[i], out + out_offsets[i]);
        if (status) status[i] = rc;
    }
}

/* ------------------------------------------------------------------------------------------------------------------
 * Stream GENERATOR for the tests and the bench: an HT cleanup-pass block encoder written from ISO/IEC 15444-15 clause 7
 * against the decoder above (NOT a restatement of the reference's openjph_cleanup_encoder.go, and not byte-identical to
 * OpenJPH: where several VLC entries fit, it takes the one with the most known bits; MEL and VLC are not fused).  Its only
 * contract is the round trip: orc_ht_decode_block(orc_ht_encode_block(x)) == x, which tests/test_ht_oracle.py checks on
 * every generated block before the stream is used to exercise the CUDA decoder.
 * Input: width*height signed values, |x| <= 2^(missing_msbs + 1); the block decodes to x with kmax = missing_msbs + 1. */

typedef struct { uint8_t* buf; int cap, pos; uint32_t cur; int n, max; } fw_t;   /* MagSgn: forward, LSB first */
static int fw_put(fw_t* w, uint32_t v, int nbits) {
    for (int i = 0; i < nbits; i++) {
        w->cur |= ((v >> i) & 1u) << w->n;
        if (++w->n == w->max) {
            if (w->pos >= w->cap) return -1;
            w->buf[w->pos++] = (uint8_t)w->cur;
            w->max = w->cur == 0xFF ? 7 : 8;
            w->cur = 0; w->n = 0;
        }
    }
    return 0;
}
static int fw_flush(fw_t* w) {
    while (w->n != 0) if (fw_put(w, 1, 1)) return -1;   /* pad with ones (the decoder reads 0xFF beyond the end) */
    return 0;
}

typedef struct { uint8_t* buf; int cap, pos; uint32_t cur; int n, max; int k, run; } melw_t;   /* MEL: forward, MSB first */
static int melw_bit(melw_t* w, int b) {
    w->cur = (w->cur << 1) | (uint32_t)b;
    if (++w->n == w->max) {
        if (w->pos >= w->cap) return -1;
        w->buf[w->pos++] = (uint8_t)w->cur;
        w->max = w->cur == 0xFF ? 7 : 8;
        w->cur = 0; w->n = 0;
    }
    return 0;
}
static int melw_event(melw_t* w, int bit) {
    int e = MEL_E[w->k];
    if (!bit) {
        if (++w->run >= (1 << e)) {
            if (melw_bit(w, 1)) return -1;
            w->run = 0;
            if (w->k < 12) w->k++;
        }
        return 0;
    }
    if (melw_bit(w, 0)) return -1;
    for (int i = e - 1; i >= 0; i--) if (melw_bit(w, (w->run >> i) & 1)) return -1;
    w->run = 0;
    if (w->k > 0) w->k--;
    return 0;
}
static int melw_flush(melw_t* w) {
    if (w->run > 0 && melw_bit(w, 1)) return -1;
    while (w->n != 0) if (melw_bit(w, 0)) return -1;
    return 0;
}

/* VLC: written LSB first into bytes that end up in reverse order; a byte that follows one > 0x8F stops at 7 bits when those
 * are all ones (the reverse reader's un-stuffing rule, vlc_reverse_decoder.go:42-87) */
typedef struct { uint8_t* buf; int cap, pos; uint32_t cur; int n; int prev_big; } vlcw_t;
static int vlcw_emit(vlcw_t* w) {
    if (w->pos >= w->cap) return -1;
    w->buf[w->pos++] = (uint8_t)w->cur;
    w->prev_big = w->cur > 0x8F;
    w->cur = 0; w->n = 0;
    return 0;
}
static int vlcw_put(vlcw_t* w, uint32_t v, int nbits) {
    for (int i = 0; i < nbits; i++) {
        w->cur |= ((v >> i) & 1u) << w->n;
        w->n++;
        if (w->n == 7 && w->prev_big && (w->cur & 0x7F) == 0x7F) { if (vlcw_emit(w)) return -1; }
        else if (w->n == 8) { if (vlcw_emit(w)) return -1; }
    }
    return 0;
}

static int enc_lookup(const unsigned char (*src)[7], int n, int cq, int rho, int uoff, int emb, int* cwd, int* len, int* ek) {
    int best = -1, bestpop = -1;
    for (int j = 0; j < n; j++) {
        const unsigned char* e = src[j];
        if (e[0] != cq || e[1] != rho || e[2] != uoff) continue;
        if (uoff) { if ((emb & e[3]) != e[4]) continue; }
        else if (e[3] || e[4]) continue;
        int pop = __builtin_popcount(e[3]);
        if (pop > bestpop) { bestpop = pop; best = j; }
    }
    if (best < 0) return -1;
    *cwd = src[best][5]; *len = src[best][6]; *ek = src[best][3];
    return 0;
}

/* u >= 1: prefix (value, bits) and suffix (value, bits) of the U-VLC code, bits LSB first (uvlc_tables.go:41-51) */
static void uvlc_code(int u, int* pv, int* pl, int* sv, int* sl) {
    if (u == 1) { *pv = 1; *pl = 1; *sv = 0; *sl = 0; }
    else if (u == 2) { *pv = 2; *pl = 2; *sv = 0; *sl = 0; }
    else if (u <= 4) { *pv = 4; *pl = 3; *sv = u - 3; *sl = 1; }
    else { *pv = 0; *pl = 3; *sv = u - 5; *sl = 5; }
}

EXPORT int orc_ht_encode_block(const int32_t* x, int width, int height, int missing_msbs, uint8_t* out, int cap) {
    init_tables();
    const int qw = (width + 1) / 2, qh = (height + 1) / 2;
    int any = 0;
    for (int i = 0; i < width * height; i++) any |= x[i] != 0;
    if (!any) return 0;   /* not included: no bytes */
    uint8_t* msb = (uint8_t*)malloc((size_t)cap);
    uint8_t* melb = (uint8_t*)malloc((size_t)cap);
    uint8_t* vlcb = (uint8_t*)malloc((size_t)cap);
    uint8_t* rho_prev = (uint8_t*)calloc((size_t)qw + 2, 1);
    uint8_t* rho_cur = (uint8_t*)calloc((size_t)qw + 2, 1);
    uint32_t* vn1 = (uint32_t*)calloc((size_t)qw + 2, 4), *vn3 = (uint32_t*)calloc((size_t)qw + 2, 4);   /* row above, index q + 1 */
    uint32_t* cn1 = (uint32_t*)calloc((size_t)qw + 2, 4), *cn3 = (uint32_t*)calloc((size_t)qw + 2, 4);
    fw_t ms = {msb, cap, 0, 0, 0, 8};
    melw_t mel = {melb, cap, 0, 0, 0, 8, 0, 0};
    vlcw_t vlc = {vlcb, cap, 0, 0xF, 4, 1};   /* the first byte's low nibble belongs to Scup and counts as 0xF */
    int rc = 0;
    for (int qy = 0; qy < qh && !rc; qy++) {
        memset(rho_cur, 0, (size_t)qw + 2);
        memset(cn1, 0, ((size_t)qw + 2) * 4); memset(cn3, 0, ((size_t)qw + 2) * 4);
        for (int q0 = 0; q0 < qw && !rc; q0 += 2) {
            int uoff[2] = {0, 0}, u[2] = {0, 0};
            for (int k = 0; k < 2; k++) {
                int q = q0 + k;
                if (q >= qw) break;
                uint32_t vn[4] = {0, 0, 0, 0}; int sgn[4] = {0, 0, 0, 0}, rho = 0, maxe = 0;
                for (int i = 0; i < 4; i++) {
                    int xx = 2 * q + (i >> 1), yy = 2 * qy + (i & 1);
                    if (xx >= width || yy >= height) continue;
                    int32_t v = x[yy * width + xx];
                    if (!v) continue;
                    uint32_t m = v < 0 ? (uint32_t)(-(int64_t)v) : (uint32_t)v;
                    rho |= 1 << i; sgn[i] = v < 0; vn[i] = 2 * m - 1;
                    int e = bitlen32(vn[i]);
                    if (e > maxe) maxe = e;
                }
                rho_cur[q + 1] = (uint8_t)rho;
                cn1[q + 1] = vn[1]; cn3[q + 1] = vn[3];
                int cq, kappa = 1;
                uint8_t L = rho_cur[q];   /* left quad (index q - 1 + 1) */
                if (qy == 0) cq = (((L & 1) | ((L >> 1) & 1))) | (((L >> 2) & 1) << 1) | (((L >> 3) & 1) << 2);
                else {
                    uint8_t Aw = rho_prev[q], A = rho_prev[q + 1], Ae = rho_prev[q + 2];
                    cq = (((Aw >> 3) & 1) | ((A >> 1) & 1)) | ((((L >> 2) & 1) | ((L >> 3) & 1)) << 1) | ((((A >> 3) & 1) | ((Ae >> 1) & 1)) << 2);
                    if (__builtin_popcount(rho) > 1) {
                        uint32_t o = vn3[q] | vn1[q + 1] | vn3[q + 1] | vn1[q + 2];
                        kappa = bitlen32(o | 2) - 1;
                    }
                }
                if (cq == 0 && melw_event(&mel, rho != 0)) { rc = -1; break; }
                if (rho == 0 && cq == 0) continue;
                int Uq = maxe > kappa ? maxe : kappa;
                if (rho == 0) Uq = kappa;
                if (Uq > missing_msbs + 2) { rc = -2; break; }
                u[k] = Uq - kappa; uoff[k] = u[k] > 0;
                int emb = 0;
                if (uoff[k]) for (int i = 0; i < 4; i++) if ((rho >> i & 1) && bitlen32(vn[i]) == Uq) emb |= 1 << i;
                int cwd, len, ek;
                if (qy == 0 ? enc_lookup(HT_VLC_SRC0, (int)(sizeof(HT_VLC_SRC0) / 7), cq, rho, uoff[k], emb, &cwd, &len, &ek)
                            : enc_lookup(HT_VLC_SRC1, (int)(sizeof(HT_VLC_SRC1) / 7), cq, rho, uoff[k], emb, &cwd, &len, &ek)) { rc = -3; break; }
                if (vlcw_put(&vlc, (uint32_t)cwd, len)) { rc = -1; break; }
                for (int i = 0; i < 4; i++) {
                    if (!(rho >> i & 1)) continue;
                    int mn = Uq - ((ek >> i) & 1);
                    uint32_t val = (vn[i] & ~1u) | (uint32_t)sgn[i];
                    if (fw_put(&ms, mn >= 32 ? val : (val & ((1u << mn) - 1)), mn)) { rc = -1; break; }
                }
            }
            if (rc) break;
            int m0 = uoff[0], m1 = uoff[1], a = u[0], b = u[1];
            int pv, pl, sv, sl, pv2, pl2, sv2, sl2;
            if (m0 && m1) {
                if (qy == 0) {
                    int big = (a < b ? a : b) > 2;
                    if (melw_event(&mel, big)) { rc = -1; break; }
                    if (big) { a -= 2; b -= 2; }
                    else if (a > 2) {   /* u1 is 1 or 2: one bit behind the first prefix */
                        uvlc_code(a, &pv, &pl, &sv, &sl);
                        if (vlcw_put(&vlc, (uint32_t)pv, pl) || vlcw_put(&vlc, (uint32_t)(b - 1), 1) || vlcw_put(&vlc, (uint32_t)sv, sl)) rc = -1;
                        continue;
                    }
                }
                uvlc_code(a, &pv, &pl, &sv, &sl);
                uvlc_code(b, &pv2, &pl2, &sv2, &sl2);
                if (vlcw_put(&vlc, (uint32_t)pv, pl) || vlcw_put(&vlc, (uint32_t)pv2, pl2) || vlcw_put(&vlc, (uint32_t)sv, sl) || vlcw_put(&vlc, (uint32_t)sv2, sl2)) rc = -1;
            } else if (m0 || m1) {
                uvlc_code(m0 ? a : b, &pv, &pl, &sv, &sl);
                if (vlcw_put(&vlc, (uint32_t)pv, pl) || vlcw_put(&vlc, (uint32_t)sv, sl)) rc = -1;
            }
        }
        uint8_t* t = rho_prev; rho_prev = rho_cur; rho_cur = t;
        uint32_t* tn = vn1; vn1 = cn1; cn1 = tn;
        tn = vn3; vn3 = cn3; cn3 = tn;
    }
    int lcup = rc;
    if (!rc) {
        if (fw_flush(&ms) || melw_flush(&mel)) rc = -1;
        if (!rc && vlc.n > 0 && vlcw_emit(&vlc)) rc = -1;
        if (!rc && vlc.pos == 0 && vlcw_emit(&vlc)) rc = -1;   /* the nibble byte always exists */
        int scup = mel.pos + vlc.pos + 1;
        lcup = ms.pos + scup;
        if (!rc && (scup > 4079 || lcup > cap)) rc = -4;
        if (!rc) {
            memcpy(out, msb, (size_t)ms.pos);
            memcpy(out + ms.pos, melb, (size_t)mel.pos);
            for (int i = 0; i < vlc.pos; i++) out[ms.pos + mel.pos + i] = vlcb[vlc.pos - 1 - i];
            out[lcup - 1] = (uint8_t)(scup >> 4);
            out[lcup - 2] = (uint8_t)((out[lcup - 2] & 0xF0) | (scup & 0xF));
        } else lcup = rc;
    }
    free(msb); free(melb); free(vlcb); free(rho_prev); free(rho_cur); free(vn1); free(vn3); free(cn1); free(cn3);
    return lcup;
}
