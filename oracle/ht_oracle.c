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

/* ------------------------------------------------------------------------------------------------------------------
 * ORACLE of the HTJ2K cleanup-pass block ENCODER: a loop-for-loop restatement of HTEncoder.Encode ->
 * encodeOpenJPHCleanup, /root/reference/jpeg2000/htj2k/encoder.go:54-68 and openjph_cleanup_encoder.go (line numbers below).
 * PINNED: tests/test_ht_oracle.py re-encodes every code-block of the 14 OpenJPH interop codestreams from its decoded
 * coefficients and compares with the bytes in the codestream -- the block-level form of the reference's own byte-parity
 * test (htj2k/go_byte_parity_test.go:11-44). */

typedef struct { uint8_t* buf; int n; int tmp, remaining, run, k, threshold; } omel_t;       /* :8-62 */
static void omel_emit(omel_t* m, int v) {
    m->tmp = (m->tmp << 1) | (v & 1);
    if (--m->remaining == 0) {
        m->buf[m->n++] = (uint8_t)m->tmp;
        m->remaining = m->tmp == 0xFF ? 7 : 8;
        m->tmp = 0;
    }
}
static void omel_encode(omel_t* m, int bit) {
    if (!bit) {
        if (++m->run >= m->threshold) {
            omel_emit(m, 1);
            m->run = 0;
            if (m->k < 12) m->k++;
            m->threshold = 1 << MEL_E[m->k];
        }
        return;
    }
    omel_emit(m, 0);
    for (int t = MEL_E[m->k]; t > 0;) { t--; omel_emit(m, (m->run >> t) & 1); }
    m->run = 0;
    if (m->k > 0) m->k--;
    m->threshold = 1 << MEL_E[m->k];
}

typedef struct { uint8_t* buf; int n; int used, tmp, last_gt_8f; } ovlc_t;                   /* :64-112 */
static void ovlc_encode(ovlc_t* v, int cwd, int len) {
    while (len > 0) {
        int avail = 8;
        if (v->last_gt_8f) avail--;
        avail -= v->used;
        int t = avail < len ? avail : len;
        v->tmp |= (cwd & ((1 << t) - 1)) << v->used;
        v->used += t;
        avail -= t;
        len -= t;
        cwd >>= t;
        if (avail == 0) {
            if (v->last_gt_8f && v->tmp != 0x7F) { v->last_gt_8f = 0; continue; }
            v->buf[v->n++] = (uint8_t)v->tmp;
            v->last_gt_8f = v->tmp > 0x8F;
            v->tmp = 0;
            v->used = 0;
        }
    }
}

typedef struct { uint8_t* buf; int n; int max_bits, used; uint32_t tmp; } oms_t;            /* :114-166 */
static void oms_encode(oms_t* m, uint32_t cwd, int len) {
    while (len > 0) {
        int t = m->max_bits - m->used < len ? m->max_bits - m->used : len;
        m->tmp |= (cwd & (uint32_t)(((uint64_t)1 << t) - 1)) << m->used;
        m->used += t;
        cwd = t >= 32 ? 0 : cwd >> t;
        len -= t;
        if (m->used >= m->max_bits) {
            uint8_t b = (uint8_t)m->tmp;
            m->buf[m->n++] = b;
            m->max_bits = b == 0xFF ? 7 : 8;
            m->tmp = 0;
            m->used = 0;
        }
    }
}
static void oms_terminate(oms_t* m) {
    if (m->used != 0) {
        int t = m->max_bits - m->used;
        m->tmp |= (uint32_t)((0xFF & ((1 << t) - 1)) << m->used);
        m->used += t;
        if ((uint8_t)m->tmp != 0xFF) m->buf[m->n++] = (uint8_t)m->tmp;
    } else if (m->max_bits == 7 && m->n > 0) m->n--;
}

typedef struct { int pre, pre_len, suf, suf_len, ext, ext_len; } ouvlc_t;                    /* :168-198 */
static ouvlc_t ouvlc(int code) {
    ouvlc_t e = {0, 0, 0, 0, 0, 0};
    if (code <= 0) return e;
    if (code == 1) { e.pre = 1; e.pre_len = 1; return e; }
    if (code == 2) { e.pre = 2; e.pre_len = 2; return e; }
    if (code <= 4) { e.pre = 4; e.pre_len = 3; e.suf = code - 3; e.suf_len = 1; return e; }
    if (code <= 32) { e.pre = 0; e.pre_len = 3; e.suf = code - 5; e.suf_len = 5; return e; }
    e.pre = 0; e.pre_len = 3; e.suf = 28 + ((code - 33) % 4); e.suf_len = 5; e.ext = (code - 33) / 4; e.ext_len = 4;
    return e;
}

static uint16_t ENC0[2048], ENC1[2048];
static int enc_ready;
static void init_enc_table(const unsigned char (*src)[7], int n, uint16_t* dst) {            /* :432-470 */
    for (int i = 0; i < 2048; i++) {
        int cq = i >> 8, rho = (i >> 4) & 0xF, eps = i & 0xF;
        dst[i] = 0;
        if ((eps & rho) != eps || (rho == 0 && cq == 0)) continue;
        const unsigned char* best = NULL;
        if (eps != 0) {
            int best_ek = -1;
            for (int j = 0; j < n; j++) {
                const unsigned char* e = src[j];
                if (e[0] == cq && e[1] == rho && e[2] == 1 && (eps & e[3]) == e[4]) {
                    int ones = __builtin_popcount(e[3]);
                    if (ones >= best_ek) { best = e; best_ek = ones; }
                }
            }
        } else {
            for (int j = 0; j < n; j++) {
                const unsigned char* e = src[j];
                if (e[0] == cq && e[1] == rho && e[2] == 0) { best = e; break; }
            }
        }
        if (best) dst[i] = (uint16_t)((best[5] << 8) | (best[6] << 4) | best[3]);
    }
}
static void init_enc(void) {
    if (enc_ready) return;
    init_enc_table(HT_VLC_SRC0, (int)(sizeof(HT_VLC_SRC0) / 7), ENC0);
    init_enc_table(HT_VLC_SRC1, (int)(sizeof(HT_VLC_SRC1) / 7), ENC1);
    enc_ready = 1;
}
/* which: 0 ojphEncoderVLCTable0, 1 ojphEncoderVLCTable1 (2048 entries each) */
EXPORT int orc_ht_enc_table(int which, uint16_t* out) {
    init_enc();
    memcpy(out, which ? ENC1 : ENC0, 2048 * 2);
    return 2048;
}

typedef struct { const uint32_t* cb; int width, height; unsigned p; } oenc_t;

static void prep_sample(const oenc_t* h, int x, int y, int idx, int* rho, int* eqmax, int* eq, uint32_t* s) {   /* :396-413 */
    if (x >= h->width || y >= h->height) return;
    uint32_t t = h->cb[y * h->width + x];
    uint32_t val = (t + t) >> h->p;
    val &= ~(uint32_t)1;
    if (val == 0) return;
    *rho += 1 << (idx % 4);
    val--;
    eq[idx] = bitlen32(val);
    if (eq[idx] > *eqmax) *eqmax = eq[idx];
    val--;
    s[idx] = val + (t >> 31);
}
static void prep_quad(const oenc_t* h, int x, int y, int off, int* rho, int* eqmax, int* eq, uint32_t* s) {      /* :382-394 */
    prep_sample(h, x, y, off, rho, eqmax, eq, s);
    prep_sample(h, x, y + 1, off + 1, rho, eqmax, eq, s);
    prep_sample(h, x + 1, y, off + 2, rho, eqmax, eq, s);
    prep_sample(h, x + 1, y + 1, off + 3, rho, eqmax, eq, s);
}
static int oeps(const int* eq, int eqmax, int u) {                                          /* :415-426 */
    if (u <= 0) return 0;
    int eps = 0;
    for (int i = 0; i < 4; i++) if (eq[i] == eqmax) eps |= 1 << i;
    return eps;
}
static int otuple(int initial, int cq, int rho, int eps) {                                  /* :428-438 */
    if (rho == 0 && cq == 0) return 0;
    return initial ? ENC0[(cq << 8) | (rho << 4) | eps] : ENC1[(cq << 8) | (rho << 4) | eps];
}
static void oms_quad(oms_t* ms, int rho, int uq, int tuple, const uint32_t* s) {            /* :472-483 */
    for (int i = 0; i < 4; i++) {
        if (!(rho & (1 << i))) continue;
        int m = uq - ((tuple >> i) & 1);
        if (m < 0) m = 0;
        oms_encode(ms, s[i] & (uint32_t)(((uint64_t)1 << m) - 1), m);
    }
}
static void ouvlc_initial(ovlc_t* v, int u0, int u1) {                                      /* :485-511 */
    ouvlc_t c0, c1;
    if (u0 > 2 && u1 > 2) {
        c0 = ouvlc(u0 - 2); c1 = ouvlc(u1 - 2);
        ovlc_encode(v, c0.pre, c0.pre_len); ovlc_encode(v, c1.pre, c1.pre_len);
        ovlc_encode(v, c0.suf, c0.suf_len); ovlc_encode(v, c1.suf, c1.suf_len);
        return;
    }
    if (u0 > 2 && u1 > 0) {
        c0 = ouvlc(u0);
        ovlc_encode(v, c0.pre, c0.pre_len); ovlc_encode(v, u1 - 1, 1); ovlc_encode(v, c0.suf, c0.suf_len);
        return;
    }
    c0 = ouvlc(u0); c1 = ouvlc(u1);
    ovlc_encode(v, c0.pre, c0.pre_len); ovlc_encode(v, c1.pre, c1.pre_len);
    ovlc_encode(v, c0.suf, c0.suf_len); ovlc_encode(v, c1.suf, c1.suf_len);
}
static void ouvlc_noninitial(ovlc_t* v, int u0, int u1) {                                   /* :513-520 */
    ouvlc_t c0 = ouvlc(u0), c1 = ouvlc(u1);
    ovlc_encode(v, c0.pre, c0.pre_len); ovlc_encode(v, c1.pre, c1.pre_len);
    ovlc_encode(v, c0.suf, c0.suf_len); ovlc_encode(v, c1.suf, c1.suf_len);
}
static int imax(int a, int b) { return a > b ? a : b; }
static int imin(int a, int b) { return a < b ? a : b; }

/* HTEncoder.Encode (encoder.go:54-68) -> encodeOpenJPHCleanup (openjph_cleanup_encoder.go:200-252).
 * Returns the number of bytes written to out (0: the block is empty, "return nil, nil"), or
 * -1 invalid Kmax (:201-203 / encoder.go:64-66), -5 "HTJ2K cleanup suffix is empty" (:246-248), -6 out too small. */
EXPORT int orc_ht_encode_ref(const int32_t* data, int width, int height, int kmax, uint8_t* out, int cap) {
    init_tables();
    init_enc();
    if (kmax <= 0 || kmax >= 31) return -1;
    const int n = width * height;
    uint32_t* cb = (uint32_t*)malloc((size_t)n * 4);
    unsigned shift = (unsigned)(31 - kmax);
    uint32_t max_val = 0;
    for (int i = 0; i < n; i++) {
        uint32_t sign = 0;
        int32_t mag = data[i];
        if (mag < 0) { sign = 0x80000000u; mag = (int32_t)(0u - (uint32_t)mag); }
        uint32_t val = (uint32_t)mag << shift;
        cb[i] = sign | val;
        max_val |= val;
    }
    if (max_val < ((uint32_t)1 << shift)) { free(cb); return 0; }
    int missing_msbs = kmax - 1;
    oenc_t h = {cb, width, height, (unsigned)(30 - missing_msbs)};
    size_t room = (size_t)n * 5 + 256;
    omel_t mel = {(uint8_t*)malloc(room), 0, 0, 8, 0, 0, 1};
    ovlc_t vlc = {(uint8_t*)malloc(room), 0, 4, 0xF, 1};
    vlc.buf[vlc.n++] = 0xFF;
    oms_t ms = {(uint8_t*)malloc(room), 0, 8, 0, 0};
    int qn = (width + 1) / 2 + 2;
    uint8_t* e_val = (uint8_t*)calloc((size_t)qn + 2, 1);
    uint8_t* cx_val = (uint8_t*)calloc((size_t)qn + 2, 1);
    {   /* encodeOJPHInitialRows, :254-313 */
        int lep = 0, lcxp = 0, cq0 = 0;
        e_val[lep] = 0; cx_val[lcxp] = 0;
        for (int x = 0; x < width; x += 4) {
            int eqmax[2] = {0, 0}, eq[8] = {0}, rho[2] = {0, 0};
            uint32_t s[8] = {0};
            prep_quad(&h, x, 0, 0, &rho[0], &eqmax[0], eq, s);
            int uq0 = imax(eqmax[0], 1), u0 = uq0 - 1;
            int eps0 = oeps(eq, eqmax[0], u0);
            e_val[lep] = (uint8_t)imax(e_val[lep], eq[1]);
            lep++;
            e_val[lep] = (uint8_t)eq[3];
            cx_val[lcxp] |= (uint8_t)((rho[0] & 2) >> 1);
            lcxp++;
            cx_val[lcxp] = (uint8_t)((rho[0] & 8) >> 3);
            int t0 = otuple(1, cq0, rho[0], eps0);
            ovlc_encode(&vlc, t0 >> 8, (t0 >> 4) & 7);
            if (cq0 == 0) omel_encode(&mel, rho[0] != 0);
            oms_quad(&ms, rho[0], uq0, t0, s);
            int u1 = 0;
            if (x + 2 < width) {
                prep_quad(&h, x + 2, 0, 4, &rho[1], &eqmax[1], eq, s);
                int cq1 = (rho[0] >> 1) | (rho[0] & 1);
                int uq1 = imax(eqmax[1], 1);
                u1 = uq1 - 1;
                int eps1 = oeps(eq + 4, eqmax[1], u1);
                e_val[lep] = (uint8_t)imax(e_val[lep], eq[5]);
                lep++;
                e_val[lep] = (uint8_t)eq[7];
                cx_val[lcxp] |= (uint8_t)((rho[1] & 2) >> 1);
                lcxp++;
                cx_val[lcxp] = (uint8_t)((rho[1] & 8) >> 3);
                int t1 = otuple(1, cq1, rho[1], eps1);
                ovlc_encode(&vlc, t1 >> 8, (t1 >> 4) & 7);
                if (cq1 == 0) omel_encode(&mel, rho[1] != 0);
                oms_quad(&ms, rho[1], uq1, t1, s + 4);
            }
            if (u0 > 0 && u1 > 0) omel_encode(&mel, imin(u0, u1) > 2);
            ouvlc_initial(&vlc, u0, u1);
            cq0 = (rho[1] >> 1) | (rho[1] & 1);
        }
        e_val[lep + 1] = 0;
    }
    for (int y = 2; y < height; y += 2) {   /* encodeOJPHSubsequentRows, :315-380 */
        int lep = 0;
        int max_e = imax(e_val[lep], e_val[lep + 1]) - 1;
        e_val[lep] = 0;
        int lcxp = 0;
        int cq0 = cx_val[lcxp] + (cx_val[lcxp + 1] << 2);
        cx_val[lcxp] = 0;
        for (int x = 0; x < width; x += 4) {
            int eqmax[2] = {0, 0}, eq[8] = {0}, rho[2] = {0, 0};
            uint32_t s[8] = {0};
            prep_quad(&h, x, y, 0, &rho[0], &eqmax[0], eq, s);
            int kappa = 1;
            if (rho[0] & (rho[0] - 1)) kappa = imax(1, max_e);
            int uq0 = imax(eqmax[0], kappa), u0 = uq0 - kappa;
            int eps0 = oeps(eq, eqmax[0], u0);
            e_val[lep] = (uint8_t)imax(e_val[lep], eq[1]);
            lep++;
            max_e = imax(e_val[lep], e_val[lep + 1]) - 1;
            e_val[lep] = (uint8_t)eq[3];
            cx_val[lcxp] |= (uint8_t)((rho[0] & 2) >> 1);
            lcxp++;
            int cq1 = cx_val[lcxp] + (cx_val[lcxp + 1] << 2);
            cx_val[lcxp] = (uint8_t)((rho[0] & 8) >> 3);
            int t0 = otuple(0, cq0, rho[0], eps0);
            ovlc_encode(&vlc, t0 >> 8, (t0 >> 4) & 7);
            if (cq0 == 0) omel_encode(&mel, rho[0] != 0);
            oms_quad(&ms, rho[0], uq0, t0, s);
            int u1 = 0;
            if (x + 2 < width) {
                prep_quad(&h, x + 2, y, 4, &rho[1], &eqmax[1], eq, s);
                kappa = 1;
                if (rho[1] & (rho[1] - 1)) kappa = imax(1, max_e);
                cq1 |= ((rho[0] & 4) >> 1) | ((rho[0] & 8) >> 2);
                int uq1 = imax(eqmax[1], kappa);
                u1 = uq1 - kappa;
                int eps1 = oeps(eq + 4, eqmax[1], u1);
                e_val[lep] = (uint8_t)imax(e_val[lep], eq[5]);
                lep++;
                max_e = imax(e_val[lep], e_val[lep + 1]) - 1;
                e_val[lep] = (uint8_t)eq[7];
                cx_val[lcxp] |= (uint8_t)((rho[1] & 2) >> 1);
                lcxp++;
                cq0 = cx_val[lcxp] + (cx_val[lcxp + 1] << 2);
                cx_val[lcxp] = (uint8_t)((rho[1] & 8) >> 3);
                int t1 = otuple(0, cq1, rho[1], eps1);
                ovlc_encode(&vlc, t1 >> 8, (t1 >> 4) & 7);
                if (cq1 == 0) omel_encode(&mel, rho[1] != 0);
                oms_quad(&ms, rho[1], uq1, t1, s + 4);
            }
            ouvlc_noninitial(&vlc, u0, u1);
            cq0 |= ((rho[1] & 4) >> 1) | ((rho[1] & 8) >> 2);
        }
    }
    /* terminateOJPHMELVLC, :522-545 */
    if (mel.run > 0) omel_emit(&mel, 1);
    mel.tmp <<= mel.remaining;
    int mel_mask = (0xFF << mel.remaining) & 0xFF;
    int vlc_mask = vlc.used > 0 ? 0xFF >> (8 - vlc.used) : 0;
    if ((mel_mask | vlc_mask) != 0) {
        int fuse = mel.tmp | vlc.tmp;
        if ((((fuse ^ mel.tmp) & mel_mask) | ((fuse ^ vlc.tmp) & vlc_mask)) == 0 && fuse != 0xFF && vlc.n > 1) {
            mel.buf[mel.n++] = (uint8_t)fuse;
        } else {
            mel.buf[mel.n++] = (uint8_t)mel.tmp;
            vlc.buf[vlc.n++] = (uint8_t)vlc.tmp;
        }
    }
    oms_terminate(&ms);
    int rc;
    int suffix = mel.n + vlc.n;
    if (suffix == 0) rc = -5;
    else if (ms.n + suffix > cap) rc = -6;
    else {
        memcpy(out, ms.buf, (size_t)ms.n);
        memcpy(out + ms.n, mel.buf, (size_t)mel.n);
        /* ojphVLCWriter.bytes(), :104-112: newest byte first, the Scup placeholder last */
        int o = ms.n + mel.n;
        for (int i = vlc.n - 1; i >= 1; i--) out[o++] = vlc.buf[i];
        out[o++] = vlc.buf[0];
        rc = o;
        if (rc >= 2) {   /* writeScupLocator, encoder.go:84-90 */
            out[rc - 1] = (uint8_t)(suffix >> 4);
            out[rc - 2] = (uint8_t)((out[rc - 2] & 0xF0) | (suffix & 0x0F));
        }
    }
    free(cb); free(mel.buf); free(vlc.buf); free(ms.buf); free(e_val); free(cx_val);
    return rc;
}
