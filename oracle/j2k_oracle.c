/*
 * j2k_oracle.c — CPU restatement of go-dicom-codec's JPEG 2000 sample-domain path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call this file, and only as the checker / the CPU
 * baseline.  The product (go-dicom-codec_b200/csrc) never links it and has no
 * CPU fallback.
 *
 * What it is: a loop-for-loop C restatement of the reference's Go code for the
 * path (same pass order, same column gather through a scratch line, same
 * per-call scratch allocation, same float32/float64 rounding points), written
 * from the Go source text because no Go toolchain exists in this image (the
 * reference cannot be compiled or run here: `go`, `gccgo` absent).  Each function
 * cites the reference file:line it follows (paths relative to the reference root).
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fPIC -shared (oracle/Makefile).
 * x86-64 SSE scalar float arithmetic == Go on amd64: every float32/float64
 * operation is individually rounded, nothing is fused.
 *
 * Parity pinning (tests/test_oracle_*.py, -m "not gpu"):
 *   - the reference's own known-answer tests for this path: QCD bytes
 *     (jpeg2000/quantization_test.go:68-86, openjpeg_lossless_flow_test.go:70-92),
 *     quantizer rounding 0.49/1.0 -> 31 (openjpeg_lossless_flow_test.go:94-105),
 *     float->int KATs (wavelet/dwt97_test.go:405-454), LL sizes
 *     (wavelet/layout_test.go), tile bounds (tile_assembler_test.go), the 5/3 and
 *     9/7 round-trip / identity contracts (wavelet/dwt53_test.go, dwt97_test.go);
 *   - golden vectors produced by OpenJPEG 2.5.4 (the library the Go code clones,
 *     jpeg2000/encoder.go:1790) through Pillow in the build container:
 *     tests/golden/make_golden.py -> tests/golden/ (9/7 + quantization +
 *     dequantization end-to-end pixels; 5/3 LL bands via reduced-resolution decode);
 *   - the 7 raw images of test-data/htj2k/interop (copied as fixtures).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/j2k_b200.h"

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ parity.go */

/* jpeg2000/wavelet/parity.go:3-8 */
static int split_lengths(int n, int even) { return even ? (n + 1) / 2 : n / 2; }
/* jpeg2000/wavelet/parity.go:10-12 */
static int is_even(int v) { return (v & 1) == 0; }
/* jpeg2000/wavelet/parity.go:14-16 */
static int next_coord(int v) { return (v + 1) >> 1; }

/* jpeg2000/wavelet/layout.go:35-44 */
static void next_lowpass_window(int* w, int* h, int* x0, int* y0) {
    int even_row = is_even(*x0), even_col = is_even(*y0);
    *w = split_lengths(*w, even_row);
    *h = split_lengths(*h, even_col);
    *x0 = next_coord(*x0);
    *y0 = next_coord(*y0);
}

/* jpeg2000/wavelet/layout.go:11-33 */
ORC_API void orc_ll_dimensions(int width, int height, int levels, int x0, int y0, int* llw, int* llh) {
    if (width <= 0 || height <= 0) { *llw = 0; *llh = 0; return; }
    if (levels <= 0) { *llw = width; *llh = height; return; }
    int cw = width, ch = height, cx = x0, cy = y0;
    for (int l = 0; l < levels; l++) {
        if (cw <= 1 && ch <= 1) break;
        next_lowpass_window(&cw, &ch, &cx, &cy);
    }
    *llw = cw; *llh = ch;
}

/* ------------------------------------------------------------------- dwt53.go */

/* jpeg2000/wavelet/dwt53.go:27-103 (Forward53_1DWithParity) */
ORC_API void orc_fwd53_1d(int32_t* data, int width, int even) {
    if (even) {
        if (width <= 1) return;
        int32_t sn = (int32_t)((width + 1) >> 1);
        int32_t dn = (int32_t)(width - sn);
        int32_t* tmp = (int32_t*)calloc((size_t)width, sizeof(int32_t));
        int32_t i;
        for (i = 0; i < sn - 1; i++)
            tmp[sn + i] = data[2 * i + 1] - ((data[i * 2] + data[(i + 1) * 2]) >> 1);
        if ((width % 2) == 0) tmp[sn + i] = data[2 * i + 1] - data[i * 2];
        data[0] += (tmp[sn] + tmp[sn] + 2) >> 2;
        for (i = 1; i < dn; i++)
            data[i] = data[2 * i] + ((tmp[sn + (i - 1)] + tmp[sn + i] + 2) >> 2);
        if ((width % 2) == 1)
            data[i] = data[2 * i] + ((tmp[sn + (i - 1)] + tmp[sn + (i - 1)] + 2) >> 2);
        memcpy(data + sn, tmp + sn, (size_t)dn * sizeof(int32_t));
        free(tmp);
    } else {
        if (width == 1) { data[0] *= 2; return; }
        if (width <= 0) return;
        int32_t sn = (int32_t)(width >> 1);
        int32_t dn = (int32_t)(width - sn);
        int32_t* tmp = (int32_t*)calloc((size_t)width, sizeof(int32_t));
        tmp[sn + 0] = data[0] - data[1];
        int32_t i;
        for (i = 1; i < sn; i++)
            tmp[sn + i] = data[2 * i] - ((data[2 * i + 1] + data[2 * (i - 1) + 1]) >> 1);
        if ((width % 2) == 1) tmp[sn + i] = data[2 * i] - data[2 * (i - 1) + 1];
        for (i = 0; i < dn - 1; i++)
            data[i] = data[2 * i + 1] + ((tmp[sn + i] + tmp[sn + i + 1] + 2) >> 2);
        if ((width % 2) == 0) data[i] = data[2 * i + 1] + ((tmp[sn + i] + tmp[sn + i] + 2) >> 2);
        memcpy(data + sn, tmp + sn, (size_t)dn * sizeof(int32_t));
        free(tmp);
    }
}

/* jpeg2000/wavelet/dwt53.go:123-234 (Inverse53_1DWithParity) */
ORC_API void orc_inv53_1d(int32_t* data, int width, int even) {
    if (even) {
        if (width <= 1) return;
        int32_t sn = (int32_t)((width + 1) >> 1);
        int32_t* tmp = (int32_t*)calloc((size_t)width, sizeof(int32_t));
        int32_t d1c, d1n, s1n, s0c, s0n;
        s1n = data[0];
        d1n = data[sn];
        s0n = s1n - ((d1n + 1) >> 1);
        int32_t i, j;
        for (i = 0, j = 1; i < (int32_t)width - 3; i += 2, j++) {
            d1c = d1n;
            s0c = s0n;
            s1n = data[j];
            d1n = data[sn + j];
            s0n = s1n - ((d1c + d1n + 2) >> 2);
            tmp[i] = s0c;
            tmp[i + 1] = d1c + ((s0c + s0n) >> 1);
        }
        tmp[i] = s0n;
        if ((width & 1) != 0) {
            tmp[width - 1] = data[(width - 1) / 2] - ((d1n + 1) >> 1);
            tmp[width - 2] = d1n + ((s0n + tmp[width - 1]) >> 1);
        } else {
            tmp[width - 1] = d1n + s0n;
        }
        memcpy(data, tmp, (size_t)width * sizeof(int32_t));
        free(tmp);
    } else {
        if (width == 1) { data[0] /= 2; return; } /* Go `/` truncates toward zero, like C */
        if (width <= 0) return;
        if (width == 2) {
            int32_t out1 = data[0] - ((data[1] + 1) >> 1);
            int32_t out0 = data[1] + out1;
            data[0] = out0;
            data[1] = out1;
            return;
        }
        int32_t sn = (int32_t)(width >> 1);
        int32_t* tmp = (int32_t*)calloc((size_t)width, sizeof(int32_t));
        int32_t s1, s2, dc, dn_var;
        s1 = data[sn + 1];
        dc = data[0] - ((data[sn] + s1 + 2) >> 2);
        tmp[0] = data[sn] + dc;
        int32_t i, j;
        int32_t not_odd = ((width & 1) == 0) ? 1 : 0;
        int32_t limit = (int32_t)width - 2 - not_odd;
        for (i = 1, j = 1; i < limit; i += 2, j++) {
            s2 = data[sn + j + 1];
            dn_var = data[j] - ((s1 + s2 + 2) >> 2);
            tmp[i] = dc;
            tmp[i + 1] = s1 + ((dn_var + dc) >> 1);
            dc = dn_var;
            s1 = s2;
        }
        tmp[i] = dc;
        if ((width & 1) == 0) {
            dn_var = data[width / 2 - 1] - ((s1 + 1) >> 1);
            tmp[width - 2] = s1 + ((dn_var + dc) >> 1);
            tmp[width - 1] = dn_var;
        } else {
            tmp[width - 1] = s1 + dc;
        }
        memcpy(data, tmp, (size_t)width * sizeof(int32_t));
        free(tmp);
    }
}

/* jpeg2000/wavelet/dwt53.go:259-301 (Forward53_2DWithParity): all columns, then all rows */
ORC_API void orc_fwd53_2d(int32_t* data, int width, int height, int stride, int even_row, int even_col) {
    if (width <= 1 && height <= 1) return;
    if (height > 1) {
        int32_t* col = (int32_t*)malloc((size_t)height * sizeof(int32_t));
        for (int x = 0; x < width; x++) {
            for (int y = 0; y < height; y++) col[y] = data[(size_t)y * stride + x];
            orc_fwd53_1d(col, height, even_col);
            for (int y = 0; y < height; y++) data[(size_t)y * stride + x] = col[y];
        }
        free(col);
    }
    if (width > 1) {
        int32_t* row = (int32_t*)malloc((size_t)width * sizeof(int32_t));
        for (int y = 0; y < height; y++) {
            for (int x = 0; x < width; x++) row[x] = data[(size_t)y * stride + x];
            orc_fwd53_1d(row, width, even_row);
            for (int x = 0; x < width; x++) data[(size_t)y * stride + x] = row[x];
        }
        free(row);
    }
}

/* jpeg2000/wavelet/dwt53.go:313-355 (Inverse53_2DWithParity): all rows, then all columns */
ORC_API void orc_inv53_2d(int32_t* data, int width, int height, int stride, int even_row, int even_col) {
    if (width <= 1 && height <= 1) return;
    if (width > 1) {
        int32_t* row = (int32_t*)malloc((size_t)width * sizeof(int32_t));
        for (int y = 0; y < height; y++) {
            for (int x = 0; x < width; x++) row[x] = data[(size_t)y * stride + x];
            orc_inv53_1d(row, width, even_row);
            for (int x = 0; x < width; x++) data[(size_t)y * stride + x] = row[x];
        }
        free(row);
    }
    if (height > 1) {
        int32_t* col = (int32_t*)malloc((size_t)height * sizeof(int32_t));
        for (int x = 0; x < width; x++) {
            for (int y = 0; y < height; y++) col[y] = data[(size_t)y * stride + x];
            orc_inv53_1d(col, height, even_col);
            for (int y = 0; y < height; y++) data[(size_t)y * stride + x] = col[y];
        }
        free(col);
    }
}

/* jpeg2000/wavelet/dwt53.go:365-394 (ForwardMultilevelWithParity) */
ORC_API void orc_fwd53_multilevel(int32_t* data, int width, int height, int levels, int x0, int y0) {
    int stride = width, cw = width, ch = height, cx = x0, cy = y0;
    for (int l = 0; l < levels; l++) {
        if (cw <= 1 && ch <= 1) break;
        orc_fwd53_2d(data, cw, ch, stride, is_even(cx), is_even(cy));
        next_lowpass_window(&cw, &ch, &cx, &cy);
    }
}

/* jpeg2000/wavelet/dwt53.go:404-434 (InverseMultilevelWithParity) */
ORC_API void orc_inv53_multilevel(int32_t* data, int width, int height, int levels, int x0, int y0) {
    if (levels < 0) levels = 0;
    int* lw = (int*)malloc((size_t)(levels + 1) * 4 * sizeof(int));
    int *lh = lw + (levels + 1), *lx = lh + (levels + 1), *ly = lx + (levels + 1);
    lw[0] = width; lh[0] = height; lx[0] = x0; ly[0] = y0;
    for (int i = 1; i <= levels; i++) {
        lw[i] = lw[i - 1]; lh[i] = lh[i - 1]; lx[i] = lx[i - 1]; ly[i] = ly[i - 1];
        next_lowpass_window(&lw[i], &lh[i], &lx[i], &ly[i]);
    }
    for (int l = levels - 1; l >= 0; l--)
        orc_inv53_2d(data, lw[l], lh[l], width, is_even(lx[l]), is_even(ly[l]));
    free(lw);
}

/* ------------------------------------------------------------------- dwt97.go */

/* jpeg2000/wavelet/dwt97.go:11-22.  The Go constants are float64 literals that are
 * converted to float32 at the use site (float32(c) at :99, float32(K97) at :207):
 * decimal -> float64 -> float32, which is what a C double literal cast to float does. */
static const double ALPHA97 = -1.586134342;
static const double BETA97 = -0.052980118;
static const double GAMMA97 = 0.882911075;
static const double DELTA97 = 0.443506852;
static const double K97 = 1.230174105;
static const double INVK97 = 0.812893066;
static const double TWOINVK97 = 1.625732422;

static int32_t min32(int32_t a, int32_t b) { return a < b ? a : b; }

/* jpeg2000/wavelet/dwt97.go:97-117 (encodeStep2_97Float32) and :269-287 (decodeStep2OpenJPEG97Float32) */
static void step2_97(float* data, int32_t fl_start, int32_t fw_start, int32_t end, int32_t m, float c32) {
    int32_t imax = min32(end, m);
    if (imax > 0) {
        int32_t fw = fw_start, fl = fl_start;
        data[fw - 1] += (data[fl] + data[fw]) * c32;
        fw += 2;
        for (int32_t i = 1; i < imax; i++) {
            data[fw - 1] += (data[fw - 2] + data[fw]) * c32;
            fw += 2;
        }
    }
    if (m < end) {
        int32_t fw = fw_start + 2 * m;
        data[fw - 1] += (2 * data[fw - 2]) * c32;
    }
}

/* jpeg2000/wavelet/dwt97.go:119-137 (encodeStep1Combined97Float32) */
static void step1_combined_97(float* data, int32_t iters_c1, int32_t iters_c2, float c1, float c2) {
    int32_t common = min32(iters_c1, iters_c2);
    int32_t i, fw = 0;
    for (i = 0; i < common; i++) {
        data[fw] *= c1;
        data[fw + 1] *= c2;
        fw += 2;
    }
    if (i < iters_c1) data[fw] *= c1;
    else if (i < iters_c2) data[fw + 1] *= c2;
}

/* jpeg2000/wavelet/dwt97.go:139-160 (deinterleaveH97Float32) */
static void deinterleave_97(float* data, int32_t dn, int32_t sn, int even) {
    int width = (int)(dn + sn);
    float* tmp = (float*)calloc((size_t)width, sizeof(float));
    if (even) {
        for (int32_t i = 0; i < sn; i++) tmp[i] = data[2 * i];
        for (int32_t i = 0; i < dn; i++) tmp[sn + i] = data[2 * i + 1];
    } else {
        for (int32_t i = 0; i < sn; i++) tmp[i] = data[2 * i + 1];
        for (int32_t i = 0; i < dn; i++) tmp[sn + i] = data[2 * i];
    }
    memcpy(data, tmp, (size_t)width * sizeof(float));
    free(tmp);
}

/* jpeg2000/wavelet/dwt97.go:225-246 (interleaveH97Float32) */
static void interleave_97(float* data, int32_t dn, int32_t sn, int even) {
    int width = (int)(dn + sn);
    float* tmp = (float*)calloc((size_t)width, sizeof(float));
    if (even) {
        for (int32_t i = 0; i < sn; i++) tmp[2 * i] = data[i];
        for (int32_t i = 0; i < dn; i++) tmp[2 * i + 1] = data[sn + i];
    } else {
        for (int32_t i = 0; i < sn; i++) tmp[2 * i + 1] = data[i];
        for (int32_t i = 0; i < dn; i++) tmp[2 * i] = data[sn + i];
    }
    memcpy(data, tmp, (size_t)width * sizeof(float));
    free(tmp);
}

/* jpeg2000/wavelet/dwt97.go:47-95 (Forward97_1DFloat32WithParity) */
ORC_API void orc_fwd97_1d(float* data, int width, int even) {
    if (width <= 1) return;
    int32_t sn, dn, a, b;
    if (even) { sn = (int32_t)((width + 1) >> 1); dn = (int32_t)width - sn; a = 0; b = 1; }
    else { sn = (int32_t)(width >> 1); dn = (int32_t)width - sn; a = 1; b = 0; }
    step2_97(data, a, b + 1, dn, min32(dn, sn - b), (float)ALPHA97);
    step2_97(data, b, a + 1, sn, min32(sn, dn - a), (float)BETA97);
    step2_97(data, a, b + 1, dn, min32(dn, sn - b), (float)GAMMA97);
    step2_97(data, b, a + 1, sn, min32(sn, dn - a), (float)DELTA97);
    if (a == 0) step1_combined_97(data, sn, dn, (float)INVK97, (float)K97);
    else step1_combined_97(data, dn, sn, (float)K97, (float)INVK97);
    deinterleave_97(data, dn, sn, even);
}

/* jpeg2000/wavelet/dwt97.go:263-267 (decodeStep1OpenJPEG97Float32) */
static void decode_step1_97(float* data, int32_t start, int32_t end, float c) {
    for (int32_t i = 0; i < end; i++) data[start + 2 * i] *= c;
}

/* jpeg2000/wavelet/dwt97.go:192-223 (Inverse97_1DOpenJPEGWithParity) */
ORC_API void orc_inv97_1d(float* data, int width, int even) {
    if (width <= 1) return;
    int32_t sn, dn, a, b;
    if (even) { sn = (int32_t)((width + 1) >> 1); dn = (int32_t)width - sn; a = 0; b = 1; }
    else { sn = (int32_t)(width >> 1); dn = (int32_t)width - sn; a = 1; b = 0; }
    interleave_97(data, dn, sn, even);
    decode_step1_97(data, a, sn, (float)K97);
    decode_step1_97(data, b, dn, (float)TWOINVK97);
    step2_97(data, b, a + 1, sn, min32(sn, dn - a), (float)(-DELTA97));
    step2_97(data, a, b + 1, dn, min32(dn, sn - b), (float)(-GAMMA97));
    step2_97(data, b, a + 1, sn, min32(sn, dn - a), (float)(-BETA97));
    step2_97(data, a, b + 1, dn, min32(dn, sn - b), (float)(-ALPHA97));
}

/* jpeg2000/wavelet/dwt97.go:290-322 (Forward97_2DFloat32WithParity): columns then rows */
ORC_API void orc_fwd97_2d(float* data, int width, int height, int stride, int even_row, int even_col) {
    if (width <= 1 && height <= 1) return;
    if (height > 1) {
        float* col = (float*)malloc((size_t)height * sizeof(float));
        for (int x = 0; x < width; x++) {
            for (int y = 0; y < height; y++) col[y] = data[(size_t)y * stride + x];
            orc_fwd97_1d(col, height, even_col);
            for (int y = 0; y < height; y++) data[(size_t)y * stride + x] = col[y];
        }
        free(col);
    }
    if (width > 1) {
        float* row = (float*)malloc((size_t)width * sizeof(float));
        for (int y = 0; y < height; y++) {
            for (int x = 0; x < width; x++) row[x] = data[(size_t)y * stride + x];
            orc_fwd97_1d(row, width, even_row);
            for (int x = 0; x < width; x++) data[(size_t)y * stride + x] = row[x];
        }
        free(row);
    }
}

/* jpeg2000/wavelet/dwt97.go:355-385 (Inverse97_2DOpenJPEGWithParity): rows then columns */
ORC_API void orc_inv97_2d(float* data, int width, int height, int stride, int even_row, int even_col) {
    if (width <= 1 && height <= 1) return;
    if (width > 1) {
        float* row = (float*)malloc((size_t)width * sizeof(float));
        for (int y = 0; y < height; y++) {
            for (int x = 0; x < width; x++) row[x] = data[(size_t)y * stride + x];
            orc_inv97_1d(row, width, even_row);
            for (int x = 0; x < width; x++) data[(size_t)y * stride + x] = row[x];
        }
        free(row);
    }
    if (height > 1) {
        float* col = (float*)malloc((size_t)height * sizeof(float));
        for (int x = 0; x < width; x++) {
            for (int y = 0; y < height; y++) col[y] = data[(size_t)y * stride + x];
            orc_inv97_1d(col, height, even_col);
            for (int y = 0; y < height; y++) data[(size_t)y * stride + x] = col[y];
        }
        free(col);
    }
}

/* jpeg2000/wavelet/dwt97.go:388-407 (ForwardMultilevel97Float32WithParity) */
ORC_API void orc_fwd97_multilevel(float* data, int width, int height, int levels, int x0, int y0) {
    int stride = width, cw = width, ch = height, cx = x0, cy = y0;
    for (int l = 0; l < levels; l++) {
        if (cw <= 1 && ch <= 1) break;
        orc_fwd97_2d(data, cw, ch, stride, is_even(cx), is_even(cy));
        next_lowpass_window(&cw, &ch, &cx, &cy);
    }
}

/* jpeg2000/wavelet/dwt97.go:425-451 (InverseMultilevel97OpenJPEGWithParity) */
ORC_API void orc_inv97_multilevel(float* data, int width, int height, int levels, int x0, int y0) {
    if (levels < 0) levels = 0;
    int* lw = (int*)malloc((size_t)(levels + 1) * 4 * sizeof(int));
    int *lh = lw + (levels + 1), *lx = lh + (levels + 1), *ly = lx + (levels + 1);
    lw[0] = width; lh[0] = height; lx[0] = x0; ly[0] = y0;
    for (int i = 1; i <= levels; i++) {
        lw[i] = lw[i - 1]; lh[i] = lh[i - 1]; lx[i] = lx[i - 1]; ly[i] = ly[i - 1];
        next_lowpass_window(&lw[i], &lh[i], &lx[i], &ly[i]);
    }
    for (int l = levels - 1; l >= 0; l--)
        orc_inv97_2d(data, lw[l], lh[l], width, is_even(lx[l]), is_even(ly[l]));
    free(lw);
}

/* jpeg2000/wavelet/dwt97.go:483-503 (roundFloat32ToNearestEven) */
static int64_t round_f32_nearest_even(float v) {
    int64_t i = (int64_t)v;
    double frac = (double)(v - (float)i);
    if (frac < 0) frac = -frac;
    if (frac > 0.5) return v >= 0 ? i + 1 : i - 1;
    if (frac < 0.5) return i;
    if (i % 2 == 0) return i;
    return v >= 0 ? i + 1 : i - 1;
}

/* jpeg2000/wavelet/dwt97.go:473-479 (ConvertFloat32ToInt32OpenJPEG) */
ORC_API void orc_convert_f32_to_i32(const float* in, int32_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = (int32_t)round_f32_nearest_even(in[i]);
}

/* jpeg2000/wavelet/dwt97.go:515-526 (ConvertFloat64ToInt32): half away from zero by +-0.5 and truncation */
ORC_API void orc_convert_f64_to_i32(const double* in, int32_t* out, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = in[i] >= 0 ? (int32_t)(in[i] + 0.5) : (int32_t)(in[i] - 0.5);
}

/* float64 wrappers, jpeg2000/wavelet/dwt97.go:30-44,181-187,325-351,410-421:
 * convert to float32, run the float32 routine, convert back. */
ORC_API void orc_fwd97_multilevel_f64(double* data, int width, int height, int levels, int x0, int y0) {
    size_t n = (size_t)width * height;
    float* f = (float*)malloc(n * sizeof(float));
    for (size_t i = 0; i < n; i++) f[i] = (float)data[i];
    orc_fwd97_multilevel(f, width, height, levels, x0, y0);
    for (size_t i = 0; i < n; i++) data[i] = (double)f[i];
    free(f);
}
ORC_API void orc_inv97_multilevel_f64(double* data, int width, int height, int levels, int x0, int y0) {
    size_t n = (size_t)width * height;
    float* f = (float*)malloc(n * sizeof(float));
    for (size_t i = 0; i < n; i++) f[i] = (float)data[i];
    orc_inv97_multilevel(f, width, height, levels, x0, y0);
    for (size_t i = 0; i < n; i++) data[i] = (double)f[i];
    free(f);
}

/* ----------------------------------------------------------------- colorspace */

/* jpeg2000/colorspace/rct.go:26-35 (ApplyRCTToComponents; RCTForward :6-11) */
ORC_API void orc_rct_forward(size_t n, const int32_t* r, const int32_t* g, const int32_t* b,
                             int32_t* y, int32_t* cb, int32_t* cr) {
    for (size_t i = 0; i < n; i++) {
        int32_t R = r[i], G = g[i], B = b[i];
        y[i] = (R + 2 * G + B) >> 2;
        cb[i] = B - G;
        cr[i] = R - G;
    }
}

/* jpeg2000/colorspace/rct.go:40-49 (ApplyInverseRCTToComponents; RCTInverse :16-21) */
ORC_API void orc_rct_inverse(size_t n, const int32_t* y, const int32_t* cb, const int32_t* cr,
                             int32_t* r, int32_t* g, int32_t* b) {
    for (size_t i = 0; i < n; i++) {
        int32_t Y = y[i], Cb = cb[i], Cr = cr[i];
        int32_t G = Y - ((Cb + Cr) >> 2);
        r[i] = Cr + G;
        g[i] = G;
        b[i] = Cb + G;
    }
}

/* Go math.Round: half away from zero; C round() has the same definition. */
static int32_t go_round_i32(double v) { return (int32_t)round(v); }

/* jpeg2000/colorspace/ict.go:24-34 (ApplyICTToComponents; ICTForward :8-14) */
ORC_API void orc_ict_forward(size_t n, const int32_t* r, const int32_t* g, const int32_t* b,
                             int32_t* y, int32_t* cb, int32_t* cr) {
    for (size_t i = 0; i < n; i++) {
        double R = (double)r[i], G = (double)g[i], B = (double)b[i];
        y[i] = go_round_i32(0.299 * R + 0.587 * G + 0.114 * B);
        cb[i] = go_round_i32(-0.16875 * R - 0.331260 * G + 0.5 * B);
        cr[i] = go_round_i32(0.5 * R - 0.41869 * G - 0.08131 * B);
    }
}

/* jpeg2000/colorspace/ict.go:36-45 (ApplyInverseICTToComponents; ICTInverse :16-21) */
ORC_API void orc_ict_inverse(size_t n, const int32_t* y, const int32_t* cb, const int32_t* cr,
                             int32_t* r, int32_t* g, int32_t* b) {
    for (size_t i = 0; i < n; i++) {
        double Y = (double)y[i], Cb = (double)cb[i], Cr = (double)cr[i];
        r[i] = go_round_i32(Y + 1.402 * Cr);
        g[i] = go_round_i32(Y - 0.34413 * Cb - 0.71414 * Cr);
        b[i] = go_round_i32(Y + 1.772 * Cb);
    }
}

/* jpeg2000/encoder.go:277-288 (applyOpenJPEGIrreversibleMCT): float32, result kept float32 */
ORC_API void orc_ict_forward_f32(size_t n, const int32_t* r, const int32_t* g, const int32_t* b,
                                 float* y, float* cb, float* cr) {
    for (size_t i = 0; i < n; i++) {
        float red = (float)r[i], green = (float)g[i], blue = (float)b[i];
        y[i] = (red * 0.299f + green * 0.587f) + blue * 0.114f;
        cb[i] = (red * -0.16875f + green * -0.331260f) + blue * 0.5f;
        cr[i] = (red * 0.5f + green * -0.41869f) + blue * -0.08131f;
    }
}

/* -------------------------------------------------------------- quantization.go */

/* jpeg2000/quantization.go:17-22 (dwtNorms97) */
static const double DWT_NORMS_97[4][10] = {
    {1.000, 1.965, 4.177, 8.403, 16.90, 33.84, 67.69, 135.3, 270.6, 540.9},
    {2.022, 3.989, 8.355, 17.04, 34.27, 68.63, 137.3, 274.6, 549.0, 0.0},
    {2.022, 3.989, 8.355, 17.04, 34.27, 68.63, 137.3, 274.6, 549.0, 0.0},
    {2.080, 3.865, 8.307, 17.18, 34.71, 69.59, 139.3, 278.6, 557.2, 0.0},
};

/* jpeg2000/quantization.go:39-52 (dwtNorm97) */
static double dwt_norm_97(int level, int orient) {
    if (level < 0) level = 0;
    if (orient == 0 && level >= 10) level = 9;
    else if (orient > 0 && level >= 9) level = 8;
    if (orient < 0 || orient > 3) return 1.0;
    return DWT_NORMS_97[orient][level];
}

/* jpeg2000/quantization.go:54-66 (qualityScale) */
static double quality_scale(int quality) {
    if (quality < 1) quality = 1;
    if (quality > 100) quality = 100;
    double scale = pow(2.0, (100.0 - (double)quality) / 12.5);
    if (scale < 0.01) scale = 0.01;
    return scale * 0.05;
}

/* jpeg2000/quantization.go:68-83 (subbandParams) */
static void subband_params(int idx, int num_levels, int* orient, int* level) {
    int resno;
    if (idx == 0) { resno = 0; *orient = 0; }
    else { resno = (idx - 1) / 3 + 1; *orient = (idx - 1) % 3 + 1; }
    *level = num_levels - resno;
    if (*level < 0) *level = 0;
}

/* jpeg2000/quantization.go:102-128 (encodeQuantizationStep) */
ORC_API uint16_t orc_encode_quant_step(double step_size, int numbps) {
    if (step_size <= 0) return 0;
    int32_t fixed = (int32_t)floor(step_size * 8192.0);
    if (fixed <= 0) fixed = 1;
    int log2v = 31 - __builtin_clz((uint32_t)fixed); /* bits.Len32(x)-1 */
    int p = log2v - 13;
    int n = 11 - log2v;
    int32_t mant;
    if (n < 0) mant = fixed >> -n; else mant = (int32_t)((uint32_t)fixed << n);
    mant &= 0x7ff;
    int expn = numbps - p;
    if (expn < 0) expn = 0;
    if (expn > 0x1f) expn = 0x1f;
    return (uint16_t)((expn << 11) | (int)mant);
}

/* jpeg2000/quantization.go:130-135 (decodeQuantizationStepWithGain) */
static double decode_quant_step_with_gain(uint16_t encoded, int bit_depth, int log2_gain) {
    int expn = (int)((encoded >> 11) & 0x1f);
    double mant = (double)(encoded & 0x7ff);
    int rb = bit_depth + log2_gain;
    return ldexp(1.0 + mant / 2048.0, rb - expn);
}

/* jpeg2000/quantization.go:140-154 (OpenJPEGRuntimeQuantizationSteps) */
ORC_API void orc_runtime_quant_steps(const uint16_t* encoded, int n, int num_levels, int bit_depth, double* steps) {
    for (int idx = 0; idx < n; idx++) {
        int orient, level;
        subband_params(idx, num_levels, &orient, &level);
        int log2_gain = 0;
        if (orient == 3) log2_gain = 2;
        else if (orient == 1 || orient == 2) log2_gain = 1;
        steps[idx] = (double)(float)decode_quant_step_with_gain(encoded[idx], bit_depth, log2_gain);
    }
}

/* jpeg2000/quantization.go:212-236 (CalculateOpenJPEGQuantizationParams) */
ORC_API int orc_openjpeg_quant_params(int num_levels, int bit_depth, uint16_t* encoded, double* step_sizes) {
    if (num_levels < 0) num_levels = 0;
    int nb = 3 * num_levels + 1;
    for (int bandno = 0; bandno < nb; bandno++) {
        int orient, level;
        subband_params(bandno, num_levels, &orient, &level);
        double norm = dwt_norm_97(level, orient);
        double stepsize = 1.0;
        if (norm > 0) stepsize = 1.0 / norm;
        step_sizes[bandno] = stepsize;
        encoded[bandno] = orc_encode_quant_step(stepsize, bit_depth);
    }
    return nb;
}

/* jpeg2000/quantization.go:180-208 (CalculateQuantizationParams) with calcOpenJPEGStepSizes97 :85-100 */
ORC_API int orc_quality_quant_params(int quality, int num_levels, int bit_depth, uint16_t* encoded, double* step_sizes) {
    if (quality < 1) quality = 1;
    if (quality > 100) quality = 100;
    double scale = quality_scale(quality);
    int nb;
    if (num_levels <= 0) {
        /* calcOpenJPEGStepSizes97 returns a single entry; the caller's slice has 3L+1 = 1 entries (L=0) */
        nb = 1;
        step_sizes[0] = scale;
    } else {
        nb = 3 * num_levels + 1;
        for (int idx = 0; idx < nb; idx++) {
            int orient, level;
            subband_params(idx, num_levels, &orient, &level);
            double norm = dwt_norm_97(level, orient);
            step_sizes[idx] = norm <= 0 ? scale : scale / norm;
        }
    }
    for (int i = 0; i < nb; i++) encoded[i] = orc_encode_quant_step(step_sizes[i], bit_depth);
    return nb;
}

/* jpeg2000/t2/tile_decoder.go:1018-1043 (decodeQuantizationSteps, style 2) with
 * log2GainForSubband :1048-1060 and decodeQuantStep :1062-1065 */
ORC_API void orc_decode_quant_steps(const uint16_t* encoded, int n, int num_levels, int bit_depth,
                                    int reversible, double* steps) {
    (void)num_levels;
    for (int idx = 0; idx < n; idx++) {
        int expn = (int)((encoded[idx] >> 11) & 0x1f);
        int mant = (int)(encoded[idx] & 0x7ff);
        int log2_gain = 0;
        if (reversible && idx != 0) {
            int orient = (idx - 1) % 3 + 1;
            log2_gain = (orient == 3) ? 2 : 1;
        }
        steps[idx] = ldexp(1.0 + (double)mant / 2048.0, (bit_depth + log2_gain) - expn);
    }
}

/* jpeg2000/t2/tile_decoder.go:1003-1017 (decodeQuantizationSteps, style 1 = scalar derived) */
ORC_API void orc_decode_quant_steps_derived(uint16_t encoded, int num_levels, int bit_depth, int reversible,
                                            double* steps) {
    int nb = 3 * num_levels + 1;
    int base_expn = (int)((encoded >> 11) & 0x1f);
    int base_mant = (int)(encoded & 0x7ff);
    for (int idx = 0; idx < nb; idx++) {
        int expn = base_expn;
        if (idx > 0) {
            expn -= (idx - 1) / 3;
            if (expn < 0) expn = 0;
        }
        int log2_gain = 0;
        if (reversible && idx != 0) {
            int orient = (idx - 1) % 3 + 1;
            log2_gain = (orient == 3) ? 2 : 1;
        }
        steps[idx] = ldexp(1.0 + (double)base_mant / 2048.0, (bit_depth + log2_gain) - expn);
    }
}

/* Go math.RoundToEven on float64 == rint() in the default rounding mode. */
static double round_to_even(double v) { return rint(v); }

/* jpeg2000/quantization.go:310-324 (QuantizeCoefficients) */
ORC_API void orc_quantize_coefficients(const int32_t* in, int32_t* out, size_t n, double step) {
    if (step <= 0) { memcpy(out, in, n * sizeof(int32_t)); return; }
    for (size_t i = 0; i < n; i++) out[i] = (int32_t)round_to_even((double)in[i] / step);
}

/* jpeg2000/quantization.go:326-340 (DequantizeCoefficients) */
ORC_API void orc_dequantize_coefficients(const int32_t* in, int32_t* out, size_t n, double step) {
    if (step <= 0) { memcpy(out, in, n * sizeof(int32_t)); return; }
    for (size_t i = 0; i < n; i++) out[i] = (int32_t)round_to_even((double)in[i] * step);
}

/* ---------------------------------------------------- band geometry (encoder/t2) */

typedef struct { int band, width, height, offx, offy; } band_info;

/* jpeg2000/encoder.go:2352-2370 / jpeg2000/t2/geometry.go:53-71 (resolutionDimsWithOrigin) */
static void resolution_dims_with_origin(int width, int height, int x0, int y0, int num_levels, int res,
                                        int* rw, int* rh) {
    int level_no = num_levels - res;
    if (level_no < 0) level_no = 0;
    int w = width, h = height, rx = x0, ry = y0;
    for (int i = 0; i < level_no; i++) {
        int lw = split_lengths(w, is_even(rx));
        int lh = split_lengths(h, is_even(ry));
        w = lw; h = lh;
        rx = next_coord(rx); ry = next_coord(ry);
    }
    *rw = w; *rh = h;
}

/* jpeg2000/encoder.go:2372-2389 / jpeg2000/t2/geometry.go:73-92 (bandInfosForResolution) */
static int band_infos_for_resolution(int width, int height, int x0, int y0, int num_levels, int res, band_info out[3]) {
    int rw, rh;
    resolution_dims_with_origin(width, height, x0, y0, num_levels, res, &rw, &rh);
    if (res == 0) {
        out[0].band = 0; out[0].width = rw; out[0].height = rh; out[0].offx = 0; out[0].offy = 0;
        return 1;
    }
    int lw, lh;
    resolution_dims_with_origin(width, height, x0, y0, num_levels, res - 1, &lw, &lh);
    int hw = rw - lw, hh = rh - lh;
    out[0] = (band_info){1, hw, lh, lw, 0};
    out[1] = (band_info){2, lw, hh, 0, lh};
    out[2] = (band_info){3, hw, hh, lw, lh};
    return 3;
}

/* Exposes the rectangles in QCD order for tests: rects[4*i..] = offx, offy, width, height. */
ORC_API int orc_band_rects(int width, int height, int x0, int y0, int num_levels, int32_t* rects) {
    int k = 0;
    band_info b[3];
    band_infos_for_resolution(width, height, x0, y0, num_levels, 0, b);
    rects[0] = b[0].offx; rects[1] = b[0].offy; rects[2] = b[0].width; rects[3] = b[0].height;
    k = 1;
    for (int res = 1; res <= num_levels; res++) {
        int nb = band_infos_for_resolution(width, height, x0, y0, num_levels, res, b);
        for (int j = 0; j < nb; j++, k++) {
            rects[4 * k] = b[j].offx; rects[4 * k + 1] = b[j].offy;
            rects[4 * k + 2] = b[j].width; rects[4 * k + 3] = b[j].height;
        }
    }
    return k;
}

/* jpeg2000/encoder.go:2311-2329 (quantizeSubbandFloat) */
static void quantize_subband_float(const float* coeffs, int32_t* out, size_t len, int x0, int y0, int w, int h,
                                   int stride, double step_size, int htj2k) {
    float scale = (float)(1 << 6); /* t1NMSEDecFracBits = 6, encoder.go:3347 */
    if (htj2k) scale = 1;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            size_t idx = (size_t)(y0 + y) * stride + (size_t)(x0 + x);
            if (idx < len) {
                if (step_size <= 0) {
                    out[idx] = (int32_t)round_to_even((double)coeffs[idx]);
                } else {
                    float quantized = (coeffs[idx] / (float)step_size) * scale;
                    out[idx] = (int32_t)round_to_even((double)quantized);
                }
            }
        }
    }
}

/* jpeg2000/encoder.go:2265-2302 (applyQuantizationBySubbandFloat) */
static void apply_quantization_by_subband_float(const float* coeffs, int32_t* quantized, int width, int height,
                                                int x0, int y0, int num_levels, const double* steps, int n_steps,
                                                int htj2k) {
    size_t len = (size_t)width * height;
    if (n_steps == 0 || num_levels == 0) {
        for (size_t i = 0; i < len; i++) quantized[i] = (int32_t)round_to_even((double)coeffs[i]);
        return;
    }
    memset(quantized, 0, len * sizeof(int32_t));
    int sb = 0;
    band_info b[3];
    band_infos_for_resolution(width, height, x0, y0, num_levels, 0, b);
    if (sb < n_steps && b[0].width > 0 && b[0].height > 0)
        quantize_subband_float(coeffs, quantized, len, b[0].offx, b[0].offy, b[0].width, b[0].height, width, steps[sb], htj2k);
    sb++;
    for (int res = 1; res <= num_levels; res++) {
        int nb = band_infos_for_resolution(width, height, x0, y0, num_levels, res, b);
        for (int j = 0; j < nb; j++) {
            if (sb < n_steps && b[j].width > 0 && b[j].height > 0)
                quantize_subband_float(coeffs, quantized, len, b[j].offx, b[j].offy, b[j].width, b[j].height, width, steps[sb], htj2k);
            sb++;
        }
    }
}

/* jpeg2000/t2/tile_decoder.go:970-987 (dequantizeSubbandFloat) */
static void dequantize_subband_float(float* data, size_t len, int x0, int y0, int w, int h, int stride,
                                     double step_size, int htj2k) {
    if (step_size <= 0) return;
    double scale = 0.5 * step_size;
    if (htj2k) scale = step_size;
    for (int y = 0; y < h; y++) {
        for (int x = 0; x < w; x++) {
            size_t idx = (size_t)(y0 + y) * stride + (size_t)(x0 + x);
            if (idx < len) data[idx] *= (float)scale;
        }
    }
}

/* jpeg2000/t2/tile_decoder.go:925-962 (applyDequantizationBySubbandFloat) */
static void apply_dequantization_by_subband_float(const int32_t* coeffs, float* out, int width, int height,
                                                  int num_levels, int x0, int y0, const double* steps, int n_steps,
                                                  int htj2k) {
    size_t len = (size_t)width * height;
    for (size_t i = 0; i < len; i++) out[i] = (float)coeffs[i];
    if (n_steps == 0) return;
    int sb = 0;
    band_info b[3];
    band_infos_for_resolution(width, height, x0, y0, num_levels, 0, b);
    if (sb < n_steps && b[0].width > 0 && b[0].height > 0)
        dequantize_subband_float(out, len, b[0].offx, b[0].offy, b[0].width, b[0].height, width, steps[sb], htj2k);
    sb++;
    for (int res = 1; res <= num_levels; res++) {
        int nb = band_infos_for_resolution(width, height, x0, y0, num_levels, res, b);
        for (int j = 0; j < nb; j++) {
            if (sb < n_steps && b[j].width > 0 && b[j].height > 0)
                dequantize_subband_float(out, len, b[j].offx, b[j].offy, b[j].width, b[j].height, width, steps[sb], htj2k);
            sb++;
        }
    }
}

/* ------------------------------------------------------------- encoder pipeline */

/* jpeg2000/encoder.go:341-383 (convertPixelData) */
static void convert_pixel_data(const j2k_fwd_params* p, const uint8_t* px, int32_t** data) {
    size_t np = (size_t)p->width * p->height;
    int C = p->components;
    if (p->bit_depth <= 8) {
        for (size_t i = 0; i < np; i++)
            for (int c = 0; c < C; c++) {
                int32_t val = (int32_t)px[i * C + c];
                if (p->is_signed && val >= 128) val -= 256;
                data[c][i] = val;
            }
    } else {
        for (size_t i = 0; i < np; i++)
            for (int c = 0; c < C; c++) {
                size_t idx = (i * C + c) * 2;
                int32_t val = (int32_t)px[idx] | ((int32_t)px[idx + 1] << 8);
                if (p->is_signed && val >= (1 << (p->bit_depth - 1))) val -= (1 << p->bit_depth);
                data[c][i] = val;
            }
    }
}

/* jpeg2000/encoder.go:3698-3711 (applyDCLevelShift) */
static void apply_dc_level_shift(const j2k_fwd_params* p, int32_t** data) {
    if (p->is_signed) return;
    int32_t shift = (int32_t)(1 << (p->bit_depth - 1));
    size_t np = (size_t)p->width * p->height;
    for (int c = 0; c < p->components; c++)
        for (size_t i = 0; i < np; i++) data[c][i] -= shift;
}

/* jpeg2000/encoder.go:662-665 (mctFixedMul) */
static int32_t mct_fixed_mul(int32_t a, int32_t b) {
    int64_t temp = (int64_t)a * (int64_t)b + 4096;
    return (int32_t)(temp >> 13);
}

/* jpeg2000/encoder.go:465-525 (applyCustomMCT); `q13` selects the else-branch at :506 */
static void apply_custom_mct(const j2k_fwd_params* p, int32_t** data, int q13) {
    int C = p->components;
    size_t n = (size_t)p->width * p->height;
    int32_t* out[J2K_MAX_COMPONENTS];
    for (int c = 0; c < C; c++) out[c] = (int32_t*)malloc(n * sizeof(int32_t));
    if (p->mct_has_offsets)
        for (int c = 0; c < C; c++) {
            int32_t off = p->mct_offsets[c];
            if (off == 0) continue;
            for (size_t i = 0; i < n; i++) data[c][i] -= off;
        }
    int32_t m[J2K_MAX_COMPONENTS][J2K_MAX_COMPONENTS];
    if (!q13) {
        for (int r = 0; r < C; r++) for (int k = 0; k < C; k++) m[r][k] = (int32_t)p->mct_matrix[r * C + k];
        for (size_t i = 0; i < n; i++)
            for (int r = 0; r < C; r++) {
                int64_t sum = 0;
                for (int k = 0; k < C; k++) sum += (int64_t)m[r][k] * (int64_t)data[k][i];
                out[r][i] = (int32_t)sum;
            }
    } else {
        for (int r = 0; r < C; r++) for (int k = 0; k < C; k++) m[r][k] = (int32_t)(p->mct_matrix[r * C + k] * (double)(1 << 13));
        for (size_t i = 0; i < n; i++)
            for (int r = 0; r < C; r++) {
                int32_t sum = 0;
                for (int k = 0; k < C; k++) sum = (int32_t)((uint32_t)sum + (uint32_t)mct_fixed_mul(m[r][k], data[k][i]));
                out[r][i] = sum;
            }
    }
    for (int c = 0; c < C; c++) { memcpy(data[c], out[c], n * sizeof(int32_t)); free(out[c]); }
}

/* jpeg2000/encoder.go:549-660 (applyMCTBinding and helpers) */
static void apply_mct_binding_fwd(const j2k_mct_binding* b, int32_t** data, size_t n, int components) {
    int idx[J2K_MAX_COMPONENTS];
    int nc = b->n_components;
    if (nc == 0 && components > 0) { nc = components; for (int i = 0; i < nc; i++) idx[i] = i; }
    else for (int i = 0; i < nc; i++) idx[i] = b->component_ids[i];
    if (nc == 0) return;
    if (b->has_offsets) /* applyMCTOffsets :582-594 */
        for (int k = 0; k < nc; k++) {
            int32_t off = b->offsets[k];
            if (off == 0) continue;
            for (size_t i = 0; i < n; i++) data[idx[k]][i] -= off;
        }
    double mat[J2K_MAX_COMPONENTS][J2K_MAX_COMPONENTS];
    for (int r = 0; r < nc; r++) for (int k = 0; k < nc; k++) /* prepareTransformMatrix :596-608 */
        mat[r][k] = b->has_matrix ? b->matrix[r * nc + k] : (r == k ? 1.0 : 0.0);
    int32_t im[J2K_MAX_COMPONENTS][J2K_MAX_COMPONENTS];
    if (b->element_type == 0) { /* applyIntegerMatrixTransform :610-633 */
        for (int r = 0; r < nc; r++) for (int k = 0; k < nc; k++) im[r][k] = (int32_t)mat[r][k];
        for (size_t i = 0; i < n; i++) {
            int32_t out[J2K_MAX_COMPONENTS];
            for (int r = 0; r < nc; r++) {
                int64_t sum = 0;
                for (int k = 0; k < nc; k++) sum += (int64_t)im[r][k] * (int64_t)data[idx[k]][i];
                out[r] = (int32_t)sum;
            }
            for (int r = 0; r < nc; r++) data[idx[r]][i] = out[r];
        }
    } else { /* applyFixedPointMatrixTransform :635-660 */
        for (int r = 0; r < nc; r++) for (int k = 0; k < nc; k++) im[r][k] = (int32_t)(mat[r][k] * (double)(1 << 13));
        for (size_t i = 0; i < n; i++) {
            int32_t out[J2K_MAX_COMPONENTS];
            for (int r = 0; r < nc; r++) {
                int32_t sum = 0;
                for (int k = 0; k < nc; k++) sum = (int32_t)((uint32_t)sum + (uint32_t)mct_fixed_mul(im[r][k], data[idx[k]][i]));
                out[r] = sum;
            }
            for (int r = 0; r < nc; r++) data[idx[r]][i] = out[r];
        }
    }
}

/* jpeg2000/encoder.go:1966-1983 (tileBounds) + tile grid :1990-1997 */
ORC_API int orc_fwd_tile_bounds(const j2k_fwd_params* p, int idx, int32_t b[4]) {
    int tw = p->tile_width ? p->tile_width : p->width;
    int th = p->tile_height ? p->tile_height : p->height;
    int ntx = (p->width + tw - 1) / tw, nty = (p->height + th - 1) / th;
    if (idx >= 0 && idx < ntx * nty) {
        int tx = idx % ntx, ty = idx / ntx;
        int x0 = tx * tw, y0 = ty * th, x1 = x0 + tw, y1 = y0 + th;
        if (x1 > p->width) x1 = p->width;
        if (y1 > p->height) y1 = p->height;
        b[0] = x0; b[1] = y0; b[2] = x1; b[3] = y1;
    }
    return ntx * nty;
}

/* The forward path for one frame.  Follows Encoder.Encode (jpeg2000/encoder.go:180-218) up to
 * buildCodestream, then for every tile Encoder.transformTile (:2213-2237) with
 * applyWaveletTransform (:2187-2211) / applyIrreversibleWaveletTransform (:2239-2259).
 * `planes` non-NULL selects the EncodeComponents entry (:221-273). */
static int forward_impl(const j2k_fwd_params* p, const uint8_t* pixels, const int32_t* const* planes, int32_t* coeffs_out) {
    int C = p->components;
    size_t np = (size_t)p->width * p->height;
    int32_t* data[J2K_MAX_COMPONENTS] = {0};
    float* fdata[J2K_MAX_COMPONENTS] = {0};
    for (int c = 0; c < C; c++) data[c] = (int32_t*)malloc(np * sizeof(int32_t));
    if (planes) for (int c = 0; c < C; c++) memcpy(data[c], planes[c], np * sizeof(int32_t));
    else convert_pixel_data(p, pixels, data);
    apply_dc_level_shift(p, data);
    int have_f = 0;
    switch (p->mct_mode) { /* dispatch :196-209 (the shim maps EnableMCT & friends onto mct_mode) */
    case J2K_MCT_BINDINGS:
        for (int i = 0; i < p->n_bindings; i++) apply_mct_binding_fwd(&p->bindings[i], data, np, C);
        break;
    case J2K_MCT_CUSTOM_INT: apply_custom_mct(p, data, 0); break;
    case J2K_MCT_CUSTOM_Q13: apply_custom_mct(p, data, 1); break;
    case J2K_MCT_RCT: {
        int32_t* y = (int32_t*)malloc(np * 4); int32_t* cb = (int32_t*)malloc(np * 4); int32_t* cr = (int32_t*)malloc(np * 4);
        orc_rct_forward(np, data[0], data[1], data[2], y, cb, cr);
        free(data[0]); free(data[1]); free(data[2]);
        data[0] = y; data[1] = cb; data[2] = cr;
        break;
    }
    case J2K_MCT_ICT:
        for (int c = 0; c < 3; c++) fdata[c] = (float*)malloc(np * sizeof(float));
        orc_ict_forward_f32(np, data[0], data[1], data[2], fdata[0], fdata[1], fdata[2]);
        have_f = 1;
        break;
    default: break;
    }
    int32_t tb[4];
    int ntiles = orc_fwd_tile_bounds(p, 0, tb);
    size_t out_off = 0;
    for (int t = 0; t < ntiles; t++) {
        orc_fwd_tile_bounds(p, t, tb);
        int x0 = tb[0], y0 = tb[1], tw = tb[2] - tb[0], th = tb[3] - tb[1];
        size_t tn = (size_t)tw * th;
        for (int c = 0; c < C; c++) {
            int32_t* out = coeffs_out + out_off;
            out_off += tn;
            if (p->reversible) {
                /* transformTile int branch :2227-2236 + applyWaveletTransform :2194-2205 */
                for (int ty = 0; ty < th; ty++)
                    memcpy(out + (size_t)ty * tw, data[c] + (size_t)(y0 + ty) * p->width + x0, (size_t)tw * 4);
                if (p->num_levels > 0) orc_fwd53_multilevel(out, tw, th, p->num_levels, x0, y0);
                if (p->fuse_t1_shift && !p->htj2k) /* encodeCodeBlock :3294-3300 */
                    for (size_t i = 0; i < tn; i++) out[i] = (int32_t)((uint32_t)out[i] << 6);
            } else {
                float* tile = (float*)malloc(tn * sizeof(float));
                if (have_f && c < 3) { /* :2214-2224 */
                    for (int ty = 0; ty < th; ty++)
                        memcpy(tile + (size_t)ty * tw, fdata[c] + (size_t)(y0 + ty) * p->width + x0, (size_t)tw * 4);
                } else { /* :2227-2236 then ConvertInt32ToFloat32 :2206-2209 */
                    for (int ty = 0; ty < th; ty++)
                        for (int tx = 0; tx < tw; tx++)
                            tile[(size_t)ty * tw + tx] = (float)data[c][(size_t)(y0 + ty) * p->width + x0 + tx];
                }
                if (p->num_levels == 0) { /* :2240-2245 (and :2188-2191 for the int branch: identical values) */
                    orc_convert_f32_to_i32(tile, out, tn);
                } else {
                    orc_fwd97_multilevel(tile, tw, th, p->num_levels, x0, y0);
                    apply_quantization_by_subband_float(tile, out, tw, th, x0, y0, p->num_levels, p->steps, p->n_steps, p->htj2k);
                }
                free(tile);
            }
        }
    }
    for (int c = 0; c < C; c++) { free(data[c]); free(fdata[c]); }
    return 0;
}

ORC_API int orc_forward(const j2k_fwd_params* p, const void* pixels, int32_t* coeffs_out) {
    return forward_impl(p, (const uint8_t*)pixels, NULL, coeffs_out);
}
ORC_API int orc_forward_planar(const j2k_fwd_params* p, const int32_t* const* planes, int32_t* coeffs_out) {
    return forward_impl(p, NULL, planes, coeffs_out);
}

/* ------------------------------------------------------------- decoder pipeline */

static int ceil_div(int a, int b) { /* jpeg2000/tile_assembler.go:207-215 */
    if (b <= 0) return 0;
    if (a >= 0) return (a + b - 1) / b;
    return a / b;
}

/* jpeg2000/tile_assembler.go:33-101 (NewTileLayout, GetTileBounds): image-local bounds.
 * canvas[2] (optional) receives the tile's canvas origin used for DWT parity
 * (jpeg2000/t2/tile_decoder.go:269-294,336-350). */
ORC_API int orc_inv_tile_bounds(const j2k_inv_params* p, int idx, int32_t b[4], int32_t canvas[2]) {
    int ntx = ceil_div(p->xsiz - p->xtosiz, p->xtsiz);
    int nty = ceil_div(p->ysiz - p->ytosiz, p->ytsiz);
    if (idx >= 0 && idx < ntx * nty) {
        int tx = idx % ntx, ty = idx / ntx;
        int gx0 = tx * p->xtsiz + p->xtosiz, gy0 = ty * p->ytsiz + p->ytosiz;
        int gx1 = gx0 + p->xtsiz, gy1 = gy0 + p->ytsiz;
        if (gx0 < p->xosiz) gx0 = p->xosiz;
        if (gy0 < p->yosiz) gy0 = p->yosiz;
        if (gx1 > p->xsiz) gx1 = p->xsiz;
        if (gy1 > p->ysiz) gy1 = p->ysiz;
        b[0] = gx0 - p->xosiz; b[1] = gy0 - p->yosiz; b[2] = gx1 - p->xosiz; b[3] = gy1 - p->yosiz;
        if (canvas) { canvas[0] = gx0; canvas[1] = gy0; }
    }
    return ntx * nty;
}

/* jpeg2000/decoder.go:630-694 (applyDecoderMCTBindings and helpers) */
static void apply_mct_binding_inv(const j2k_mct_binding* b, int32_t** data, size_t n) {
    int nc = b->n_components;
    if (nc == 0) return; /* :633-635 */
    const int32_t* ids = b->component_ids;
    if (b->has_matrix) {
        if (b->element_type == 0) { /* applyIntegerMatrixTransform :646-662 */
            int32_t im[J2K_MAX_COMPONENTS][J2K_MAX_COMPONENTS];
            for (int r = 0; r < nc; r++) for (int k = 0; k < nc; k++) im[r][k] = (int32_t)b->matrix[r * nc + k];
            for (size_t i = 0; i < n; i++) {
                int32_t out[J2K_MAX_COMPONENTS];
                for (int r = 0; r < nc; r++) {
                    int64_t sum = 0;
                    for (int k = 0; k < nc; k++) sum += (int64_t)im[r][k] * (int64_t)data[ids[k]][i];
                    out[r] = (int32_t)sum;
                }
                for (int r = 0; r < nc; r++) data[ids[r]][i] = out[r];
            }
        } else { /* applyFloatMatrixTransform :664-681 */
            for (size_t i = 0; i < n; i++) {
                int32_t out[J2K_MAX_COMPONENTS];
                for (int r = 0; r < nc; r++) {
                    double sum = 0.0;
                    for (int k = 0; k < nc; k++) sum += b->matrix[r * nc + k] * (double)data[ids[k]][i];
                    out[r] = go_round_i32(sum);
                }
                for (int r = 0; r < nc; r++) data[ids[r]][i] = out[r];
            }
        }
    }
    if (b->has_offsets) /* applyBindingOffsets :683-694 */
        for (int k = 0; k < nc; k++) {
            int32_t off = b->offsets[k];
            if (off != 0) for (size_t i = 0; i < n; i++) data[ids[k]][i] += off;
        }
}

/* jpeg2000/decoder.go:696-723 (applyDecoderInverseCustomMCT) */
static void apply_inverse_custom_mct(const j2k_inv_params* p, int32_t** data, size_t n) {
    int C = p->components;
    int32_t* out[J2K_MAX_COMPONENTS];
    for (int c = 0; c < C; c++) out[c] = (int32_t*)malloc(n * sizeof(int32_t));
    for (size_t i = 0; i < n; i++)
        for (int r = 0; r < C; r++) {
            double sum = 0.0;
            for (int k = 0; k < C; k++) sum += p->mct_matrix[r * C + k] * (double)data[k][i];
            out[r][i] = go_round_i32(sum);
        }
    for (int c = 0; c < C; c++) { memcpy(data[c], out[c], n * sizeof(int32_t)); free(out[c]); }
    if (p->mct_has_offsets)
        for (int c = 0; c < C; c++) {
            int32_t off = p->mct_offsets[c];
            if (off != 0) for (size_t i = 0; i < n; i++) data[c][i] += off;
        }
}

/* jpeg2000/decoder.go:777-944 (GetPixelData: getGrayscalePixelData / getInterleavedPixelData) */
static void get_pixel_data(const j2k_inv_params* p, int32_t** data, int width, int height, uint8_t* result) {
    size_t np = (size_t)width * height;
    int C = p->components, B = p->bit_depth;
    for (size_t i = 0; i < np; i++)
        for (int c = 0; c < C; c++) {
            int32_t val = data[c][i];
            if (p->is_signed) {
                int32_t min_val = -(1 << (B - 1)), max_val = (1 << (B - 1)) - 1;
                if (val < min_val) val = min_val; else if (val > max_val) val = max_val;
                if (val < 0) val += (1 << B);
            } else {
                if (val < 0) val = 0;
                int32_t max_val = (1 << B) - 1;
                if (val > max_val) val = max_val;
            }
            if (B <= 8) result[i * C + c] = (uint8_t)val;
            else {
                size_t idx = (i * C + c) * 2;
                result[idx] = (uint8_t)val;
                result[idx + 1] = (uint8_t)(val >> 8);
            }
        }
}

/* The inverse path for one frame: for every tile-component TileDecoder.applyIDWT
 * (jpeg2000/t2/tile_decoder.go:886-919), TileAssembler.AssembleTile (jpeg2000/tile_assembler.go:138-178),
 * then Decoder.applyInverseTransforms / applyInverseDCLevelShift (jpeg2000/decoder.go:540-542) and
 * GetPixelData (:777).  planes_out (optional) = GetImageData. */
ORC_API int orc_inverse(const j2k_inv_params* p, const int32_t* coeffs_in, void* pixels_out, int32_t* planes_out) {
    int C = p->components;
    int iw = p->xsiz - p->xosiz, ih = p->ysiz - p->yosiz;
    size_t np = (size_t)iw * ih;
    int32_t* data[J2K_MAX_COMPONENTS] = {0};
    for (int c = 0; c < C; c++) data[c] = (int32_t*)calloc(np, sizeof(int32_t));
    int32_t tb[4], cv[2];
    int ntiles = orc_inv_tile_bounds(p, 0, tb, cv);
    size_t in_off = 0;
    for (int t = 0; t < ntiles; t++) {
        orc_inv_tile_bounds(p, t, tb, cv);
        int tw = tb[2] - tb[0], th = tb[3] - tb[1];
        if (tw < 0) tw = 0;
        if (th < 0) th = 0;
        size_t tn = (size_t)tw * th;
        for (int c = 0; c < C; c++) {
            const int32_t* coeffs = coeffs_in + in_off;
            in_off += tn;
            int32_t* samples = (int32_t*)malloc((tn ? tn : 1) * sizeof(int32_t));
            if (p->num_levels == 0 || p->reversible) { /* :887-891, :894-898 */
                memcpy(samples, coeffs, tn * 4);
                if (p->reversible && p->fuse_t1_halve && !p->htj2k) /* normalizeOpenJPEGReversibleT1Coefficients :989-993 (decodeCodeBlock :732-734) */
                    for (size_t i = 0; i < tn; i++) samples[i] /= 2;
                if (p->num_levels > 0) orc_inv53_multilevel(samples, tw, th, p->num_levels, cv[0], cv[1]);
            } else { /* :899-913 */
                float* f = (float*)malloc((tn ? tn : 1) * sizeof(float));
                apply_dequantization_by_subband_float(coeffs, f, tw, th, p->num_levels, cv[0], cv[1], p->steps, p->n_steps, p->htj2k);
                orc_inv97_multilevel(f, tw, th, p->num_levels, cv[0], cv[1]);
                orc_convert_f32_to_i32(f, samples, tn);
                free(f);
            }
            for (int ty = 0; ty < th; ty++) /* AssembleTile :164-175 */
                memcpy(data[c] + (size_t)(tb[1] + ty) * iw + tb[0], samples + (size_t)ty * tw, (size_t)tw * 4);
            free(samples);
        }
    }
    switch (p->mct_mode) { /* applyInverseTransforms :620-628 */
    case J2K_MCT_BINDINGS:
        for (int i = 0; i < p->n_bindings; i++) apply_mct_binding_inv(&p->bindings[i], data, np);
        break;
    case J2K_MCT_CUSTOM_FLOAT: apply_inverse_custom_mct(p, data, np); break;
    case J2K_MCT_RCT: orc_rct_inverse(np, data[0], data[1], data[2], data[0], data[1], data[2]); break;
    case J2K_MCT_ICT: orc_ict_inverse(np, data[0], data[1], data[2], data[0], data[1], data[2]); break;
    default: break;
    }
    if (!p->is_signed) { /* applyInverseDCLevelShift :948-962 */
        int32_t shift = (int32_t)(1 << (p->bit_depth - 1));
        for (int c = 0; c < C; c++) for (size_t i = 0; i < np; i++) data[c][i] += shift;
    }
    if (planes_out) for (int c = 0; c < C; c++) memcpy(planes_out + (size_t)c * np, data[c], np * 4);
    if (pixels_out) get_pixel_data(p, data, iw, ih, (uint8_t*)pixels_out);
    for (int c = 0; c < C; c++) free(data[c]);
    return 0;
}

/* Frame-parallel drivers for the CPU baseline (bench.py): `threads` > 1 runs one frame per
 * worker thread ("goroutine-per-frame upper bound"); the reference itself is single-goroutine
 * (jpeg2000/lossless/codec.go:246-261), which is threads == 1. */
typedef struct {
    int is_fwd, nframes, next;
    pthread_mutex_t mu;
    const j2k_fwd_params* fp; const j2k_inv_params* ip;
    const uint8_t* in8; const int32_t* in32; uint8_t* out8; int32_t* out32;
    size_t frame_stride, nc;
} batch_job;

static void* batch_worker(void* arg) {
    batch_job* j = (batch_job*)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int f = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (f >= j->nframes) break;
        if (j->is_fwd) forward_impl(j->fp, j->in8 + (size_t)f * j->frame_stride, NULL, j->out32 + (size_t)f * j->nc);
        else orc_inverse(j->ip, j->in32 + (size_t)f * j->nc, j->out8 + (size_t)f * j->frame_stride, NULL);
    }
    return NULL;
}

static int run_batch(batch_job* j, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (threads > j->nframes) threads = j->nframes > 0 ? j->nframes : 1;
    pthread_mutex_init(&j->mu, NULL);
    j->next = 0;
    if (threads == 1) { batch_worker(j); }
    else {
        pthread_t th[256];
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, batch_worker, j);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    }
    pthread_mutex_destroy(&j->mu);
    return 0;
}

ORC_API int orc_forward_batch(const j2k_fwd_params* p, int nframes, const void* pixels, size_t frame_stride,
                              int32_t* coeffs_out, int threads) {
    batch_job j; memset(&j, 0, sizeof j);
    j.is_fwd = 1; j.nframes = nframes; j.fp = p; j.in8 = (const uint8_t*)pixels; j.out32 = coeffs_out;
    j.frame_stride = frame_stride; j.nc = (size_t)p->width * p->height * p->components;
    return run_batch(&j, threads);
}

ORC_API int orc_inverse_batch(const j2k_inv_params* p, int nframes, const int32_t* coeffs_in, void* pixels_out,
                              size_t frame_stride, int threads) {
    batch_job j; memset(&j, 0, sizeof j);
    j.is_fwd = 0; j.nframes = nframes; j.ip = p; j.in32 = coeffs_in; j.out8 = (uint8_t*)pixels_out;
    j.frame_stride = frame_stride; j.nc = (size_t)(p->xsiz - p->xosiz) * (p->ysiz - p->yosiz) * p->components;
    return run_batch(&j, threads);
}

/* ------------------------------------------------------------------ code-block interface
 * jpeg2000/encoder.go:3059-3197 (getSubbandsForResolution), :3215-3285 (partitionIntoCodeBlocks),
 * :3294-3300 (T1 scaling), :3349-3362 + :3643-3667 (codeBlockNumBps / calculateMaxBitplane),
 * consumed in buildTilePacketEncoder :2424-2431 (res 0..L, sub-bands in order, blocks in order). */

typedef struct { int32_t* data; int x0, y0, width, height, band, res; } orc_subband;

static int orc_ceildivpow2(int a, int b) { return (a + (1 << b) - 1) >> b; }

/* getSubbandsForResolution: returns the number of sub-bands (1 or 3), each with a freshly extracted copy */
static int get_subbands_for_resolution(const int32_t* data, size_t len, int width, int height, int num_levels, int resolution,
                                       orc_subband sb[3]) {
    if (resolution == 0) {
        int divisor = 1 << num_levels;
        int llw = (width + divisor - 1) / divisor, llh = (height + divisor - 1) / divisor;
        int32_t* d = (int32_t*)calloc((size_t)llw * llh + 1, sizeof(int32_t));
        for (int y = 0; y < llh; y++)
            for (int x = 0; x < llw; x++) {
                size_t src = (size_t)y * width + x;
                if (src < len && y < height && x < width) d[(size_t)y * llw + x] = data[src];
            }
        sb[0] = (orc_subband){d, 0, 0, llw, llh, 0, 0};
        return 1;
    }
    int level = num_levels - resolution;
    if (level < 0) level = 0;
    int llw = orc_ceildivpow2(width - (0 << level), level + 1), llh = orc_ceildivpow2(height - (0 << level), level + 1);
    for (int b = 1; b <= 3; b++) {
        int x0b = b & 1, y0b = b >> 1;
        int bw = orc_ceildivpow2(width - (x0b << level), level + 1), bh = orc_ceildivpow2(height - (y0b << level), level + 1);
        int ox = x0b ? llw : 0, oy = y0b ? llh : 0;
        int32_t* d = (int32_t*)calloc((size_t)(bw > 0 ? bw : 0) * (bh > 0 ? bh : 0) + 1, sizeof(int32_t));
        for (int y = 0; y < bh; y++)
            for (int x = 0; x < bw; x++) {
                size_t src = (size_t)(oy + y) * width + (ox + x);
                if (src < len && oy + y < height && ox + x < width) d[(size_t)y * bw + x] = data[src];
            }
        sb[b - 1] = (orc_subband){d, ox, oy, bw, bh, b, resolution};
    }
    return 3;
}

/* calculateMaxBitplane :3643-3667 */
static int calculate_max_bitplane(const int32_t* data, size_t n) {
    int32_t max_abs = 0;
    for (size_t i = 0; i < n; i++) {
        int32_t a = data[i];
        if (a < 0) a = (int32_t)(0u - (uint32_t)a);
        if (a > max_abs) max_abs = a;
    }
    if (max_abs == 0) return -1;
    int bitplane = 0;
    while (max_abs > 0) { max_abs >>= 1; bitplane++; }
    return bitplane - 1;
}

/* Walks one tile-component plane the way buildTilePacketEncoder does.  table / blocks / numbps may be NULL.
 * Returns the number of code-blocks. */
static int walk_code_blocks(const int32_t* plane, int width, int height, int num_levels, int cbw, int cbh, int shift6, int htj2k,
                            j2k_cblk* table, int max_blocks, int32_t* blocks, int32_t* numbps) {
    int n = 0;
    int64_t off = 0;
    size_t len = (size_t)width * height;
    for (int res = 0; res <= num_levels; res++) {
        orc_subband sb[3];
        int nsb;
        if (plane) nsb = get_subbands_for_resolution(plane, len, width, height, num_levels, res, sb);
        else { /* geometry only: same formulas, no data */
            int32_t* zero = (int32_t*)calloc(len + 1, sizeof(int32_t));
            nsb = get_subbands_for_resolution(zero, len, width, height, num_levels, res, sb);
            free(zero);
        }
        for (int s = 0; s < nsb; s++) {
            const orc_subband* b = &sb[s];
            int ncbx = b->width > 0 ? (b->width + cbw - 1) / cbw : 0, ncby = b->height > 0 ? (b->height + cbh - 1) / cbh : 0;
            for (int cby = 0; cby < ncby; cby++)
                for (int cbx = 0; cbx < ncbx; cbx++) { /* partitionIntoCodeBlocks :3226-3283 */
                    int x0 = cbx * cbw, y0 = cby * cbh, x1 = x0 + cbw, y1 = y0 + cbh;
                    if (x1 > b->width) x1 = b->width;
                    if (y1 > b->height) y1 = b->height;
                    int aw = x1 - x0, ah = y1 - y0;
                    if (table && n < max_blocks)
                        table[n] = (j2k_cblk){b->x0 + x0, b->y0 + y0, aw, ah, cbx, cby, b->band, b->res, off};
                    if (blocks) {
                        int32_t* cb = blocks + off;
                        for (int y = 0; y < ah; y++)
                            for (int x = 0; x < aw; x++) cb[(size_t)y * aw + x] = b->data[(size_t)(y0 + y) * b->width + (x0 + x)];
                        if (shift6) /* encodeCodeBlock :3294-3300 */
                            for (int i = 0; i < aw * ah; i++) cb[i] = (int32_t)((uint32_t)cb[i] << 6);
                        if (numbps) { /* codeBlockNumBps :3349-3362 */
                            int raw = calculate_max_bitplane(cb, (size_t)aw * ah);
                            int v = raw < 0 ? 0 : raw + 1 - (htj2k ? 0 : 6);
                            numbps[n] = v < 0 ? 0 : v;
                        }
                    }
                    off += (int64_t)aw * ah;
                    n++;
                }
            free(b->data);
        }
    }
    return n;
}

ORC_API int orc_codeblock_layout(int width, int height, int num_levels, int cbw, int cbh, j2k_cblk* out, int max_blocks) {
    return walk_code_blocks(NULL, width, height, num_levels, cbw, cbh, 0, 0, out, max_blocks, NULL, NULL);
}

/* plane (Mallat layout, stride = width) -> block-major plane + numbps, one tile-component */
ORC_API int orc_gather_blocks(const int32_t* plane, int width, int height, int num_levels, int cbw, int cbh, int shift6, int htj2k,
                              int32_t* blocks, int32_t* numbps) {
    return walk_code_blocks(plane, width, height, num_levels, cbw, cbh, shift6, htj2k, NULL, 0, blocks, numbps);
}

/* TileDecoder.assembleSubbands (jpeg2000/t2/tile_decoder.go:840-883): block-major plane -> plane */
ORC_API int orc_scatter_blocks(const int32_t* blocks, int width, int height, int num_levels, int cbw, int cbh, int32_t* plane) {
    int n = orc_codeblock_layout(width, height, num_levels, cbw, cbh, NULL, 0);
    j2k_cblk* t = (j2k_cblk*)calloc((size_t)n + 1, sizeof(j2k_cblk));
    orc_codeblock_layout(width, height, num_levels, cbw, cbh, t, n);
    memset(plane, 0, (size_t)width * height * sizeof(int32_t));
    for (int i = 0; i < n; i++)
        for (int y = 0; y < t[i].height; y++)
            for (int x = 0; x < t[i].width; x++) {
                size_t dst = (size_t)(t[i].y0 + y) * width + (t[i].x0 + x);
                if (dst < (size_t)width * height) plane[dst] = blocks[t[i].offset + (int64_t)y * t[i].width + x];
            }
    free(t);
    return n;
}

/* applyInverseGeneralScaling / applyInverseGeneralScalingMasked (jpeg2000/t2/tile_decoder.go:1082-1111): data /= 2^shift with
 * Go's truncating division, for the whole block or for the samples its mask names (mask == NULL: whole block). */
ORC_API void orc_inverse_general_scaling(int32_t* data, const uint8_t* mask, size_t n, int shift) {
    if (shift <= 0) return;
    const int32_t factor = (int32_t)((int64_t)1 << shift);
    for (size_t i = 0; i < n; i++)
        if (!mask || mask[i]) data[i] /= factor;
}

/* applyInverseMaxShift (jpeg2000/t2/tile_decoder.go:1113-1138), the Srgn = 0 branch of decodeCodeBlock (:726-730):
 * shift <= 0 leaves the block alone, shift >= 31 zeroes it, otherwise magnitudes >= 2^shift come down by shift
 * (Go's -val wraps for INT_MIN: the magnitude stays negative, fails the threshold test and the value is kept). */
ORC_API void orc_inverse_max_shift(int32_t* data, size_t n, int shift) {
    if (shift <= 0) return;
    if (shift >= 31) { for (size_t i = 0; i < n; i++) data[i] = 0; return; }
    const int32_t thresh = (int32_t)1 << shift;
    for (size_t i = 0; i < n; i++) {
        const int32_t val = data[i];
        int32_t mag = val;
        if (mag < 0) mag = (int32_t)(0u - (uint32_t)mag);
        if (mag >= thresh) {
            mag >>= shift;
            data[i] = val < 0 ? -mag : mag;
        }
    }
}

ORC_API int orc_abi_sizes(int* fwd, int* inv, int* binding) {
    *fwd = (int)sizeof(j2k_fwd_params); *inv = (int)sizeof(j2k_inv_params); *binding = (int)sizeof(j2k_mct_binding);
    return J2K_B200_ABI_VERSION;
}
