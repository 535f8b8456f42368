/* A plain C caller of the C ABI: include/j2k_b200.h must compile as C, and a C program must be able to drive a whole
 * forward / inverse round trip through it.  Built by tests/test_abi_symbols.py against the CPU-emulator flavour of the
 * library (tests/emu) so that it runs without a GPU; on a GPU box the same program links against libj2kb200.so. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "j2k_b200.h"

int main(void) {
    j2k_ctx* ctx = NULL;
    int rc = j2k_init(&ctx, NULL, 0);
    if (rc != 0) { fprintf(stderr, "init: %d %s\n", rc, j2k_last_error(NULL)); return 2; }
    enum { W = 64, H = 48, L = 3 };
    j2k_fwd_params fp;
    j2k_inv_params ip;
    memset(&fp, 0, sizeof fp);
    memset(&ip, 0, sizeof ip);
    fp.width = W; fp.height = H; fp.components = 1; fp.bit_depth = 12; fp.num_levels = L; fp.reversible = 1; fp.mct_mode = J2K_MCT_NONE;
    ip.xsiz = W; ip.ysiz = H; ip.xtsiz = W; ip.ytsiz = H; ip.components = 1; ip.bit_depth = 12; ip.num_levels = L; ip.reversible = 1;
    ip.mct_mode = J2K_MCT_NONE;
    unsigned char px[W * H * 2], back[W * H * 2];
    int32_t co[W * H];
    for (int i = 0; i < W * H; i++) { unsigned v = (unsigned)(i * 2654435761u) >> 20; px[2 * i] = (unsigned char)(v & 0xFF); px[2 * i + 1] = (unsigned char)(v >> 8); }
    if (j2k_fwd_pixel_bytes(&fp) != sizeof px || j2k_fwd_coeff_count(&fp) != (size_t)(W * H)) { fprintf(stderr, "sizes\n"); return 3; }
    rc = j2k_forward(ctx, &fp, px, sizeof px, co, W * H);
    if (rc != 0) { fprintf(stderr, "forward: %d %s\n", rc, j2k_last_error(ctx)); return 4; }
    rc = j2k_inverse(ctx, &ip, co, W * H, back, sizeof back, NULL);
    if (rc != 0) { fprintf(stderr, "inverse: %d %s\n", rc, j2k_last_error(ctx)); return 5; }
    if (memcmp(px, back, sizeof px) != 0) { fprintf(stderr, "lossless round trip differs\n"); return 6; }
    j2k_cblk tab[64];
    int nb = j2k_codeblock_layout(W, H, L, 16, 16, tab, 64);
    if (nb <= 0 || tab[0].band != 0 || tab[0].offset != 0) { fprintf(stderr, "layout\n"); return 7; }
    rc = j2k_forward(ctx, &fp, px, 10, co, W * H);  /* short buffer: the reference's error, not a crash */
    if (rc != J2K_ERR_SIZE || strstr(j2k_last_error(ctx), "insufficient pixel data") == NULL) { fprintf(stderr, "error path\n"); return 8; }
    j2k_shutdown(ctx);
    printf("c abi ok: %d code-blocks\n", nb);
    return 0;
}
