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
    /* code-block interface with a decode-side MaxShift ROI: blocks out, every magnitude pushed up by 5 bits as an encoder
     * with an all-covering region would leave them (and halved as the classic T1 hands them back), shift undone on the way in */
    {
        static int32_t blk[W * H], nbps[64];
        int32_t shift[1] = {5};
        rc = j2k_forward_blocks(ctx, &fp, 16, 16, 1, px, sizeof px, blk, nbps);
        if (rc != 0 || j2k_fwd_block_count(&fp, 16, 16) != (size_t)nb) { fprintf(stderr, "forward_blocks: %d %s\n", rc, j2k_last_error(ctx)); return 9; }
        for (int i = 0; i < W * H; i++) {
            int32_t v = (blk[i] >> 6) * 2, m = v < 0 ? -v : v;  /* T1 output of the classic lossless path: one half bit */
            m <<= 5;
            blk[i] = v < 0 ? -m : m;
        }
        ip.fuse_t1_halve = 1;
        rc = j2k_inverse_blocks_roi(ctx, &ip, 16, 16, 1, blk, shift, back, sizeof back, NULL);
        ip.fuse_t1_halve = 0;
        if (rc != 0) { fprintf(stderr, "inverse_blocks_roi: %d %s\n", rc, j2k_last_error(ctx)); return 10; }
        /* magnitudes below 2^5 after the shift are only the zeros, so the MaxShift rule restores every sample */
        if (memcmp(px, back, sizeof px) != 0) { fprintf(stderr, "ROI block round trip differs\n"); return 11; }
    }
    rc = j2k_forward(ctx, &fp, px, 10, co, W * H);  /* short buffer: the reference's error, not a crash */
    if (rc != J2K_ERR_SIZE || strstr(j2k_last_error(ctx), "insufficient pixel data") == NULL) { fprintf(stderr, "error path\n"); return 8; }
    j2k_shutdown(ctx);
    printf("c abi ok: %d code-blocks\n", nb);
    return 0;
}
