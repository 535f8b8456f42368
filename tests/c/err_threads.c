/* The error text of a failing call is kept per CONTEXT, not per thread: a cgo caller's goroutine may be moved to another
 * OS thread between the failing call and j2k_last_error().  Thread A provokes an error and exits; thread B - which has
 * never made a failing call - reads the message through the context.  Also: two contexts do not see each other's
 * messages, and j2k_last_error_copy() returns the same text without handing out a pointer.
 * Built by tests/test_abi_symbols.py against the CPU-emulator flavour of the library. */
#include <pthread.h>
#include <stdio.h>
#include <string.h>

#include "j2k_b200.h"

static j2k_ctx* g_ctx;
static int g_rc;

static void* thread_a(void* arg) {
    (void)arg;
    j2k_fwd_params fp;
    unsigned char px[16] = {0};
    static int32_t co[64 * 48];
    memset(&fp, 0, sizeof fp);
    fp.width = 64; fp.height = 48; fp.components = 1; fp.bit_depth = 12; fp.num_levels = 3; fp.reversible = 1;
    g_rc = j2k_forward(g_ctx, &fp, px, sizeof px, co, 64 * 48);  /* short pixel buffer */
    return NULL;
}

static char g_seen[256];
static void* thread_b(void* arg) {
    (void)arg;
    strncpy(g_seen, j2k_last_error(g_ctx), sizeof g_seen - 1);
    return NULL;
}

int main(void) {
    j2k_ctx* other = NULL;
    pthread_t t;
    char buf[256];
    if (j2k_init(&g_ctx, NULL, 0) != 0 || j2k_init(&other, NULL, 0) != 0) { fprintf(stderr, "init: %s\n", j2k_last_error(NULL)); return 2; }
    pthread_create(&t, NULL, thread_a, NULL);
    pthread_join(t, NULL);
    if (g_rc != J2K_ERR_SIZE) { fprintf(stderr, "thread A: rc %d\n", g_rc); return 3; }
    pthread_create(&t, NULL, thread_b, NULL);
    pthread_join(t, NULL);
    if (strstr(g_seen, "insufficient pixel data") == NULL) { fprintf(stderr, "thread B saw '%s'\n", g_seen); return 4; }
    if (j2k_last_error(other)[0] != 0) { fprintf(stderr, "the other context saw '%s'\n", j2k_last_error(other)); return 5; }
    if (j2k_last_error_copy(g_ctx, buf, sizeof buf) != strlen(g_seen) || strcmp(buf, g_seen) != 0) { fprintf(stderr, "copy form differs\n"); return 6; }
    if (j2k_last_error_copy(g_ctx, buf, 8) != strlen(g_seen) || strlen(buf) != 7) { fprintf(stderr, "truncating copy\n"); return 7; }
    j2k_shutdown(other);
    j2k_shutdown(g_ctx);
    printf("error text crosses threads\n");
    return 0;
}
