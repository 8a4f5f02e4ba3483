/*
 * tests/stubs/cli_stub_backend.c -- TEST INFRASTRUCTURE, never shipped: the handful of nblic_b200_* entry points that
 * csrc/nblic_batch_cli.c and csrc/nblic_dropin.c call, implemented on the CPU by the parity oracle
 * (oracle/nblic_oracle.c), so that the HOST-side C above the batch ABI -- the CLI's switch parsing, PGM / BMP readers,
 * grouping, reader / coder / writer threads and error paths; the drop-in wrappers' header-before-validation, in-place
 * clipping, reconstruction copy-back and extent probing -- runs under `pytest -m "not gpu"` in a container without a GPU.
 * The product never links this file: nblic_batch and the drop-in symbols proper live on libnblic_b200.so, which has no CPU path.
 */
#include <stdlib.h>
#include <string.h>

#include "../../include/nblic_b200.h"
#include "../../oracle/nblic_oracle.h"

struct nblic_b200_ctx { int unused; };

nblic_b200_ctx *nblic_b200_create(int device) { (void)device; return (nblic_b200_ctx *)calloc(1, sizeof(nblic_b200_ctx)); }
void nblic_b200_destroy(nblic_b200_ctx *c) { free(c); }
const char *nblic_b200_last_error(const nblic_b200_ctx *c) { (void)c; return "stub backend"; }
void *nblic_b200_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void nblic_b200_host_free(void *p) { free(p); }
size_t nblic_b200_stream_bound(int h, int w) { return h > 0 && w > 0 ? 2 * (size_t)h * (size_t)w + 8192 : 8192; }

int nblic_b200_peek(const uint8_t *p, size_t len, int *h, int *w, int *near, int *effort) { /* same rules as csrc/kernels.cuh: peek_bytes */
    if (len >= 8 && p[0] == 0x51 && p[1] == 0x30 && p[2] == 0x2e && p[3] == 0x32) {
        *h = p[4] | (p[5] << 8); *w = p[6] | (p[7] << 8); *near = 0; *effort = 0;
        return *h > 0 && *w > 0 && (long long)*h * *w <= 100000000LL ? 0 : -1;
    }
    if (len < 16 || memcmp(p, "NBLIC0.3", 8) != 0) return -1;
    *h = (p[9] << 8) | p[10]; *w = (p[11] << 8) | p[12]; *near = p[13]; *effort = p[15];
    return *h > 0 && *w > 0 && (long long)*h * *w <= 100000000LL && p[8] <= 1 && p[13] <= 9 && p[14] >= 3 && p[14] <= 16 && p[15] >= 1 && p[15] <= 3 ? 0 : -1;
}

int nblic_b200_encode_batch(nblic_b200_ctx *c, int n, const uint8_t *const *images, const int *hs, const int *ws, int near, int effort,
                            uint8_t *const *outs, const size_t *caps, size_t *lens, uint8_t *const *recon, int *status) {
    int i, failed = 0;
    (void)c;
    for (i = 0; i < n; i++) {
        const size_t px = (size_t)hs[i] * ws[i];
        uint8_t *tmp = (uint8_t *)malloc(px), *out = (uint8_t *)malloc(2 * px + 65536);
        int len, n_ = near, e_ = effort;
        memcpy(tmp, images[i], px);
        if (near == 0 && effort == 0) len = 2 * oracle_q_encode(tmp, hs[i], ws[i], (uint16_t *)out);
        else len = oracle_n_encode(tmp, hs[i], ws[i], &n_, &e_, out);
        status[i] = len > 0 && (size_t)len <= caps[i] ? NBLIC_B200_OK : NBLIC_B200_OVERFLOW;
        if (status[i] == NBLIC_B200_OK) { memcpy(outs[i], out, (size_t)len); lens[i] = (size_t)len; } else { lens[i] = 0; failed++; }
        if (status[i] == NBLIC_B200_OK && recon && recon[i] && n_ > 0) memcpy(recon[i], tmp, px); /* the oracle codes in place, like the reference */
        free(tmp); free(out);
    }
    return failed;
}

int nblic_b200_decode_batch(nblic_b200_ctx *c, int n, const uint8_t *const *streams, const size_t *lens, uint8_t *const *images, const size_t *caps,
                            int *hs, int *ws, int *nears, int *efforts, int *status) {
    int i, failed = 0;
    (void)c;
    for (i = 0; i < n; i++) {
        int h = 0, w = 0, near = 0, effort = 0, rc;
        if (nblic_b200_peek(streams[i], lens[i], &h, &w, &near, &effort) != 0) { status[i] = NBLIC_B200_BAD_HEADER; failed++; continue; }
        if ((size_t)h * w > caps[i]) { status[i] = NBLIC_B200_OVERFLOW; failed++; continue; }
        if (effort == 0) rc = oracle_q_decode((const uint16_t *)streams[i], (long)(lens[i] / 2), images[i], &h, &w);
        else rc = oracle_n_decode(streams[i], (long)lens[i], images[i], &h, &w, &near, &effort);
        if (hs) hs[i] = h;
        if (ws) ws[i] = w;
        if (nears) nears[i] = near;
        if (efforts) efforts[i] = effort;
        status[i] = rc == 0 ? NBLIC_B200_OK : NBLIC_B200_CORRUPT;
        failed += rc != 0;
    }
    return failed;
}
