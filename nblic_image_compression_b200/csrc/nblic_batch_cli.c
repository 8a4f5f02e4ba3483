/*
 * nblic_batch_cli.c -- batch front end for libnblic_b200.so: the reference CLI
 * (src/NBLIC_main.c) handles one image per process launch; this one loads many gray images (PGM P5 or
 * 8-bit BMP), encodes or decodes them in ONE batch call so every image gets its own warp, and writes the
 * results.  Switches follow the reference's (-c / -d, -n<N>, -e<E>, -v; src/NBLIC_main.c:52-95).
 *
 *   nblic_batch -c [-n<near>] [-e<effort>] [-g<gpu>] [-v] <out_dir> <in.pgm|in.bmp>...   -> <out_dir>/<stem>.nblic
 *   nblic_batch -d [-g<gpu>] [-v] <out_dir> <in.nblic>...                              -> <out_dir>/<stem>.pgm
 *
 * The file formats are read and written by this file's own small parsers (binary PGM; 8-bit palettised
 * BMP, bottom-up or top-down, rows padded to 4 bytes, the palette index taken as the gray value as the
 * reference's loader does, src/FileIO.c:170-245).  Host side only; all coding happens on the GPU.
 */
#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/nblic_b200.h"

typedef struct { uint8_t *data; size_t len; } blob_t;

static blob_t read_file(const char *path) {
    blob_t b = {NULL, 0};
    FILE *f = fopen(path, "rb");
    long n;
    if (!f) return b;
    if (fseek(f, 0, SEEK_END) != 0 || (n = ftell(f)) < 0 || fseek(f, 0, SEEK_SET) != 0) { fclose(f); return b; }
    b.data = (uint8_t *)malloc((size_t)n + 16);
    if (b.data && fread(b.data, 1, (size_t)n, f) == (size_t)n) { b.len = (size_t)n; memset(b.data + n, 0, 16); }
    else { free(b.data); b.data = NULL; }
    fclose(f);
    return b;
}

static int write_file(const char *path, const void *head, size_t head_len, const void *body, size_t body_len) {
    FILE *f = fopen(path, "wb");
    int ok;
    if (!f) return -1;
    ok = (head_len == 0 || fwrite(head, 1, head_len, f) == head_len) && fwrite(body, 1, body_len, f) == body_len;
    return fclose(f) == 0 && ok ? 0 : -1;
}

static uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

/* binary PGM: "P5" <ws> width <ws> height <ws> maxval(<256) <single ws> raster; '#' comments allowed in the header */
static int parse_pgm(const blob_t *b, uint8_t **pix, int *h, int *w) {
    size_t p = 2;
    long v[3];
    int k;
    if (b->len < 7 || b->data[0] != 'P' || b->data[1] != '5') return -1;
    for (k = 0; k < 3; k++) {
        for (;;) {
            while (p < b->len && isspace(b->data[p])) p++;
            if (p < b->len && b->data[p] == '#') { while (p < b->len && b->data[p] != '\n') p++; continue; }
            break;
        }
        if (p >= b->len || !isdigit(b->data[p])) return -1;
        v[k] = 0;
        while (p < b->len && isdigit(b->data[p])) { v[k] = v[k] * 10 + (b->data[p] - '0'); if (v[k] > 100000000L) return -1; p++; }
    }
    if (p >= b->len || !isspace(b->data[p])) return -1;
    p++;
    if (v[0] <= 0 || v[1] <= 0 || v[2] <= 0 || v[2] > 255 || (size_t)v[0] * (size_t)v[1] > b->len - p) return -1;
    *w = (int)v[0]; *h = (int)v[1];
    *pix = (uint8_t *)malloc((size_t)v[0] * (size_t)v[1]);
    if (!*pix) return -1;
    memcpy(*pix, b->data + p, (size_t)v[0] * (size_t)v[1]);
    return 0;
}

static int parse_bmp(const blob_t *b, uint8_t **pix, int *h, int *w) {
    uint32_t off, bpp, compression;
    int32_t bw, bh;
    size_t stride;
    int y, rows;
    if (b->len < 54 || b->data[0] != 'B' || b->data[1] != 'M') return -1;
    off = le32(b->data + 10);
    bw = (int32_t)le32(b->data + 18); bh = (int32_t)le32(b->data + 22);
    bpp = b->data[28] | ((uint32_t)b->data[29] << 8);
    compression = le32(b->data + 30);
    rows = bh < 0 ? -bh : bh;
    if (bpp != 8 || compression != 0 || bw <= 0 || rows <= 0) return -1;
    stride = ((size_t)bw + 3) & ~(size_t)3;
    if (off > b->len || stride * (size_t)rows > b->len - off) return -1;
    *pix = (uint8_t *)malloc((size_t)bw * (size_t)rows);
    if (!*pix) return -1;
    for (y = 0; y < rows; y++) { /* positive height = bottom-up */
        const uint8_t *src = b->data + off + stride * (size_t)(bh > 0 ? rows - 1 - y : y);
        memcpy(*pix + (size_t)y * (size_t)bw, src, (size_t)bw);
    }
    *w = bw; *h = rows;
    return 0;
}

static void out_path(char *dst, size_t cap, const char *dir, const char *in, const char *suffix) {
    const char *base = strrchr(in, '/');
    const char *dot;
    size_t stem;
    base = base ? base + 1 : in;
    dot = strrchr(base, '.');
    stem = dot ? (size_t)(dot - base) : strlen(base);
    snprintf(dst, cap, "%s/%.*s%s", dir, (int)stem, base, suffix);
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

int main(int argc, char **argv) {
    int decode = -1, near = 0, effort = 1, gpu = 0, verbose = 0, a = 1, n, i, failed, rc = 0;
    const char *dir;
    nblic_b200_ctx *ctx;
    uint8_t **in_data, **out_data;
    size_t *in_len, *out_cap, *out_len;
    int *hs, *ws, *nears, *efforts, *status;
    double t0, t1, pixels = 0, bytes = 0;

    for (; a < argc && argv[a][0] == '-'; a++) { /* compact switches, any order, as the reference CLI */
        const char *s = argv[a] + 1;
        for (; *s; s++) {
            if (*s == 'c') decode = 0;
            else if (*s == 'd') decode = 1;
            else if (*s == 'v' || *s == 'V') verbose = 1;
            else if (*s == 't') { /* accepted and ignored: every image already gets its own warp */ }
            else if (*s == 'n' || *s == 'e' || *s == 'g') {
                int val = 0, digits = 0;
                const char which = *s;
                while (isdigit((unsigned char)s[1])) { val = val * 10 + (s[1] - '0'); s++; digits++; }
                if (!digits) { fprintf(stderr, "-%c needs a number\n", which); return -1; }
                if (which == 'n') near = val; else if (which == 'e') effort = val; else gpu = val;
            } else { fprintf(stderr, "unknown switch -%c\n", *s); return -1; }
        }
    }
    if (decode < 0 || argc - a < 2) {
        fprintf(stderr, "usage: %s -c [-n<near>] [-e<effort>] [-g<gpu>] [-v] <out_dir> <in.pgm|in.bmp>...\n"
                        "       %s -d [-g<gpu>] [-v] <out_dir> <in.nblic>...\n", argv[0], argv[0]);
        return -1;
    }
    dir = argv[a++];
    n = argc - a;
    in_data = (uint8_t **)calloc((size_t)n, sizeof *in_data); out_data = (uint8_t **)calloc((size_t)n, sizeof *out_data);
    in_len = (size_t *)calloc((size_t)n, sizeof *in_len); out_cap = (size_t *)calloc((size_t)n, sizeof *out_cap);
    out_len = (size_t *)calloc((size_t)n, sizeof *out_len);
    hs = (int *)calloc((size_t)n, sizeof(int)); ws = (int *)calloc((size_t)n, sizeof(int)); nears = (int *)calloc((size_t)n, sizeof(int));
    efforts = (int *)calloc((size_t)n, sizeof(int)); status = (int *)calloc((size_t)n, sizeof(int));
    if (!in_data || !out_data || !in_len || !out_cap || !out_len || !hs || !ws || !nears || !efforts || !status) return -1;

    for (i = 0; i < n; i++) { /* load everything first: one batch call needs all inputs resident */
        blob_t b = read_file(argv[a + i]);
        if (!b.data) { fprintf(stderr, "  ***Error : open %s failed\n", argv[a + i]); return -1; }
        if (!decode) {
            if (parse_pgm(&b, &in_data[i], &hs[i], &ws[i]) != 0 && parse_bmp(&b, &in_data[i], &hs[i], &ws[i]) != 0) {
                fprintf(stderr, "  ***Error : %s is neither a binary PGM nor an 8-bit BMP\n", argv[a + i]);
                return -1;
            }
            free(b.data);
            out_cap[i] = nblic_b200_stream_bound(hs[i], ws[i]);
            pixels += (double)hs[i] * ws[i];
        } else {
            int e_ = 0, n_ = 0;
            in_data[i] = b.data; in_len[i] = b.len;
            if (nblic_b200_peek(b.data, b.len, &hs[i], &ws[i], &n_, &e_) != 0) { fprintf(stderr, "  ***Error : %s is not a .nblic stream\n", argv[a + i]); return -1; }
            out_cap[i] = (size_t)hs[i] * (size_t)ws[i];
            pixels += (double)out_cap[i];
        }
        out_data[i] = (uint8_t *)malloc(out_cap[i] ? out_cap[i] : 1);
        if (!out_data[i]) return -1;
    }

    ctx = nblic_b200_create(gpu);
    if (!ctx) { fprintf(stderr, "  ***Error : %s\n", nblic_b200_last_error(NULL)); return -1; }
    t0 = now_s();
    if (!decode)
        failed = nblic_b200_encode_batch(ctx, n, (const uint8_t *const *)in_data, hs, ws, near, effort, out_data, out_cap, out_len, NULL, status);
    else
        failed = nblic_b200_decode_batch(ctx, n, (const uint8_t *const *)in_data, in_len, out_data, out_cap, hs, ws, nears, efforts, status);
    t1 = now_s();
    if (failed < 0) { fprintf(stderr, "  ***Error : %s\n", nblic_b200_last_error(ctx)); return -1; }

    for (i = 0; i < n; i++) {
        char path[4096], head[64];
        if (status[i] != NBLIC_B200_OK) { fprintf(stderr, "  ***Error : %s failed (status %d)\n", argv[a + i], status[i]); rc = -1; continue; }
        if (!decode) {
            out_path(path, sizeof path, dir, argv[a + i], ".nblic");
            if (write_file(path, NULL, 0, out_data[i], out_len[i]) != 0) { fprintf(stderr, "  ***Error : write %s failed\n", path); rc = -1; }
            bytes += (double)out_len[i];
        } else {
            const int hl = snprintf(head, sizeof head, "P5\n%d %d\n255\n", ws[i], hs[i]);
            out_path(path, sizeof path, dir, argv[a + i], ".pgm");
            if (write_file(path, head, (size_t)hl, out_data[i], (size_t)hs[i] * (size_t)ws[i]) != 0) { fprintf(stderr, "  ***Error : write %s failed\n", path); rc = -1; }
            bytes += (double)in_len[i];
        }
        if (verbose) printf("  %s -> %s  %d x %d\n", argv[a + i], path, ws[i], hs[i]);
    }
    if (verbose)
        printf("  %d images, %.3f MPixel, %.0f stream bytes (%.4f bpp), batch call %.3f s (%.1f MPixel/s), %d failed\n", n, pixels / 1e6, bytes,
               pixels > 0 ? 8.0 * bytes / pixels : 0.0, t1 - t0, pixels / 1e6 / (t1 - t0 > 0 ? t1 - t0 : 1), failed);
    nblic_b200_destroy(ctx);
    return rc;
}
