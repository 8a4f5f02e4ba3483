/*
 * nblic_dropin.c -- the reference codec's five entry points on top of the batch C ABI (host side, plain C).
 *
 * Replaces src/NBLIC.c:915-926 and src/QNBLIC.c:493-655,872-883 behind the unchanged prototypes of
 * src/NBLIC.h:54,72 and src/QNBLIC.h:14-18: the reference CLI (src/NBLIC_main.c) links against this
 * file + libnblic_b200 without modification.  Every call is a batch of one image on a lazily
 * created per-process context (device NBLIC_B200_DEVICE, default 0); calls are serialised by a mutex,
 * so the functions stay thread-safe like the reference's.  There is no CPU path: without a GPU the
 * calls return -1.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "../../include/nblic_b200.h"

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
static nblic_b200_ctx *g_ctx = NULL;
static __thread size_t g_hint_len = 0;

static nblic_b200_ctx *shared_ctx(void) { /* call with g_lock held */
    if (!g_ctx) {
        const char *dev = getenv("NBLIC_B200_DEVICE");
        g_ctx = nblic_b200_create(dev ? atoi(dev) : 0);
        if (!g_ctx) fprintf(stderr, "nblic_b200: %s\n", nblic_b200_last_error(NULL));
    }
    return g_ctx;
}

void nblic_b200_hint_input_len(size_t n_bytes) { g_hint_len = n_bytes; }

static int dims_ok(int h, int w) { /* src/NBLIC.c:717-729, src/QNBLIC.c:33-45 */
    return h > 0 && w > 0 && h <= NBLIC_MAX_HEIGHT && w <= NBLIC_MAX_WIDTH && (long long)h * w <= NBLIC_MAX_IMG_SIZE;
}

static int clip(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

static int encode_one(unsigned char *p_buf, unsigned char *p_img, int height, int width, int near, int effort, int want_recon) {
    const uint8_t *img = p_img;
    uint8_t *out = p_buf, *rec = want_recon ? p_img : NULL;
    size_t cap = nblic_b200_stream_bound(height, width), len = 0;
    int status = 0, rc;
    pthread_mutex_lock(&g_lock);
    rc = shared_ctx() ? nblic_b200_encode_batch(g_ctx, 1, &img, &height, &width, near, effort, &out, &cap, &len, want_recon ? &rec : NULL, &status) : -1;
    pthread_mutex_unlock(&g_lock);
    return rc == 0 ? (int)len : -1;
}

int NBLICcompress(int verbose, unsigned char *p_buf, unsigned char *p_img, int height, int width, int *p_near, int *p_effort) {
    int len;
    /* parameters are clipped in place and the header is emitted before validation (src/NBLIC.c:768-775) */
    *p_near = clip(*p_near, 0, 9);
    *p_effort = clip(*p_effort, 1, 3);
    memcpy(p_buf, "NBLIC0.3", 8);
    p_buf[8] = 1;
    p_buf[9] = (unsigned char)(height >> 8); p_buf[10] = (unsigned char)height;
    p_buf[11] = (unsigned char)(width >> 8); p_buf[12] = (unsigned char)width;
    p_buf[13] = (unsigned char)*p_near; p_buf[14] = (unsigned char)clip(3 + 2 * *p_near, 3, 16); p_buf[15] = (unsigned char)*p_effort;
    if (!dims_ok(height, width)) return -1;
    /* near > 0: the reference codes in place, leaving the reconstruction in p_img (src/NBLIC.c:876,916) */
    len = encode_one(p_buf, p_img, height, width, *p_near, *p_effort, *p_near > 0);
    if (verbose && len > 0) printf("\r    %d rows (B200), compressed length=%d B\n", height, len);
    return len;
}

/*
 * The reference decoders get no input length (src/NBLIC.h:72, src/QNBLIC.h:14): they read forward until the last
 * pixel is out.  A GPU decoder has to upload the stream first, so the legacy wrappers need an extent.  With the
 * side-channel hint that is the caller's figure.  Without it the wrappers take the worst-case stream size for the
 * parsed dimensions, but never touch memory the caller does not own: the bytes are copied through a pipe
 * (write(2) from the caller's buffer, read(2) into a bounce buffer), and the kernel stops a write at the first
 * unreadable page instead of faulting.  The copy therefore ends at min(worst case, readable extent); whatever lies
 * between the true end of the stream and that point is uploaded and never looked at (a valid stream is consumed
 * exactly to its end).
 */
static uint8_t *readable_prefix(const uint8_t *src, size_t want, size_t *got) {
    int fd[2];
    uint8_t *bounce = (uint8_t *)malloc(want ? want : 1);
    size_t done = 0;
    *got = 0;
    if (!bounce) return NULL;
    if (pipe(fd) != 0) { free(bounce); return NULL; }
    while (done < want) {
        /* pieces stay below the default pipe capacity (write never blocks) and, after a head piece that ends at a page
         * boundary of the source, start page-aligned: the kernel copies pipe pages whole and drops a partly copied one,
         * so only an aligned piece is guaranteed to deliver every readable byte in front of a protected page */
        const size_t to_page = 4096 - (size_t)((uintptr_t)(src + done) & 4095);
        size_t piece = to_page < 4096 ? to_page : 32768;
        ssize_t wr, rd = 0;
        if (piece > want - done) piece = want - done;
        wr = write(fd[1], src + done, piece);
        if (wr <= 0) break; /* EFAULT: the first byte of the piece is already unreadable */
        while (rd < wr) {
            ssize_t r = read(fd[0], bounce + done + rd, (size_t)(wr - rd));
            if (r <= 0) break;
            rd += r;
        }
        done += (size_t)rd;
        if ((size_t)wr < piece || rd < wr) break; /* short write: the piece ran into an unreadable page */
    }
    close(fd[0]); close(fd[1]);
    *got = done;
    return bounce;
}

static int decode_one(const uint8_t *stream, size_t len, unsigned char *p_img, int h, int w) {
    uint8_t *img = p_img;
    size_t cap = (size_t)h * w;
    int status = 0, rc;
    pthread_mutex_lock(&g_lock);
    rc = shared_ctx() ? nblic_b200_decode_batch(g_ctx, 1, &stream, &len, &img, &cap, NULL, NULL, NULL, NULL, &status) : -1;
    pthread_mutex_unlock(&g_lock);
    return rc == 0 ? 0 : -1;
}

/* decode p_buf with the extent rules above */
static int decode_legacy(const uint8_t *p_buf, unsigned char *p_img, int h, int w) {
    size_t n = g_hint_len, got = 0;
    uint8_t *bounce;
    int rc;
    g_hint_len = 0;
    if (n) return decode_one(p_buf, n, p_img, h, w);
    n = nblic_b200_stream_bound(h, w);
    if (n > (size_t)2 * NBLIC_MAX_IMG_SIZE) n = (size_t)2 * NBLIC_MAX_IMG_SIZE; /* src/NBLIC_main.c:141 */
    bounce = readable_prefix(p_buf, n, &got);
    if (!bounce) return -1;
    rc = got ? decode_one(bounce, got, p_img, h, w) : -1;
    free(bounce);
    return rc;
}

int NBLICdecompress(int verbose, unsigned char *p_buf, unsigned char *p_img, int *p_height, int *p_width, int *p_near, int *p_effort) {
    int h = 0, w = 0, near = 0, effort = 0, rc;
    if (memcmp(p_buf, "NBLIC0.3", 8) != 0) return -1; /* src/NBLIC.c:700-702: nothing is written on a bad magic */
    rc = nblic_b200_peek(p_buf, 16, &h, &w, &near, &effort);
    *p_height = h; *p_width = w; *p_near = near; *p_effort = effort; /* src/NBLIC.c:703-711 */
    if (rc != 0) return -1;
    rc = decode_legacy(p_buf, p_img, h, w);
    if (verbose && rc == 0) printf("\r    %d rows (B200)\n", h);
    return rc;
}

int QNBLICcompress(uint16_t *p_buf, unsigned char *p_img, int height, int width) {
    int len;
    if (!dims_ok(height, width)) return -1; /* src/QNBLIC.c:575 */
    len = encode_one((unsigned char *)p_buf, p_img, height, width, 0, 0, 0);
    return len < 0 ? -1 : len / 2;
}

/* The reference's -t pipeline produces the same words as QNBLICcompress (src/QNBLIC.c:660-883). */
int QNBLICcompressMultiThread(uint16_t *p_buf, unsigned char *p_img, int height, int width) {
    return QNBLICcompress(p_buf, p_img, height, width);
}

int QNBLICdecompress(uint16_t *p_buf, unsigned char *p_img, int *p_height, int *p_width) {
    const uint8_t *bytes = (const uint8_t *)p_buf;
    int h = 0, w = 0, near = 0, effort = 0, rc;
    if (memcmp(bytes, "Q0.2", 4) != 0) return -1; /* format sniff: no CUDA work for foreign input (src/QNBLIC.c:475-486) */
    rc = nblic_b200_peek(bytes, 8, &h, &w, &near, &effort);
    *p_height = h; *p_width = w;
    if (rc != 0) return -1;
    return decode_legacy(bytes, p_img, h, w);
}
