/*
 * oracle/nblic_oracle.c -- CPU restatement of NBLIC / QNBLIC (TEST INFRASTRUCTURE, not product code).
 *
 * Plain C99, integer arithmetic with explicit two's-complement wrap (build with -fwrapv
 * -ffp-contract=off, see oracle/Makefile).  Written from the behaviour of the reference files
 * (citations "R:" are /root/reference/src/<file>:<lines>); the structure is this repo's own:
 * one table-driven predictor shared by both codecs, explicit stream objects for the two entropy
 * coders, and a single raster walker per codec.  Parity is pinned against the reference's own
 * output (tests/golden/, tests/test_oracle.py).
 */
#include "nblic_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef int64_t i64;
typedef uint64_t u64;
typedef uint32_t u32;

/* ------------------------------------------------------------------------------------------ */
/* shared helpers                                                                             */
/* ------------------------------------------------------------------------------------------ */

/* causal neighbourhood slots (R: NBLIC.c:287-304 / QNBLIC.c:48-64)
 *      s h f g r      row i-2
 *      q c b d t      row i-1
 *      e a X          row i      */
enum { NA, NB, NC, ND, NE, NF, NG, NH, NQ, NR, NS, NT, N_SLOTS };

static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline i64 clampl(i64 v, i64 lo, i64 hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline i64 wmul(i64 a, i64 b) { return (i64)((u64)a * (u64)b); }
static inline i64 wshl(i64 a, int s) { return (i64)((u64)a << s); }

static int dims_ok(int h, int w) { /* R: NBLIC.c:717-729, QNBLIC.c:33-45 */
    return h > 0 && w > 0 && h <= 65535 && w <= 65535 && (long)h * w <= 100000000L;
}

/* Positional sampling with cascading fallbacks. R: NBLIC.c:287-304 (t) and QNBLIC.c:48-64 (no t). */
static void sample_positional(const uint8_t *img, int w, int i, int j, int nb[N_SLOTS]) {
#define PIX(di, dj, dflt) (((i + (di)) >= 0 && (j + (dj)) >= 0 && (j + (dj)) < w) ? (int)img[(long)(i + (di)) * w + (j + (dj))] : (dflt))
    int a = PIX(0, -1, 128);
    int b = PIX(-1, 0, 128);
    if (i == 0) b = a; else if (j == 0) a = b;
    nb[NA] = a; nb[NB] = b;
    nb[NE] = PIX(0, -2, a);
    nb[NC] = PIX(-1, -1, b);
    nb[ND] = PIX(-1, 1, b);
    nb[NF] = PIX(-2, 0, b);
    nb[NG] = PIX(-2, 1, nb[NF]);
    nb[NH] = PIX(-2, -1, nb[NF]);
    nb[NQ] = PIX(-1, -2, nb[NC]);
    nb[NR] = PIX(-2, 2, nb[NG]);
    nb[NS] = PIX(-2, -2, nb[NH]);
    nb[NT] = PIX(-1, 2, nb[ND]);
#undef PIX
}

/* The 7-direction gradient predictor, table driven.  R: NBLIC.c:307-370, QNBLIC.c:94-149.
 * Every directional cost is sum_t |2*base[t] - nb[u[t]] - nb[v[t]]| over base = (a,c,b,d); the
 * four axis-aligned directions use u == v (the reference writes those as 2*|base - nb[u]|). */
static const unsigned char PROBE[4][4] = {
    {NE, NQ, NC, NB}, /* one step "west"       */
    {NQ, NS, NH, NF}, /* one step "north-west" */
    {NC, NH, NF, NG}, /* one step "north"      */
    {NB, NF, NG, NR}, /* one step "north-east" */
};
static const unsigned char DIR_PROBES[7][2] = {{0, 0}, {2, 2}, {1, 1}, {3, 3}, {0, 1}, {1, 2}, {2, 3}};
static const unsigned char DIR_SOURCE[7][2] = {{NA, NA}, {NB, NB}, {NC, NC}, {ND, ND}, {NA, NC}, {NC, NB}, {NB, ND}};

typedef struct { int ang2, lin16, spread; } pred_terms_t;

static pred_terms_t predictor_terms(const int nb[N_SLOTS]) {
    const int base[4] = {nb[NA], nb[NC], nb[NB], nb[ND]};
    pred_terms_t r;
    int best = 0x7fffffff, total = 0, k, t;
    r.ang2 = 0;
    for (k = 0; k < 7; k++) {
        const unsigned char *u = PROBE[DIR_PROBES[k][0]], *v = PROBE[DIR_PROBES[k][1]];
        int cost = 0;
        for (t = 0; t < 4; t++) cost += iabs(2 * base[t] - nb[u[t]] - nb[v[t]]);
        total += cost;
        if (cost < best) { /* first minimum wins (strict <) */
            best = cost;
            r.ang2 = nb[DIR_SOURCE[k][0]] + nb[DIR_SOURCE[k][1]];
        }
    }
    r.lin16 = clampi(9 * nb[NA] + 9 * nb[NB] + 2 * nb[ND] - 2 * nb[NC] - nb[NE] - nb[NF], 0, 16 * 255);
    r.spread = total - 7 * best;
    return r;
}

static inline int blend_prediction(pred_terms_t p, int wt) {
    return (8 * wt * p.ang2 + (8 - wt) * p.lin16 + 64) >> 7;
}

static inline int activity(const int nb[N_SLOTS], int err) { /* R: NBLIC.c:376, QNBLIC.c:531 */
    return iabs(nb[NA] - nb[NE]) + iabs(nb[NB] - nb[NC]) + iabs(nb[NB] - nb[ND]) + iabs(nb[NA] - nb[NC]) +
           iabs(nb[NB] - nb[NF]) + iabs(nb[ND] - nb[NG]) + 2 * iabs(err);
}

static inline int texture_bits(const int nb[N_SLOTS], int px) { /* R: NBLIC.c:401-408; QNBLIC.c:164-173 has them MSB-first */
    int t = 0;
    t |= (px > nb[NA]) << 0;
    t |= (px > nb[NB]) << 1;
    t |= (px > nb[NC]) << 2;
    t |= (px > nb[ND]) << 3;
    t |= (px > nb[NE]) << 4;
    t |= (px > nb[NF]) << 5;
    t |= (px > 2 * nb[NA] - nb[NE]) << 6;
    t |= (px > 2 * nb[NB] - nb[NF]) << 7;
    return t;
}

/* ------------------------------------------------------------------------------------------ */
/* QNBLIC  ("Q0.2")                                                                           */
/* ------------------------------------------------------------------------------------------ */

#define Q_CLASSES 12
#define Q_CTX_SHIFT 11
#define Q_NORM_BITS 15
#define Q_NORM_SUM (1u << Q_NORM_BITS)

static int q_weight(int spread) { /* R: QNBLIC.c:82-91,144-146 */
    static const int bound[7] = {5, 12, 34, 78, 194, 431, 601};
    int s = spread >> 3, wt = 0, k;
    if (s > 607) s = 607;
    for (k = 0; k < 7; k++) wt += (s >= bound[k]);
    return wt;
}

static int q_class(int delta) { /* R: QNBLIC.c:152-161,532-533 */
    static const int bound[11] = {1, 2, 4, 6, 9, 15, 25, 39, 63, 101, 151};
    int c = 0, k;
    for (k = 0; k < 11; k++) c += (delta >= bound[k]);
    return c;
}

/* QNBLIC's texture bits are the same 8 comparisons but packed MSB-first (a is bit 7).  R: QNBLIC.c:164-173 */
static int q_ctx_address(const int nb[N_SLOTS], int px, int cls) {
    int t = texture_bits(nb, px), rev = 0, k;
    for (k = 0; k < 8; k++) rev |= ((t >> k) & 1) << (7 - k);
    return (cls << 8) | rev;
}

/* QNBLIC's sliding neighbourhood: positional only at j == 0, then a literal shift register whose
 * two fresh taps are (i-1, j+2) and (i-2, j+3).  R: QNBLIC.c:67-79.  `j` is the pixel just coded. */
static void q_window_shift(const uint8_t *img, int w, int i, int j, int x, int nb[N_SLOTS]) {
    int old_d = nb[ND], old_r = nb[NR];
    nb[NE] = nb[NA]; nb[NA] = x;
    nb[NQ] = nb[NC]; nb[NC] = nb[NB]; nb[NB] = old_d;
    nb[NS] = nb[NH]; nb[NH] = nb[NF]; nb[NF] = nb[NG]; nb[NG] = old_r;
    if (i <= 0) nb[ND] = nb[NA];
    else if (j + 2 < w) nb[ND] = img[(long)(i - 1) * w + j + 2];
    if (i <= 1) nb[NR] = nb[ND];
    else if (j + 3 < w) nb[NR] = img[(long)(i - 2) * w + j + 3];
}

static inline void q_bias_apply(int ctx, int px0, int *px, int *sign) { /* R: QNBLIC.c:176-180 */
    *sign = (ctx >> (Q_CTX_SHIFT - 1)) & 1;
    *px = clampi(px0 + (ctx >> Q_CTX_SHIFT) + *sign, 0, 255);
}
static inline int q_bias_learn(int ctx, int err) { /* R: QNBLIC.c:183-188 */
    return (ctx * 127 + err * (1 << Q_CTX_SHIFT) + 63) >> 7;
}

static int q_fold(int x, int px, int sign) { /* R: QNBLIC.c:191-202 */
    int room = px < 255 - px ? px : 255 - px, mag = iabs(x - px);
    if (mag == 0) return 0;
    if (mag <= room) return 2 * mag - ((x >= px) ^ sign);
    return mag + room;
}
static int q_unfold(int y, int px, int sign) { /* R: QNBLIC.c:205-217 */
    int room = px < 255 - px ? px : 255 - px;
    if (y <= 0) return px;
    if (y <= 2 * room) { int mag = (y + 1) >> 1; return ((y & 1) ^ sign) ? px + mag : px - mag; }
    return px < 128 ? px + (y - room) : px - (y - room);
}

/* Normalise a 256-bin histogram to sum 2^15.  R: QNBLIC.c:308-358 (double arithmetic, no FMA). */
static void q_normalise(u32 hist[256]) {
    u32 total = 0, live = 0, last = 0, k;
    for (k = 0; k < 256; k++) if (hist[k]) { total += hist[k]; live++; last = k; }
    if (live == 0) { hist[0] = Q_NORM_SUM - 1; hist[1] = 1; return; }
    if (live == 1) { hist[last] = Q_NORM_SUM - 1; hist[(last + 1) & 255] = 1; return; }
    {
        volatile double scale = (1.0 * Q_NORM_SUM) / total;
        u32 sum = 0;
        for (k = 0; k < 256; k++) if (hist[k]) {
            volatile double prod = scale * hist[k];
            u32 v = (u32)(0.49 + prod);
            hist[k] = v ? v : 1;
            sum += hist[k];
        }
        for (k = 0; sum > Q_NORM_SUM; k = (k + 1) & 255) if (hist[k] > 1) { hist[k]--; sum--; }
        for (k = 0; sum < Q_NORM_SUM; k = (k + 1) & 255) if (hist[k] > 0) { hist[k]++; sum++; }
    }
}

/* Histogram side information: 16-bit codes in five shapes.  R: QNBLIC.c:362-459. */
static uint16_t *q_put_hist(uint16_t *out, const u32 hist[256]) {
    u32 pos = 0, sum = 0;
    while (pos < 256 && sum < Q_NORM_SUM) {
        u32 head = hist[pos] & 0xffff, stop = pos + 1, follower = 0xffff, next, code;
        while (stop < 256) { follower = hist[stop] & 0xffff; if (follower != head) break; stop++; }
        if (head <= 1 && stop - pos >= 4) {            /* run of 0s or 1s, optionally closed by a 4-bit value */
            u32 run = stop - pos;
            next = stop;
            if (stop < 256 && follower <= 15) next = stop + 1; else follower = head;
            code = 0xE000u | (head << 12) | (follower << 8) | (run - 4);
        } else {
            u32 h1 = pos + 1 < 256 ? (hist[pos + 1] & 0xffff) : 0xffff;
            u32 h2 = pos + 2 < 256 ? (hist[pos + 2] & 0xffff) : 0xffff;
            u32 h3 = pos + 3 < 256 ? (hist[pos + 3] & 0xffff) : 0xffff;
            if (head <= 7 && h1 <= 7 && h2 <= 7 && h3 <= 7) { code = 0xD000u | (head << 9) | (h1 << 6) | (h2 << 3) | h3; next = pos + 4; }
            else if (head <= 15 && h1 <= 15 && h2 <= 15)   { code = 0xC000u | (head << 8) | (h1 << 4) | h2; next = pos + 3; }
            else if (head <= 127 && h1 <= 127)             { code = 0x8000u | (head << 7) | h1; next = pos + 2; }
            else                                           { code = head; next = pos + 1; }
        }
        *out++ = (uint16_t)code;
        for (; pos < next; pos++) sum += hist[pos];
    }
    return out;
}

typedef struct { const uint16_t *p; long left; } q_reader_t;
static inline u32 q_next_word(q_reader_t *r) { if (r->left > 0) { r->left--; return *r->p++; } return 0; }

static void q_get_hist(q_reader_t *rd, u32 hist[256]) {
    u32 pos = 0, sum = 0;
    memset(hist, 0, 256 * sizeof(u32));
#define PUSH(v) do { if (pos < 256) { hist[pos] = (v); sum += (v); } pos++; } while (0)
    while (pos < 256 && sum < Q_NORM_SUM) {
        u32 code = q_next_word(rd);
        if ((code >> 15) == 0) { PUSH(code); }
        else if ((code >> 14) == 2) { PUSH((code >> 7) & 0x7f); PUSH(code & 0x7f); }
        else if ((code >> 12) == 12) { PUSH((code >> 8) & 15); PUSH((code >> 4) & 15); PUSH(code & 15); }
        else if ((code >> 12) == 13) { PUSH((code >> 9) & 7); PUSH((code >> 6) & 7); PUSH((code >> 3) & 7); PUSH(code & 7); }
        else {
            u32 run = (code & 0xff) + 4, closer = (code >> 8) & 15, bit = (code >> 12) & 1;
            while (run--) PUSH(bit);
            if (closer != bit) PUSH(closer);
        }
    }
#undef PUSH
}

static void q_cumulate(const u32 hist[256], u32 acc[256]) { /* R: QNBLIC.c:290-295 */
    int k; acc[0] = 0; for (k = 1; k < 256; k++) acc[k] = acc[k - 1] + hist[k - 1];
}

int oracle_q_encode(const uint8_t *img, int height, int width, uint16_t *out) {
    static const uint16_t MAGIC[2] = {0x3051 /* "Q0" */, 0x322e /* ".2" */};
    u32 (*hist)[256], (*acc)[256];
    int *ctx, i, j, c;
    uint8_t *sym;   /* (class, y) pairs, raster order */
    uint16_t *o = out, *payload;
    long n, idx = 0;
    u32 state;

    if (!dims_ok(height, width)) return -1;
    n = (long)height * width;
    sym = malloc(2 * n);
    hist = calloc(Q_CLASSES, sizeof *hist);
    acc = calloc(Q_CLASSES, sizeof *acc);
    ctx = calloc(Q_CLASSES * 256, sizeof(int));
    if (!sym || !hist || !acc || !ctx) { free(sym); free(hist); free(acc); free(ctx); return -1; }

    /* pass 1: model every pixel, collect per-class statistics.  R: QNBLIC.c:586-623 */
    for (i = 0; i < height; i++) {
        int nb[N_SLOTS], err = 0;
        sample_positional(img, width, i, 0, nb);
        for (j = 0; j < width; j++) {
            int x = img[(long)i * width + j], px, sign, y, cls, adr;
            pred_terms_t pt = predictor_terms(nb);
            int px0 = blend_prediction(pt, q_weight(pt.spread));
            cls = q_class(activity(nb, err));
            err = x - px0;
            adr = q_ctx_address(nb, px0, cls);
            q_bias_apply(ctx[adr], px0, &px, &sign);
            y = q_fold(x, px, sign);
            sym[idx++] = (uint8_t)cls; sym[idx++] = (uint8_t)y;
            hist[cls][y]++;
            ctx[adr] = q_bias_learn(ctx[adr], err);
            q_window_shift(img, width, i, j, x, nb);
        }
    }

    /* header, side information.  R: QNBLIC.c:463-473,625-631 */
    *o++ = MAGIC[0]; *o++ = MAGIC[1]; *o++ = (uint16_t)height; *o++ = (uint16_t)width;
    for (c = 0; c < Q_CLASSES; c++) { q_normalise(hist[c]); q_cumulate(hist[c], acc[c]); o = q_put_hist(o, hist[c]); }

    /* pass 2: rANS over the symbols, last pixel first; words come out backwards.  R: QNBLIC.c:238-253,635-650 */
    payload = o;
    state = 1u << 16;
    while (idx > 0) {
        u32 f, base, quot;
        idx -= 2;
        f = hist[sym[idx]][sym[idx + 1]]; base = acc[sym[idx]][sym[idx + 1]];
        quot = state / f;
        if (quot > 0x1ffffu) { *o++ = (uint16_t)state; state >>= 16; quot = state / f; }
        state = (state % f) + (quot << Q_NORM_BITS) + base;
    }
    *o++ = (uint16_t)state; *o++ = (uint16_t)(state >> 16);
    { uint16_t *l = payload, *r = o - 1; while (l < r) { uint16_t t = *l; *l++ = *r; *r-- = t; } }

    free(sym); free(hist); free(acc); free(ctx);
    return (int)(o - out);
}

int oracle_q_decode(const uint16_t *in, long avail, uint8_t *img, int *height, int *width) {
    q_reader_t rd = {in, avail};
    u32 (*hist)[256], (*acc)[256], state;
    uint8_t (*lut)[Q_NORM_SUM];
    int *ctx, i, j, c, h, w;

    if (q_next_word(&rd) != 0x3051 || q_next_word(&rd) != 0x322e) return -1; /* R: QNBLIC.c:475-486 */
    h = (int)q_next_word(&rd); w = (int)q_next_word(&rd);
    *height = h; *width = w;
    if (!dims_ok(h, w)) return -1;

    hist = calloc(Q_CLASSES, sizeof *hist);
    acc = calloc(Q_CLASSES, sizeof *acc);
    lut = malloc(Q_CLASSES * sizeof *lut);
    ctx = calloc(Q_CLASSES * 256, sizeof(int));
    if (!hist || !acc || !lut || !ctx) { free(hist); free(acc); free(lut); free(ctx); return -1; }

    for (c = 0; c < Q_CLASSES; c++) { /* R: QNBLIC.c:298-305,512-516 */
        u32 v, k;
        q_get_hist(&rd, hist[c]);
        q_cumulate(hist[c], acc[c]);
        memset(lut[c], 255, Q_NORM_SUM);
        for (v = 0; v < 255; v++) for (k = acc[c][v]; k < acc[c][v + 1] && k < Q_NORM_SUM; k++) lut[c][k] = (uint8_t)v;
    }

    state = q_next_word(&rd) << 16; state |= q_next_word(&rd); /* R: QNBLIC.c:256-260 */

    for (i = 0; i < h; i++) { /* R: QNBLIC.c:520-552 */
        int nb[N_SLOTS], err = 0;
        sample_positional(img, w, i, 0, nb);
        for (j = 0; j < w; j++) {
            int px, sign, cls, adr, x;
            u32 slot, y;
            pred_terms_t pt = predictor_terms(nb);
            int px0 = blend_prediction(pt, q_weight(pt.spread));
            cls = q_class(activity(nb, err));
            adr = q_ctx_address(nb, px0, cls);
            q_bias_apply(ctx[adr], px0, &px, &sign);
            slot = state & (Q_NORM_SUM - 1);
            y = lut[cls][slot];
            state = (state >> Q_NORM_BITS) * hist[cls][y] + slot - acc[cls][y];
            if (state < (1u << 16)) state = (state << 16) | q_next_word(&rd);
            x = q_unfold((int)y, px, sign);
            img[(long)i * w + j] = (uint8_t)x;
            err = x - px0;
            ctx[adr] = q_bias_learn(ctx[adr], err);
            q_window_shift(img, w, i, j, x, nb);
        }
    }
    free(hist); free(acc); free(lut); free(ctx);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* NBLIC  ("NBLIC0.3")                                                                        */
/* ------------------------------------------------------------------------------------------ */

#define N_CLASSES 16
#define N_CTX_SHIFT 8
#define N_MIX 32            /* soft-class weight denominator */
#define N_RANKS 20          /* symbols tracked by the adaptive rank mapper */
#define N_PROB_ONE 4096
#define N_FRAC 12           /* fixed point of AVP predictions */
#define N_BIAS_MAX 4096
#define N_AVP_MAX_N 10
#define N_AVP_MAX_M (1 + N_AVP_MAX_N + N_AVP_MAX_N * N_AVP_MAX_N)

static const char N_MAGIC[8] = {'N', 'B', 'L', 'I', 'C', '0', '.', '3'};

static int n_weight(int spread) { /* R: NBLIC.c:308,365-367 */
    static const int bound[8] = {31, 93, 279, 620, 1550, 3410, 9300, 24800};
    int wt = 0, k;
    for (k = 0; k < 8; k++) wt += (spread >= bound[k]);
    return wt;
}

/* Soft 16-class activity quantiser: main class u, side class v, side weight wv in 0..16 (of 32).
 * R: NBLIC.c:373-395 */
static void n_soft_class(int delta, int *u, int *v, int *wv) {
    static const int mid[N_CLASSES] = {0, 2, 4, 7, 10, 14, 20, 26, 34, 42, 52, 64, 78, 95, 135, 200};
    int c = 0;
    while (c < N_CLASSES - 1 && delta > mid[c]) c++;
    *u = *v = c; *wv = 0;
    if (delta < mid[c]) {
        int w = N_MIX * (delta - mid[c - 1]) / (mid[c] - mid[c - 1]);
        if (w < N_MIX / 2) { *u = c - 1; *wv = w; } else { *v = c - 1; *wv = N_MIX - w; }
    }
}

static inline void n_bias_apply(int ctx, int px0, int *px, int *sign) { /* R: NBLIC.c:413-418 */
    *sign = (ctx >> (N_CTX_SHIFT - 1)) & 1;
    *px = clampi(px0 + (ctx >> N_CTX_SHIFT) + *sign, 0, 255);
}
static inline int n_bias_learn(int ctx, int err) { /* R: NBLIC.c:421-428 */
    return (ctx * 127 + err * (1 << N_CTX_SHIFT) + 64) >> 7;
}

/* near-aware residual fold / unfold.  R: NBLIC.c:431-466 */
static int n_fold(int x, int px, int sign, int near) {
    int q = 2 * near + 1;
    int room = (clampi(px, 0, 255 - px) + near) / q;
    int mag = (iabs(x - px) + near) / q;
    if (mag <= 0) return 0;
    if (mag <= room) return 2 * mag - ((x >= px) ^ sign);
    return mag + room;
}
static int n_unfold(int y, int px, int sign, int near) {
    int q = 2 * near + 1;
    int room = (clampi(px, 0, 255 - px) + near) / q;
    int mag, up;
    if (y <= 0) { mag = 0; up = 0; }
    else if (y <= 2 * room) { mag = (y + 1) / 2; up = (y & 1) ^ sign; }
    else { mag = y - room; up = px < 128; }
    mag *= q;
    return clampi(up ? px + mag : px - mag, 0, 255);
}

/* adaptive rank mapper: a permutation of 0..19 kept roughly sorted by frequency.  R: NBLIC.c:470-523 */
typedef struct { uint8_t rank_of[N_RANKS], sym_at[N_RANKS]; int count[N_RANKS]; } ranker_t;

static void ranker_reset(ranker_t *m) {
    int k; for (k = 0; k < N_RANKS; k++) { m->rank_of[k] = m->sym_at[k] = (uint8_t)k; m->count[k] = 2 * (N_RANKS - 1 - k); }
}
static void ranker_touch(ranker_t *m, int y) {
    int z;
    if (y >= N_RANKS) return;
    z = m->rank_of[y];
    m->count[z]++;
    if (z > 0 && m->count[z - 1] < m->count[z]) { /* one adjacent promotion */
        int other = m->sym_at[z - 1], t = m->count[z];
        m->count[z] = m->count[z - 1]; m->count[z - 1] = t;
        m->sym_at[z] = (uint8_t)other; m->sym_at[z - 1] = (uint8_t)y;
        m->rank_of[y] = (uint8_t)(z - 1); m->rank_of[other] = (uint8_t)z;
    }
}

/* carry-less 32-bit binary range coder.  R: NBLIC.c:527-586 */
typedef struct {
    uint8_t *wr; const uint8_t *rd; long rd_left;
    u32 lo, hi, code; int decoding;
} rc_t;

static inline u32 rc_read(rc_t *c) { if (c->rd_left > 0) { c->rd_left--; return *c->rd++; } return 0; }

static int rc_bit(rc_t *c, int bit, u32 p1) {
    u32 span = c->hi - c->lo;
    u32 mid = c->lo + (span >> 12) * p1 + (((span & 0xfff) * p1) >> 12);
    if (c->decoding) bit = c->code <= mid;
    if (bit) c->hi = mid; else c->lo = mid + 1;
    while (((c->lo ^ c->hi) & 0xff000000u) == 0) {
        if (c->decoding) c->code = (c->code << 8) | rc_read(c);
        else *c->wr++ = (uint8_t)(c->hi >> 24);
        c->lo <<= 8; c->hi = (c->hi << 8) | 0xff;
    }
    return bit;
}

/* pair of weighted hit counters per tree node.  R: NBLIC.c:589-637 */
typedef struct { int n0, n1; } node_t;

static inline int node_p1(const node_t *n) { return (N_PROB_ONE * n->n1) / (n->n0 + n->n1); }
static inline void node_learn(node_t *n, int bit, int weight) {
    if (bit) n->n1 += weight; else n->n0 += weight;
    if (n->n0 + n->n1 > N_MIX * 256) { n->n0 = (n->n0 + 1) >> 1; n->n1 = (n->n1 + 1) >> 1; }
}
static int mixed_bit(rc_t *c, node_t *u, node_t *v, int wv, int bit) {
    int p = (node_p1(u) * (N_MIX - wv) + node_p1(v) * wv + N_MIX / 2) / N_MIX;
    bit = rc_bit(c, bit, (u32)clampi(p, 1, N_PROB_ONE - 1));
    node_learn(u, bit, N_MIX - wv);
    node_learn(v, bit, wv);
    return bit;
}

/* adaptive-Golomb binarisation over the 16x256 node forest.  R: NBLIC.c:640-679 */
static int golomb_symbol(rc_t *c, int k_step, node_t (*forest)[256], int u, int v, int wv, int z) {
    const int top = (N_CLASSES - 1) / k_step;   /* largest Golomb order */
    int node = 0, k, bit = 0;
    if (v / k_step != u / k_step) v = u;
    for (;;) {
        k = u / k_step;
        if (!c->decoding) bit = (node >> top) < (z >> k);
        bit = mixed_bit(c, &forest[u][node], &forest[v][node], wv, bit);
        if (!bit) break;
        node += 1 << top;
        if (node >= 256) { node >>= 1; u = v = (k + 1) * k_step; } /* escape to the next order */
    }
    if (c->decoding) z = (node >> top) << k;
    for (node++, k--; k >= 0; k--) {
        if (!c->decoding) bit = (z >> k) & 1;
        bit = mixed_bit(c, &forest[u][node], &forest[v][node], wv, bit);
        if (c->decoding && bit) z += 1 << k;
        node += bit ? (1 << k) : 1;
    }
    return z;
}

/* ---- AVP: recursive weighted least squares in int64.  R: NBLIC.c:112-283 ---- */

static inline i64 avp_decay(i64 v, int slot) { /* forgetting: 2/3 for the energy slot, 4/5 otherwise */
    int ab = slot == 0 ? 3 : 5;
    return (wmul(v, ab - 1) + ab / 2) / ab;
}

static int avp_solve(int n, i64 *A, i64 *b) { /* R: NBLIC.c:112-161 */
    int k, r, c;
    for (k = 0; k + 1 < n; k++) {
        int piv = k;
        i64 d;
        for (r = k + 1; r < n; r++) {
            i64 x = A[r * n + k], y = A[piv * n + k];
            if ((x < 0 ? -x : x) > (y < 0 ? -y : y)) piv = r;
        }
        if (piv != k) {
            i64 t = b[k]; b[k] = b[piv]; b[piv] = t;
            for (c = k; c < n; c++) { t = A[k * n + c]; A[k * n + c] = A[piv * n + c]; A[piv * n + c] = t; }
        }
        d = A[k * n + k];
        if (d == 0) return 0;
        for (r = k + 1; r < n; r++) {
            i64 f = A[r * n + k];
            A[r * n + k] = 0;
            if (f == 0) continue;
            for (c = k + 1; c < n; c++) A[r * n + c] -= wmul(A[k * n + c], f) / d;
            b[r] -= wmul(b[k], f) / d;
        }
    }
    for (k = n - 1; k > 0; k--) {
        i64 d = A[k * n + k];
        if (d == 0) return 0;
        for (r = 0; r < k; r++) {
            i64 f = A[r * n + k];
            A[r * n + k] = 0;
            if (f != 0) b[r] -= wmul(b[k], f) / d;
        }
    }
    return 1;
}

static int avp_predict(int n, int m, const i64 *E, const i64 *F, const i64 *vec, i64 ridge, i64 *out) { /* R: NBLIC.c:210-239 */
    i64 ds[N_AVP_MAX_M], *b = ds + 1, *A = ds + 1 + n, px;
    int k;
    for (k = 1; k < m; k++) ds[k] = E[k] + F[k];
    for (k = 0; k < n; k++) { b[k] += wshl(ridge, N_FRAC - 2); A[k * n + k] += wmul(ridge, n); }
    if (!avp_solve(n, A, b)) return 0;
    px = (i64)128 << N_FRAC;
    for (k = 0; k < n; k++) {
        i64 d = A[k * n + k];
        px += (wshl(wmul(b[k], vec[k]), 2) + (d >> 1)) / d;
    }
    *out = clampl(px, 0, (i64)255 << N_FRAC);
    return 1;
}

static void avp_learn(int n, int m, i64 *E, i64 *B, const i64 *vec, int x, i64 s_now, i64 s_sum) { /* R: NBLIC.c:242-283 */
    i64 ds[N_AVP_MAX_M], *b = ds + 1, *A = ds + 1 + n, half;
    int r, c, k;
    ds[0] = s_now;
    x -= 128;
    s_sum = clampl(s_sum + (1 << N_FRAC), 1 << N_FRAC, 16 << N_FRAC);
    half = s_sum >> 1;
    for (k = 0; k < n; k++) b[k] = (wshl(wmul(x, vec[k]), 28) + half) / s_sum;
    for (r = 0; r < n; r++) for (c = 0; c < n; c++) A[r * n + c] = (wshl(wmul(vec[r], vec[c]), 18) + half) / s_sum;
    for (k = 0; k < m; k++) { B[k] = avp_decay(B[k], k) + ds[k]; E[k] = avp_decay(E[k], k) + B[k]; }
}

static int n_codec(int decoding, uint8_t *buf, long avail, uint8_t *img, int *ph, int *pw, int *pnear, int *peffort) {
    static const int N_OF_EFFORT[4] = {-1, 0, 6, 10};
    static const unsigned char VEC_ORDER[N_AVP_MAX_N] = {NA, NB, NC, ND, NE, NF, NT, NH, NQ, NG}; /* R: NBLIC.c:164-183 */
    uint8_t *p = buf;
    int channels = 1, k_step, h, w, near, effort, n, m, i, j, k;
    int *ctx = NULL;
    node_t (*forest)[256] = NULL;
    ranker_t (*rankers)[2] = NULL;
    i64 *Brow = NULL, *Frow = NULL, E[N_AVP_MAX_M], vec[N_AVP_MAX_N], ridge = 8;
    rc_t rc;
    int rv = -1;

    if (decoding) { /* R: NBLIC.c:698-712 */
        if (avail < 16 || memcmp(p, N_MAGIC, 8) != 0) return -1;
        p += 8;
        channels = *p++;
        *ph = (p[0] << 8) | p[1]; *pw = (p[2] << 8) | p[3]; p += 4;
        *pnear = *p++; k_step = *p++; *peffort = *p++;
    } else { /* R: NBLIC.c:682-694,768-771 -- the header is emitted before validation */
        *pnear = clampi(*pnear, 0, 9);
        k_step = clampi(3 + 2 * *pnear, 3, 16);
        *peffort = clampi(*peffort, 1, 3);
        memcpy(p, N_MAGIC, 8); p += 8;
        *p++ = (uint8_t)channels;
        *p++ = (uint8_t)(*ph >> 8); *p++ = (uint8_t)*ph; *p++ = (uint8_t)(*pw >> 8); *p++ = (uint8_t)*pw;
        *p++ = (uint8_t)*pnear; *p++ = (uint8_t)k_step; *p++ = (uint8_t)*peffort;
    }
    h = *ph; w = *pw; near = *pnear; effort = *peffort;
    if (!dims_ok(h, w) || channels < 0 || channels > 1 || near < 0 || near > 9 || k_step < 3 || k_step > 16 || effort < 1 || effort > 3)
        return -1; /* R: NBLIC.c:733-745 */

    n = N_OF_EFFORT[effort]; m = 1 + n + n * n;
    ctx = calloc((N_CLASSES / 2) * 256, sizeof(int));
    forest = malloc(N_CLASSES * sizeof *forest);
    rankers = malloc(256 * sizeof *rankers);
    if (n > 0) { Brow = calloc((size_t)w * m * 2, sizeof(i64)); Frow = Brow ? Brow + (size_t)w * m : NULL; }
    if (!ctx || !forest || !rankers || (n > 0 && !Brow)) goto done;
    for (i = 0; i < N_CLASSES; i++) for (j = 0; j < 256; j++) forest[i][j].n0 = forest[i][j].n1 = N_MIX;
    for (i = 0; i < 256; i++) { ranker_reset(&rankers[i][0]); ranker_reset(&rankers[i][1]); }

    memset(&rc, 0, sizeof rc);
    rc.hi = 0xffffffffu; rc.decoding = decoding;
    if (decoding) { rc.rd = p; rc.rd_left = avail - (p - buf); for (k = 0; k < 4; k++) rc.code = (rc.code << 8) | rc_read(&rc); }
    else rc.wr = p;

    for (i = 0; i < h; i++) { /* R: NBLIC.c:807-895 */
        int err = 0;
        if (n > 0) { /* R: NBLIC.c:186-204,817-820 */
            memset(E, 0, sizeof(i64) * m);
            for (j = w - 1; j >= 0; j--) for (k = 0; k < m; k++)
                Frow[(size_t)j * m + k] = (j == w - 1 ? 0 : avp_decay(Frow[(size_t)(j + 1) * m + k], k)) + Brow[(size_t)j * m + k];
        }
        for (j = 0; j < w; j++) {
            int nb[N_SLOTS], px0, px, sign, u, v, wv, adr, x, y, z = 0, ok1 = 0, ok2 = 0;
            i64 r1 = 0, r2 = 0, p1 = 0, p2 = 0, *B = NULL, *F = NULL;
            ranker_t *rk;

            sample_positional(img, w, i, j, nb);

            if (n > 0) { /* R: NBLIC.c:831-846 */
                for (k = 0; k < n; k++) vec[k] = nb[VEC_ORDER[k]] - 128;
                B = Brow + (size_t)j * m; F = Frow + (size_t)j * m;
                r1 = ridge * 21 / 22; r2 = ridge * 22 / 21;
                r1 = clampl(r1, -1, ridge - 1); r2 = clampl(r2, ridge + 1, N_BIAS_MAX + 1);
                r1 = clampl(r1, 0, N_BIAS_MAX); r2 = clampl(r2, 0, N_BIAS_MAX);
                ok1 = avp_predict(n, m, E, F, vec, r1, &p1);
                ok2 = avp_predict(n, m, E, F, vec, r2, &p2);
            }
            if (ok1) px0 = (int)((p1 + (1 << (N_FRAC - 1))) >> N_FRAC);
            else { pred_terms_t pt = predictor_terms(nb); px0 = blend_prediction(pt, n_weight(pt.spread)); p1 = (i64)px0 << N_FRAC; }

            n_soft_class(activity(nb, err), &u, &v, &wv);
            adr = ((u >> 1) << 8) | texture_bits(nb, px0);
            n_bias_apply(ctx[adr], px0, &px, &sign);
            rk = &rankers[px][sign];

            if (!decoding) {
                y = n_fold(img[(long)i * w + j], px, sign, near);
                z = y < N_RANKS ? rk->rank_of[y] : y;
            }
            z = golomb_symbol(&rc, k_step, forest, u, v, wv, z);
            if (decoding) y = z < N_RANKS ? rk->sym_at[z] : z;
            ranker_touch(rk, y);

            x = n_unfold(y, px, sign, near);
            img[(long)i * w + j] = (uint8_t)x;
            err = clampi(x - px0, -127, 127);
            ctx[adr] = n_bias_learn(ctx[adr], err);

            if (n > 0) { /* R: NBLIC.c:882-893 */
                i64 target = (i64)x << N_FRAC;
                i64 s_now = p1 - target < 0 ? target - p1 : p1 - target;
                i64 s_sum = (E[0] + F[0]) + s_now * 3 / 2;
                avp_learn(n, m, E, B, vec, x, s_now, s_sum);
                if (ok1 && ok2) {
                    i64 e2 = p2 - target < 0 ? target - p2 : p2 - target;
                    ridge = s_now > e2 ? r2 : r1;
                }
            }
        }
    }
    if (decoding) rv = 0;
    else { /* R: NBLIC.c:576-586 */
        for (k = 0; k < 4; k++) { *rc.wr++ = (uint8_t)(rc.lo >> 24); rc.lo <<= 8; }
        rv = (int)(rc.wr - buf);
    }
done:
    free(ctx); free(forest); free(rankers); free(Brow);
    return rv;
}

int oracle_n_encode(uint8_t *img, int height, int width, int *near, int *effort, uint8_t *out) {
    return n_codec(0, out, 0, img, &height, &width, near, effort);
}

int oracle_n_decode(const uint8_t *in, long in_avail, uint8_t *img, int *height, int *width, int *near, int *effort) {
    return n_codec(1, (uint8_t *)in, in_avail, img, height, width, near, effort);
}
