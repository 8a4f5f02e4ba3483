/*
 * codec_core.cuh -- device-side building blocks shared by every NBLIC / QNBLIC kernel.
 *
 * One *coder stream* (= one image, SURVEY.md section 0 correction 1) is advanced by one sequential
 * agent; which hardware unit plays the agent (a warp's leader lane with its adaptive state in
 * shared memory, or every lane of a warp with the state in L2) is decided by the kernels in
 * stream_kernels.cu.  Everything here is written against plain pointers so both mappings share it.
 *
 * Bit-exactness notes: C truncating '/', arithmetic '>>' on negatives,
 * two's-complement wrap on int64 products (wmul / wshl), no FMA contraction in the histogram
 * normaliser (__dmul_rn / __dadd_rn).
 *
 * Reference lines each block reproduces are cited as  R: <file>:<lines>  (paths under
 * /root/reference/src).
 */
#pragma once
#include <stdint.h>

namespace nblic {

typedef long long i64;
typedef unsigned long long u64;
typedef unsigned int u32;

#define NB_DEV __device__ __forceinline__

/* causal neighbourhood (R: NBLIC.c:287-304 / QNBLIC.c:48-64)
 *      s h f g r      row i-2
 *      q c b d t      row i-1
 *      e a X          row i      */
struct Nb { int a, b, c, d, e, f, g, h, q, r, s, t; };

NB_DEV int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }
NB_DEV i64 clampl(i64 v, i64 lo, i64 hi) { return v < lo ? lo : (v > hi ? hi : v); }
NB_DEV i64 wmul(i64 a, i64 b) { return (i64)((u64)a * (u64)b); }
NB_DEV i64 wshl(i64 a, int s) { return (i64)((u64)a << s); }
NB_DEV i64 labs64(i64 v) { return v < 0 ? -v : v; }

NB_DEV bool dims_ok(int h, int w) { /* R: NBLIC.c:717-729, QNBLIC.c:33-45 */
    return h > 0 && w > 0 && h <= 65535 && w <= 65535 && (i64)h * w <= 100000000LL;
}

/* Positional sampling with the reference's cascading border fallbacks.  R: NBLIC.c:287-304. */
NB_DEV void sample_positional(const uint8_t *img, int w, int i, int j, Nb &n) {
    const uint8_t *r0 = img + (size_t)i * w, *r1 = r0 - w, *r2 = r1 - w;
    const bool up1 = i >= 1, up2 = i >= 2;
    const bool l1 = j >= 1, l2 = j >= 2, rt1 = j + 1 < w, rt2 = j + 2 < w;
    int a = l1 ? (int)r0[j - 1] : 128;
    int b = up1 ? (int)r1[j] : 128;
    if (i == 0) b = a; else if (j == 0) a = b;
    n.a = a; n.b = b;
    n.e = l2 ? (int)r0[j - 2] : a;
    n.c = (up1 && l1) ? (int)r1[j - 1] : b;
    n.d = (up1 && rt1) ? (int)r1[j + 1] : b;
    n.f = up2 ? (int)r2[j] : b;
    n.g = (up2 && rt1) ? (int)r2[j + 1] : n.f;
    n.h = (up2 && l1) ? (int)r2[j - 1] : n.f;
    n.q = (up1 && l2) ? (int)r1[j - 2] : n.c;
    n.r = (up2 && rt2) ? (int)r2[j + 2] : n.g;
    n.s = (up2 && l2) ? (int)r2[j - 2] : n.h;
    n.t = (up1 && rt2) ? (int)r1[j + 2] : n.d;
}

/* Interior step of the positional window: valid when the new column jn satisfies i >= 2,
 * 2 <= jn <= w-3 (every slot of both windows is a real pixel, so sliding == re-sampling). */
NB_DEV void slide_interior(const uint8_t *img, int w, int i, int jn, int x, Nb &n) {
    const uint8_t *r1 = img + (size_t)(i - 1) * w, *r2 = r1 - w;
    n.e = n.a; n.a = x;
    n.q = n.c; n.c = n.b; n.b = n.d; n.d = n.t; n.t = r1[jn + 2];
    n.s = n.h; n.h = n.f; n.f = n.g; n.g = n.r; n.r = r2[jn + 2];
}

/* The 7-direction gradient predictor.  R: NBLIC.c:307-370, QNBLIC.c:94-149.
 * Each directional cost is sum over the four anchors (a, c, b, d) of |2*anchor - p - p'| where p, p'
 * are the anchor's neighbours one step along the direction pair; written here as sums of the
 * one-step differences west / north-west / north / north-east. */
struct Pred { int ang2, lin16, spread; };

NB_DEV Pred predictor_terms(const Nb &n) {
    const int w0 = n.a - n.e, w1 = n.c - n.q, w2 = n.b - n.c, w3 = n.d - n.b;
    const int x0 = n.a - n.q, x1 = n.c - n.s, x2 = n.b - n.h, x3 = n.d - n.f;
    const int y0 = n.a - n.c, y1 = n.c - n.h, y2 = n.b - n.f, y3 = n.d - n.g;
    const int z0 = n.a - n.b, z1 = n.c - n.f, z2 = n.b - n.g, z3 = n.d - n.r;
    int cost[7], src[7];
    cost[0] = 2 * (abs(w0) + abs(w1) + abs(w2) + abs(w3));                 src[0] = 2 * n.a;
    cost[1] = 2 * (abs(y0) + abs(y1) + abs(y2) + abs(y3));                 src[1] = 2 * n.b;
    cost[2] = 2 * (abs(x0) + abs(x1) + abs(x2) + abs(x3));                 src[2] = 2 * n.c;
    cost[3] = 2 * (abs(z0) + abs(z1) + abs(z2) + abs(z3));                 src[3] = 2 * n.d;
    cost[4] = abs(w0 + x0) + abs(w1 + x1) + abs(w2 + x2) + abs(w3 + x3);   src[4] = n.a + n.c;
    cost[5] = abs(x0 + y0) + abs(x1 + y1) + abs(x2 + y2) + abs(x3 + y3);   src[5] = n.c + n.b;
    cost[6] = abs(y0 + z0) + abs(y1 + z1) + abs(y2 + z2) + abs(y3 + z3);   src[6] = n.b + n.d;
    Pred r;
    int best = cost[0], total = cost[0];
    r.ang2 = src[0];
#pragma unroll
    for (int k = 1; k < 7; k++) {
        total += cost[k];
        if (cost[k] < best) { best = cost[k]; r.ang2 = src[k]; } /* first minimum wins */
    }
    r.lin16 = clampi(9 * n.a + 9 * n.b + 2 * n.d - 2 * n.c - n.e - n.f, 0, 16 * 255);
    r.spread = total - 7 * best;
    return r;
}

NB_DEV int blend_prediction(const Pred &p, int wt) { return (8 * wt * p.ang2 + (8 - wt) * p.lin16 + 64) >> 7; }

NB_DEV int activity(const Nb &n, int err) { /* R: NBLIC.c:376, QNBLIC.c:531 */
    return abs(n.a - n.e) + abs(n.b - n.c) + abs(n.b - n.d) + abs(n.a - n.c) + abs(n.b - n.f) + abs(n.d - n.g) + 2 * abs(err);
}

/* The eight texture comparisons, a in bit 0 (NBLIC order, R: NBLIC.c:401-408). */
NB_DEV int texture_bits(const Nb &n, int px) {
    int t = 0;
    t |= (px > n.a) << 0;
    t |= (px > n.b) << 1;
    t |= (px > n.c) << 2;
    t |= (px > n.d) << 3;
    t |= (px > n.e) << 4;
    t |= (px > n.f) << 5;
    t |= (px > 2 * n.a - n.e) << 6;
    t |= (px > 2 * n.b - n.f) << 7;
    return t;
}

/* ============================================================================================ */
/* NBLIC ("NBLIC0.3")                                                                           */
/* ============================================================================================ */

enum { N_CLASSES = 16, N_CTX_SHIFT = 8, N_MIX = 32, N_RANKS = 20, N_PROB_ONE = 4096, N_FRAC = 12, N_BIAS_MAX = 4096 };

/* adaptive state of one NBLIC stream */
struct NState {
    int16_t *ctx;       /* [8*256]   bias-cancel table, |v| <= 127*256 (R: NBLIC.c:60-64,421-428)        */
    u32 *forest;        /* [16*256]  node counters packed n0 | n1 << 16, each <= 8224 (R: NBLIC.c:589-617) */
    uint8_t *rank_of;   /* [512*20]  symbol -> rank      (R: NBLIC.c:470-523)                              */
    uint8_t *sym_at;    /* [512*20]  rank -> symbol                                                       */
    int *count;         /* [512*20]  running frequencies (unbounded, kept 32-bit)                          */
    i64 *Brow, *Frow;   /* [w*m]     AVP per-column accumulators (effort 2/3 only)                         */
};
enum { N_CTX_ENTRIES = 8 * 256, N_FOREST_ENTRIES = 16 * 256, N_RANK_ENTRIES = 512 * N_RANKS };

NB_DEV void nstate_reset(const NState &s, int lane, int nl) {
    for (int k = lane; k < N_CTX_ENTRIES; k += nl) s.ctx[k] = 0;
    for (int k = lane; k < N_FOREST_ENTRIES; k += nl) s.forest[k] = (u32)N_MIX | ((u32)N_MIX << 16);
    for (int k = lane; k < N_RANK_ENTRIES; k += nl) {
        int r = k % N_RANKS;
        s.rank_of[k] = (uint8_t)r; s.sym_at[k] = (uint8_t)r; s.count[k] = 2 * (N_RANKS - 1 - r);
    }
}

NB_DEV int n_weight(int spread) { /* R: NBLIC.c:308,365-367: number of thresholds {31,93,279,620,1550,3410,9300,24800} <= spread */
    int w = spread >= 620 ? 4 : 0;                       /* three-level search instead of eight compares */
    w += spread >= (w ? 3410 : 93) ? 2 : 0;
    w += spread >= (w == 0 ? 31 : w == 2 ? 279 : w == 4 ? 1550 : 9300);
    return w + (spread >= 24800);
}

/* Soft 16-class activity quantiser.  R: NBLIC.c:373-395 */
NB_DEV void n_soft_class(int delta, int &u, int &v, int &wv) {
    const int MID[N_CLASSES] = {0, 2, 4, 7, 10, 14, 20, 26, 34, 42, 52, 64, 78, 95, 135, 200};
    int c = 0, lo = 0, hi = 0;
#pragma unroll
    for (int k = 0; k < N_CLASSES - 1; k++) {
        const bool above = delta > MID[k];
        c += above;
        if (above) { lo = MID[k]; hi = MID[k + 1]; }
    }
    u = v = c; wv = 0;
    if (delta < hi) { /* only reachable with c >= 1, where lo = mid[c-1], hi = mid[c] */
        const int w = N_MIX * (delta - lo) / (hi - lo);
        if (w < N_MIX / 2) { u = c - 1; wv = w; } else { v = c - 1; wv = N_MIX - w; }
    }
}

NB_DEV void n_bias_apply(int ctx, int px0, int &px, int &sign) { /* R: NBLIC.c:413-418 */
    sign = (ctx >> (N_CTX_SHIFT - 1)) & 1;
    px = clampi(px0 + (ctx >> N_CTX_SHIFT) + sign, 0, 255);
}
NB_DEV int n_bias_learn(int ctx, int err) { /* R: NBLIC.c:421-428 */
    return (ctx * 127 + err * (1 << N_CTX_SHIFT) + 64) >> 7;
}

/* near-aware residual fold / unfold.  R: NBLIC.c:431-466 */
NB_DEV int n_fold(int x, int px, int sign, int near) {
    const int q = 2 * near + 1;
    const int room = (min(px, 255 - px) + near) / q;
    const int mag = (abs(x - px) + near) / q;
    if (mag <= 0) return 0;
    if (mag <= room) return 2 * mag - ((x >= px) ^ sign);
    return mag + room;
}
NB_DEV int n_unfold(int y, int px, int sign, int near) {
    const int q = 2 * near + 1;
    const int room = (min(px, 255 - px) + near) / q;
    int mag, up;
    if (y <= 0) { mag = 0; up = 0; }
    else if (y <= 2 * room) { mag = (y + 1) >> 1; up = (y & 1) ^ sign; }
    else { mag = y - room; up = px < 128; }
    mag *= q;
    return clampi(up ? px + mag : px - mag, 0, 255);
}

/* carry-less 32-bit binary range coder.  R: NBLIC.c:527-586 */
template <bool DEC> struct RangeCoder;

template <> struct RangeCoder<false> {
    u32 lo, hi;
    uint8_t *wr, *wr_end;
    bool overflow;
    NB_DEV void start(uint8_t *p, uint8_t *end) { lo = 0; hi = 0xffffffffu; wr = p; wr_end = end; overflow = false; }
    NB_DEV void put(u32 byte) { if (wr < wr_end) *wr = (uint8_t)byte; else overflow = true; wr++; }
    NB_DEV int bit(int b, u32 p1) {
        const u32 span = hi - lo;
        const u32 mid = lo + (span >> 12) * p1 + (((span & 0xfffu) * p1) >> 12);
        if (b) hi = mid; else lo = mid + 1;
        while (((lo ^ hi) & 0xff000000u) == 0) { put(hi >> 24); lo <<= 8; hi = (hi << 8) | 0xffu; }
        return b;
    }
    NB_DEV void finish() { for (int k = 0; k < 4; k++) { put(lo >> 24); lo <<= 8; } } /* R: NBLIC.c:576-586 */
};

template <> struct RangeCoder<true> {
    u32 lo, hi, code;
    const uint8_t *rd, *rd_end;
    NB_DEV u32 get() { u32 v = rd < rd_end ? (u32)*rd : 0u; rd++; return v; }
    NB_DEV void start(const uint8_t *p, const uint8_t *end) {
        lo = 0; hi = 0xffffffffu; rd = p; rd_end = end; code = 0;
        for (int k = 0; k < 4; k++) code = (code << 8) | get();
    }
    NB_DEV int bit(int, u32 p1) {
        const u32 span = hi - lo;
        const u32 mid = lo + (span >> 12) * p1 + (((span & 0xfffu) * p1) >> 12);
        const int b = code <= mid;
        if (b) hi = mid; else lo = mid + 1;
        while (((lo ^ hi) & 0xff000000u) == 0) { code = (code << 8) | get(); lo <<= 8; hi = (hi << 8) | 0xffu; }
        return b;
    }
};

/* Weighted pair of hit counters per tree node, two-model mix.  R: NBLIC.c:589-637 */
NB_DEV int node_p1(u32 packed) { return (int)(((packed >> 16) * (u32)N_PROB_ONE) / ((packed & 0xffffu) + (packed >> 16))); }
NB_DEV u32 node_learn(u32 packed, int bit, int weight) {
    u32 n0 = packed & 0xffffu, n1 = packed >> 16;
    if (bit) n1 += weight; else n0 += weight;
    if (n0 + n1 > (u32)(N_MIX * 256)) { n0 = (n0 + 1) >> 1; n1 = (n1 + 1) >> 1; }
    return n0 | (n1 << 16);
}

template <bool DEC>
NB_DEV int mixed_bit(RangeCoder<DEC> &rc, u32 *nu, u32 *nv, int wv, int bit) {
    const u32 cu = *nu;
    if (nu == nv) { /* both models are the same node: weights 32-wv and wv land on one counter pair */
        const int p = node_p1(cu); /* (p*(32-wv) + p*wv + 16)/32 == p */
        bit = rc.bit(bit, (u32)clampi(p, 1, N_PROB_ONE - 1));
        *nu = node_learn(node_learn(cu, bit, N_MIX - wv), bit, wv);
        return bit;
    }
    const u32 cv = *nv;
    const int p = (node_p1(cu) * (N_MIX - wv) + node_p1(cv) * wv + N_MIX / 2) / N_MIX;
    bit = rc.bit(bit, (u32)clampi(p, 1, N_PROB_ONE - 1));
    *nu = node_learn(cu, bit, N_MIX - wv);
    *nv = node_learn(cv, bit, wv);
    return bit;
}

/* adaptive-Golomb binarisation over the 16x256 node forest.  R: NBLIC.c:640-679 */
/* Returns the symbol, or -1 when a (corrupt) stream escapes past the last Golomb order -- a state no
 * encoder output reaches (the reference indexes out of bounds there); valid streams are unaffected. */
template <bool DEC>
NB_DEV int golomb_symbol(RangeCoder<DEC> &rc, int k_step, u32 *forest, int u, int v, int wv, int z, int node = 0) {
    const int top = (N_CLASSES - 1) / k_step;
    int k, bit = 0;
    if (v / k_step != u / k_step) v = u;
    for (;;) {
        k = u / k_step;
        if (!DEC) bit = (node >> top) < (z >> k);
        bit = mixed_bit<DEC>(rc, forest + u * 256 + node, forest + v * 256 + node, wv, bit);
        if (!bit) break;
        node += 1 << top;
        if (node >= 256) { /* escape to the next order */
            node >>= 1; u = v = (k + 1) * k_step;
            if (u >= N_CLASSES) return -1;
        }
    }
    if (DEC) z = (node >> top) << k;
    for (node++, k--; k >= 0; k--) {
        if (!DEC) bit = (z >> k) & 1;
        const int at = DEC ? (node & 255) : node;
        bit = mixed_bit<DEC>(rc, forest + u * 256 + at, forest + v * 256 + at, wv, bit);
        if (DEC && bit) z += 1 << k;
        node += bit ? (1 << k) : 1;
    }
    return z;
}

/* adaptive rank mapper.  R: NBLIC.c:470-523 */
NB_DEV void ranker_touch(uint8_t *rank_of, uint8_t *sym_at, int *count, int y) {
    if (y >= N_RANKS) return;
    const int z = rank_of[y];
    const int cz = count[z] + 1;
    count[z] = cz;
    if (z > 0) {
        const int cp = count[z - 1];
        if (cp < cz) { /* one adjacent promotion */
            const int other = sym_at[z - 1];
            count[z] = cp; count[z - 1] = cz;
            sym_at[z] = (uint8_t)other; sym_at[z - 1] = (uint8_t)y;
            rank_of[y] = (uint8_t)(z - 1); rank_of[other] = (uint8_t)z;
        }
    }
}

/* ---- AVP: recursive weighted least squares in int64.  R: NBLIC.c:112-283 -------------------- */

NB_DEV i64 avp_decay(i64 v, int slot) { /* forgetting: 2/3 for the energy slot, 4/5 otherwise */
    return slot == 0 ? (wmul(v, 2) + 1) / 3 : (wmul(v, 4) + 2) / 5;
}

template <int N>
__device__ __noinline__ int avp_solve(i64 *A, i64 *b) { /* R: NBLIC.c:112-161 */
    for (int k = 0; k + 1 < N; k++) {
        int piv = k;
        i64 best = labs64(A[k * N + k]);
        for (int r = k + 1; r < N; r++) {
            const i64 m = labs64(A[r * N + k]);
            if (m > best) { best = m; piv = r; }
        }
        if (piv != k) {
            i64 t = b[k]; b[k] = b[piv]; b[piv] = t;
            for (int c = k; c < N; c++) { t = A[k * N + c]; A[k * N + c] = A[piv * N + c]; A[piv * N + c] = t; }
        }
        const i64 d = A[k * N + k];
        if (d == 0) return 0;
        for (int r = k + 1; r < N; r++) {
            const i64 f = A[r * N + k];
            A[r * N + k] = 0;
            if (f == 0) continue;
            for (int c = k + 1; c < N; c++) A[r * N + c] -= wmul(A[k * N + c], f) / d;
            b[r] -= wmul(b[k], f) / d;
        }
    }
    for (int k = N - 1; k > 0; k--) {
        const i64 d = A[k * N + k];
        if (d == 0) return 0;
        for (int r = 0; r < k; r++) {
            const i64 f = A[r * N + k];
            A[r * N + k] = 0;
            if (f != 0) b[r] -= wmul(b[k], f) / d;
        }
    }
    return 1;
}

template <int N>
NB_DEV int avp_predict(const i64 *E, const i64 *F, const i64 *vec, i64 ridge, i64 &out) { /* R: NBLIC.c:210-239 */
    constexpr int M = 1 + N + N * N;
    i64 ds[M];
    i64 *b = ds + 1, *A = ds + 1 + N;
    for (int k = 1; k < M; k++) ds[k] = E[k] + F[k];
    for (int k = 0; k < N; k++) { b[k] += wshl(ridge, N_FRAC - 2); A[k * N + k] += wmul(ridge, N); }
    if (!avp_solve<N>(A, b)) return 0;
    i64 px = (i64)128 << N_FRAC;
    for (int k = 0; k < N; k++) {
        const i64 d = A[k * N + k];
        px += (wshl(wmul(b[k], vec[k]), 2) + (d >> 1)) / d;
    }
    out = clampl(px, 0, (i64)255 << N_FRAC);
    return 1;
}

template <int N>
NB_DEV void avp_learn(i64 *E, i64 *B, const i64 *vec, int x, i64 s_now, i64 s_sum) { /* R: NBLIC.c:242-283 */
    x -= 128;
    s_sum = clampl(s_sum + (1 << N_FRAC), 1 << N_FRAC, 16 << N_FRAC);
    const i64 half = s_sum >> 1;
    { const i64 nb = avp_decay(B[0], 0) + s_now; B[0] = nb; E[0] = avp_decay(E[0], 0) + nb; }
    for (int k = 0; k < N; k++) {
        const i64 t = (wshl(wmul(x, vec[k]), 28) + half) / s_sum;
        const i64 nb = avp_decay(B[1 + k], 1) + t; B[1 + k] = nb; E[1 + k] = avp_decay(E[1 + k], 1) + nb;
    }
    for (int r = 0; r < N; r++) for (int c = 0; c < N; c++) {
        const int k = 1 + N + r * N + c;
        const i64 t = (wshl(wmul(vec[r], vec[c]), 18) + half) / s_sum;
        const i64 nb = avp_decay(B[k], 1) + t; B[k] = nb; E[k] = avp_decay(E[k], 1) + nb;
    }
}

struct NJob {
    const uint8_t *src;   /* original pixels (encode) */
    uint8_t *rec;         /* reconstruction / decoded raster; may alias nothing (NULL) for lossless encode */
    uint8_t *stream;      /* encode: output slot; decode: input bytes */
    u32 stream_cap;       /* encode: capacity; decode: valid bytes */
    int h, w, near, k_step;
};

/* One NBLIC stream, start to finish.  Returns encode: bytes written (or ~0u on overflow); decode: 0,
 * or 1 for a corrupt stream.  R: NBLIC.c:749-908 */
template <int NAVP, bool DEC>
__device__ u32 nblic_stream(const NJob &job, const NState &st) {
    constexpr int N = NAVP, M = 1 + N + N * N;
    const int h = job.h, w = job.w, near = job.near, k_step = job.k_step;
    const uint8_t *nbimg = (!DEC && near == 0) ? job.src : job.rec;
    uint8_t *rec = (!DEC && near == 0) ? nullptr : job.rec;

    RangeCoder<DEC> rc;
    if constexpr (DEC) {
        rc.start(job.stream + 16, job.stream + job.stream_cap);
    } else { /* R: NBLIC.c:682-694 */
        uint8_t *p = job.stream;
        const char magic[8] = {'N', 'B', 'L', 'I', 'C', '0', '.', '3'};
        for (int k = 0; k < 8; k++) p[k] = (uint8_t)magic[k];
        p[8] = 1; p[9] = (uint8_t)(h >> 8); p[10] = (uint8_t)h; p[11] = (uint8_t)(w >> 8); p[12] = (uint8_t)w;
        p[13] = (uint8_t)near; p[14] = (uint8_t)k_step; p[15] = (uint8_t)(N == 0 ? 1 : (N == 6 ? 2 : 3));
        rc.start(p + 16, p + job.stream_cap);
    }

    i64 E[M > 1 ? M : 1], vec[N > 0 ? N : 1], ridge = 8;

    for (int i = 0; i < h; i++) {
        int err = 0;
        if constexpr (N > 0) { /* R: NBLIC.c:186-204,817-820 */
            for (int k = 0; k < M; k++) E[k] = 0;
            for (int j = w - 1; j >= 0; j--) {
                i64 *F = st.Frow + (size_t)j * M;
                const i64 *B = st.Brow + (size_t)j * M;
                if (j == w - 1) { for (int k = 0; k < M; k++) F[k] = B[k]; }
                else { const i64 *Fn = F + M; F[0] = avp_decay(Fn[0], 0) + B[0]; for (int k = 1; k < M; k++) F[k] = avp_decay(Fn[k], 1) + B[k]; }
            }
        }
        Nb nb;
        int x = 0;
        for (int j = 0; j < w; j++) {
            if (i >= 2 && j >= 2 && j + 2 < w) slide_interior(nbimg, w, i, j, x, nb);
            else sample_positional(nbimg, w, i, j, nb);

            int px0, ok1 = 0, ok2 = 0;
            i64 r1 = 0, r2 = 0, p1 = 0, p2 = 0;
            if constexpr (N > 0) { /* R: NBLIC.c:164-183,831-846 */
                vec[0] = nb.a - 128; vec[1] = nb.b - 128; vec[2] = nb.c - 128; vec[3] = nb.d - 128; vec[4] = nb.e - 128; vec[5] = nb.f - 128;
                if constexpr (N > 6) { vec[6] = nb.t - 128; vec[7] = nb.h - 128; vec[8] = nb.q - 128; vec[9] = nb.g - 128; }
                const i64 *F = st.Frow + (size_t)j * M;
                r1 = ridge * 21 / 22; r2 = ridge * 22 / 21;
                r1 = clampl(r1, -1, ridge - 1); r2 = clampl(r2, ridge + 1, N_BIAS_MAX + 1);
                r1 = clampl(r1, 0, N_BIAS_MAX); r2 = clampl(r2, 0, N_BIAS_MAX);
                ok1 = avp_predict<N>(E, F, vec, r1, p1);
                ok2 = avp_predict<N>(E, F, vec, r2, p2);
            }
            if (ok1) px0 = (int)((p1 + (1 << (N_FRAC - 1))) >> N_FRAC);
            else { const Pred pt = predictor_terms(nb); px0 = blend_prediction(pt, n_weight(pt.spread)); p1 = (i64)px0 << N_FRAC; }

            int u, v, wv, px, sign;
            n_soft_class(activity(nb, err), u, v, wv);
            const int adr = ((u >> 1) << 8) | texture_bits(nb, px0);
            const int ctx = st.ctx[adr];
            n_bias_apply(ctx, px0, px, sign);
            const int key = ((px << 1) | sign) * N_RANKS;

            int y = 0, z = 0;
            if (!DEC) {
                y = n_fold(job.src[(size_t)i * w + j], px, sign, near);
                z = y < N_RANKS ? (int)st.rank_of[key + y] : y;
            }
            z = golomb_symbol<DEC>(rc, k_step, st.forest, u, v, wv, z);
            if (z < 0) return DEC ? 1u : 0xffffffffu;
            if (DEC) y = z < N_RANKS ? (int)st.sym_at[key + z] : z;
            ranker_touch(st.rank_of + key, st.sym_at + key, st.count + key, y);

            x = n_unfold(y, px, sign, near);
            if (rec) rec[(size_t)i * w + j] = (uint8_t)x;
            err = clampi(x - px0, -127, 127);
            st.ctx[adr] = (int16_t)n_bias_learn(ctx, err);

            if constexpr (N > 0) { /* R: NBLIC.c:882-893 */
                const i64 target = (i64)x << N_FRAC;
                const i64 s_now = labs64(p1 - target);
                const i64 s_sum = (E[0] + st.Frow[(size_t)j * M]) + s_now * 3 / 2;
                avp_learn<N>(E, st.Brow + (size_t)j * M, vec, x, s_now, s_sum);
                if (ok1 && ok2) ridge = s_now > labs64(p2 - target) ? r2 : r1;
            }
        }
    }
    if constexpr (DEC) return 0;
    else {
        rc.finish();
        return rc.overflow ? 0xffffffffu : (u32)(rc.wr - job.stream);
    }
}

/* ============================================================================================ */
/* QNBLIC ("Q0.2")                                                                              */
/* ============================================================================================ */

enum { Q_CLASSES = 12, Q_CTX_SHIFT = 11, Q_NORM_BITS = 15, Q_NORM_SUM = 1 << Q_NORM_BITS };
enum { Q_CTX_ENTRIES = Q_CLASSES * 256, Q_TAB_ENTRIES = Q_CLASSES * 256 };

struct QState {
    int *ctx;   /* [12*256] bias-cancel table, |v| <= 255*2048 (R: QNBLIC.c:24-28,183-188)                      */
    u32 *tab;   /* [12*256] encode pass 1: raw counts; afterwards freq | cumulative << 16 (R: QNBLIC.c:290-358) */
};

NB_DEV int q_weight(int spread) { /* R: QNBLIC.c:82-91,144-146 */
    const int s = min(spread >> 3, 607);
    return (s >= 5) + (s >= 12) + (s >= 34) + (s >= 78) + (s >= 194) + (s >= 431) + (s >= 601);
}
NB_DEV int q_class(int delta) { /* R: QNBLIC.c:152-161,532-533 */
    return (delta >= 1) + (delta >= 2) + (delta >= 4) + (delta >= 6) + (delta >= 9) + (delta >= 15) + (delta >= 25) + (delta >= 39) +
           (delta >= 63) + (delta >= 101) + (delta >= 151);
}
/* same eight comparisons as NBLIC but packed MSB-first.  R: QNBLIC.c:164-173 */
NB_DEV int q_ctx_address(const Nb &n, int px, int cls) {
    return (cls << 8) | (int)(__brev((u32)texture_bits(n, px)) >> 24);
}
/* QNBLIC's literal shift register; fresh taps (i-1, j+2) and (i-2, j+3).  R: QNBLIC.c:67-79 */
NB_DEV void q_window_shift(const uint8_t *img, int w, int i, int j, int x, Nb &n) {
    const int old_d = n.d, old_r = n.r;
    n.e = n.a; n.a = x;
    n.q = n.c; n.c = n.b; n.b = old_d;
    n.s = n.h; n.h = n.f; n.f = n.g; n.g = old_r;
    if (i <= 0) n.d = n.a;
    else if (j + 2 < w) n.d = img[(size_t)(i - 1) * w + j + 2];
    if (i <= 1) n.r = n.d;
    else if (j + 3 < w) n.r = img[(size_t)(i - 2) * w + j + 3];
}
NB_DEV void q_bias_apply(int ctx, int px0, int &px, int &sign) { /* R: QNBLIC.c:176-180 */
    sign = (ctx >> (Q_CTX_SHIFT - 1)) & 1;
    px = clampi(px0 + (ctx >> Q_CTX_SHIFT) + sign, 0, 255);
}
NB_DEV int q_bias_learn(int ctx, int err) { return (ctx * 127 + err * (1 << Q_CTX_SHIFT) + 63) >> 7; } /* R: QNBLIC.c:183-188 */

NB_DEV int q_fold(int x, int px, int sign) { /* R: QNBLIC.c:191-202 */
    const int room = min(px, 255 - px), mag = abs(x - px);
    if (mag == 0) return 0;
    if (mag <= room) return 2 * mag - ((x >= px) ^ sign);
    return mag + room;
}
NB_DEV int q_unfold(int y, int px, int sign) { /* R: QNBLIC.c:205-217 */
    const int room = min(px, 255 - px);
    if (y <= 0) return px;
    if (y <= 2 * room) { const int mag = (y + 1) >> 1; return ((y & 1) ^ sign) ? px + mag : px - mag; }
    return px < 128 ? px + (y - room) : px - (y - room);
}

/* Normalise one 256-bin histogram to sum 2^15 (double arithmetic, no FMA).  R: QNBLIC.c:308-358 */
NB_DEV void q_normalise(u32 *hist) {
    u32 total = 0, live = 0, last = 0;
    for (u32 k = 0; k < 256; k++) if (hist[k]) { total += hist[k]; live++; last = k; }
    if (live == 0) { hist[0] = Q_NORM_SUM - 1; hist[1] = 1; return; }
    if (live == 1) { hist[last] = Q_NORM_SUM - 1; hist[(last + 1) & 255] = 1; return; }
    const double scale = __ddiv_rn((double)Q_NORM_SUM, (double)total);
    u32 sum = 0;
    for (u32 k = 0; k < 256; k++) if (hist[k]) {
        const u32 v = (u32)__dadd_rn(0.49, __dmul_rn(scale, (double)hist[k]));
        hist[k] = v ? v : 1;
        sum += hist[k];
    }
    for (u32 k = 0; sum > Q_NORM_SUM; k = (k + 1) & 255) if (hist[k] > 1) { hist[k]--; sum--; }
    for (u32 k = 0; sum < Q_NORM_SUM; k = (k + 1) & 255) if (hist[k] > 0) { hist[k]++; sum++; }
}

/* freq -> freq | cumulative << 16, in place.  R: QNBLIC.c:290-295 */
NB_DEV void q_pack_cumulative(u32 *hist) {
    u32 acc = 0;
    for (int k = 0; k < 256; k++) { const u32 f = hist[k]; hist[k] = f | (acc << 16); acc += f; }
}

/* Histogram side information: 16-bit codes in five shapes.  R: QNBLIC.c:362-412.  tab holds freq in
 * the low half.  Returns the advanced output index (words); never writes at or beyond cap. */
NB_DEV u32 q_put_hist(uint16_t *out, u32 o, u32 cap, const u32 *tab) {
    u32 pos = 0, sum = 0;
    while (pos < 256 && sum < Q_NORM_SUM) {
        const u32 head = tab[pos] & 0xffffu;
        u32 stop = pos + 1, follower = 0xffff, next, code;
        while (stop < 256) { follower = tab[stop] & 0xffffu; if (follower != head) break; stop++; }
        if (head <= 1 && stop - pos >= 4) {
            const u32 run = stop - pos;
            next = stop;
            if (stop < 256 && follower <= 15) next = stop + 1; else follower = head;
            code = 0xE000u | (head << 12) | (follower << 8) | (run - 4);
        } else {
            const u32 h1 = pos + 1 < 256 ? (tab[pos + 1] & 0xffffu) : 0xffffu;
            const u32 h2 = pos + 2 < 256 ? (tab[pos + 2] & 0xffffu) : 0xffffu;
            const u32 h3 = pos + 3 < 256 ? (tab[pos + 3] & 0xffffu) : 0xffffu;
            if (head <= 7 && h1 <= 7 && h2 <= 7 && h3 <= 7) { code = 0xD000u | (head << 9) | (h1 << 6) | (h2 << 3) | h3; next = pos + 4; }
            else if (head <= 15 && h1 <= 15 && h2 <= 15)   { code = 0xC000u | (head << 8) | (h1 << 4) | h2; next = pos + 3; }
            else if (head <= 127 && h1 <= 127)             { code = 0x8000u | (head << 7) | h1; next = pos + 2; }
            else                                           { code = head; next = pos + 1; }
        }
        if (o < cap) out[o] = (uint16_t)code;
        o++;
        for (; pos < next; pos++) sum += tab[pos] & 0xffffu;
    }
    return o;
}

struct QJob {
    const uint8_t *src;   /* encode: pixels */
    uint8_t *rec;         /* decode: raster out */
    uint16_t *stream;     /* 2-byte aligned */
    u32 stream_cap_words; /* encode: capacity; decode: valid words */
    uint8_t *sym;         /* encode scratch: (class, y) per pixel, raster order */
    int h, w;
};

/* QNBLIC encode.  The head (magic, dims, 12 histogram descriptions) is written at the start of the slot,
 * the rANS words downwards from its end -- which is already the reversed order the container wants
 * (R: QNBLIC.c:250-260,647-649); the gather kernel joins the two pieces.
 * Returns false on overflow.  R: QNBLIC.c:562-655 */
__device__ bool qnblic_encode_stream(const QJob &job, const QState &st, u32 &head_words, u32 &tail_words) {
    const int h = job.h, w = job.w;
    const uint8_t *img = job.src;
    for (int k = 0; k < Q_CTX_ENTRIES; k++) st.ctx[k] = 0;
    for (int k = 0; k < Q_TAB_ENTRIES; k++) st.tab[k] = 0;

    size_t idx = 0;
    for (int i = 0; i < h; i++) { /* pass 1: model every pixel.  R: QNBLIC.c:586-623 */
        Nb nb;
        int err = 0;
        sample_positional(img, w, i, 0, nb);
        for (int j = 0; j < w; j++) {
            const int x = img[(size_t)i * w + j];
            const Pred pt = predictor_terms(nb);
            const int px0 = blend_prediction(pt, q_weight(pt.spread));
            const int cls = q_class(activity(nb, err));
            err = x - px0;
            const int adr = q_ctx_address(nb, px0, cls);
            const int ctx = st.ctx[adr];
            int px, sign;
            q_bias_apply(ctx, px0, px, sign);
            const int y = q_fold(x, px, sign);
            job.sym[idx++] = (uint8_t)cls; job.sym[idx++] = (uint8_t)y;
            st.tab[cls * 256 + y]++;
            st.ctx[adr] = q_bias_learn(ctx, err);
            q_window_shift(img, w, i, j, x, nb);
        }
    }

    uint16_t *out = job.stream;
    const u32 cap = job.stream_cap_words;
    if (cap < 8) return false;
    out[0] = 0x3051; out[1] = 0x322e; out[2] = (uint16_t)h; out[3] = (uint16_t)w; /* R: QNBLIC.c:463-473 */
    u32 o = 4;
    for (int c = 0; c < Q_CLASSES; c++) {
        q_normalise(st.tab + c * 256);
        o = q_put_hist(out, o, cap, st.tab + c * 256);
        q_pack_cumulative(st.tab + c * 256);
    }
    if (o >= cap) return false;

    /* pass 2: rANS, last pixel first.  R: QNBLIC.c:238-253,635-650 */
    u32 state = 1u << 16, p = cap;
    bool ok = true;
    while (idx > 0) {
        idx -= 2;
        const u32 e = st.tab[(u32)job.sym[idx] * 256 + job.sym[idx + 1]];
        const u32 f = e & 0xffffu, base = e >> 16;
        u32 quot = state / f;
        if (quot > 0x1ffffu) {
            if (p > o) out[--p] = (uint16_t)state; else ok = false;
            state >>= 16; quot = state / f;
        }
        state = (state - quot * f) + (quot << Q_NORM_BITS) + base;
    }
    if (p >= o + 2) { out[--p] = (uint16_t)state; out[--p] = (uint16_t)(state >> 16); } else ok = false;
    head_words = o; tail_words = cap - p;
    return ok;
}

/* QNBLIC decode.  R: QNBLIC.c:493-555.  Symbol search replaces the reference's 12 x 32 KiB lookup
 * table by a binary search over the cumulative frequencies (same result for every slot). */
__device__ void qnblic_decode_stream(const QJob &job, const QState &st) {
    const int h = job.h, w = job.w;
    const uint16_t *in = job.stream;
    const u32 avail = job.stream_cap_words;
    u32 rd = 4;
#define Q_NEXT() (rd < avail ? (u32)in[rd++] : (rd++, 0u))
    for (int k = 0; k < Q_CTX_ENTRIES; k++) st.ctx[k] = 0;
    for (int c = 0; c < Q_CLASSES; c++) { /* R: QNBLIC.c:415-459 */
        u32 *hist = st.tab + c * 256;
        for (int k = 0; k < 256; k++) hist[k] = 0;
        u32 pos = 0, sum = 0;
#define Q_PUSH(val) do { const u32 v_ = (val); if (pos < 256) { hist[pos] = v_; sum += v_; } pos++; } while (0)
        while (pos < 256 && sum < Q_NORM_SUM) {
            const u32 code = Q_NEXT();
            if ((code >> 15) == 0) { Q_PUSH(code); }
            else if ((code >> 14) == 2) { Q_PUSH((code >> 7) & 0x7f); Q_PUSH(code & 0x7f); }
            else if ((code >> 12) == 12) { Q_PUSH((code >> 8) & 15); Q_PUSH((code >> 4) & 15); Q_PUSH(code & 15); }
            else if ((code >> 12) == 13) { Q_PUSH((code >> 9) & 7); Q_PUSH((code >> 6) & 7); Q_PUSH((code >> 3) & 7); Q_PUSH(code & 7); }
            else {
                u32 run = (code & 0xff) + 4;
                const u32 closer = (code >> 8) & 15, bit = (code >> 12) & 1;
                while (run--) Q_PUSH(bit);
                if (closer != bit) Q_PUSH(closer);
            }
        }
#undef Q_PUSH
        q_pack_cumulative(hist);
    }
    u32 state = Q_NEXT() << 16; state |= Q_NEXT(); /* R: QNBLIC.c:256-260 */

    uint8_t *img = job.rec;
    for (int i = 0; i < h; i++) { /* R: QNBLIC.c:520-552 */
        Nb nb;
        int err = 0;
        sample_positional(img, w, i, 0, nb);
        for (int j = 0; j < w; j++) {
            const Pred pt = predictor_terms(nb);
            const int px0 = blend_prediction(pt, q_weight(pt.spread));
            const int cls = q_class(activity(nb, err));
            const int adr = q_ctx_address(nb, px0, cls);
            const int ctx = st.ctx[adr];
            int px, sign;
            q_bias_apply(ctx, px0, px, sign);
            const u32 slot = state & (Q_NORM_SUM - 1);
            const u32 *tab = st.tab + cls * 256;
            u32 y = 0; /* last y with cumulative[y] <= slot */
#pragma unroll
            for (u32 step = 128; step > 0; step >>= 1) { if ((tab[y + step] >> 16) <= slot) y += step; }
            const u32 e = tab[y];
            state = (state >> Q_NORM_BITS) * (e & 0xffffu) + slot - (e >> 16);
            if (state < (1u << 16)) state = (state << 16) | Q_NEXT();
            const int x = q_unfold((int)y, px, sign);
            img[(size_t)i * w + j] = (uint8_t)x;
            err = x - px0;
            st.ctx[adr] = q_bias_learn(ctx, err);
            q_window_shift(img, w, i, j, x, nb);
        }
    }
#undef Q_NEXT
}

} /* namespace nblic */
