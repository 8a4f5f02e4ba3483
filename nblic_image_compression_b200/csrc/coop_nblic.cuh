/*
 * coop_nblic.cuh -- warp-cooperative NBLIC effort-1 coder: one coder stream per warp, the 32 lanes
 * share the work of every pixel.
 *
 *   phase P (lanes = 32 consecutive pixels)   everything that depends only on already-known pixels.
 *       lossless encode: the whole front end -- neighbourhood sampling, the 7-direction predictor,
 *       activity, soft context class, texture bits                                  R: NBLIC.c:287-410
 *       decode / near-lossless encode (reconstruction feedback): the neighbours of the two rows above
 *       and the partial sums of every predictor cost that do not involve the pixel to the left.
 *   phase S (one pixel at a time, warp-uniform registers)
 *       [feedback modes] finish the predictor with a = x[j-1], e = x[j-2]            R: NBLIC.c:307-370
 *       bias cancel + residual fold / unfold   ctx table in shared memory            R: NBLIC.c:413-466
 *       rank mapper                            lane s holds entry s of the key's table R: NBLIC.c:470-523
 *       binarisation, encoder                  lane d owns decision d of the pixel: node address, both
 *                                              counter pairs, mixed probability, counter update
 *       binarisation, decoder                  lanes evaluate every node the unary run / the suffix
 *                                              sub-tree can reach, the coder then walks them
 *                                                                                     R: NBLIC.c:589-679
 *       range coder                            consumes / produces bits through __shfl; bytes move
 *                                              through 128-byte lines held one word per lane
 *                                                                                     R: NBLIC.c:527-586
 *
 * The lossless encoder goes further: its bias table, rank mappers and counter forest advance for 32
 * pixels at once in __match_any rounds (only entries that collide are ordered), see
 * coop_e1_encode_lossless.  Pixels whose Golomb code escapes to the next order (0.1-0.5 %) run their
 * decisions one by one, warp-uniform.
 */
#pragma once
#include "codec_core.cuh"
#include "coop_avp.cuh"
#include "row_stage.cuh"

namespace nblic {

constexpr unsigned FULL = 0xffffffffu;

/* phase-P record of one pixel in the feedback modes (8 words, read back as two 16-byte loads) */
struct __align__(16) PixRec {
    u32 bcdf;   /* b | c << 8 | d << 16 | f << 24                                             */
    u32 ghqr;   /* g | h << 8 | q << 16 | r << 24                                             */
    u32 st_act; /* s | t << 8 | act << 16      act = |b-c| + |b-d| + |b-f| + |d-g|            */
    u32 k01;    /* K0 | K1 << 16               Kn = the three a-free terms of directional cost n */
    u32 k23, k45;
    u32 k6_lin; /* K6 | (lin + 1024) << 16     lin = 9b + 2d - 2c - f                         */
    u32 orig;   /* encoder: the original pixel                                                */
};

/* The record of a pixel from its ten neighbours of the two rows above (a-free partial sums of the directional costs). */
NB_DEV PixRec pixrec_from(int b, int c, int d, int f, int g, int hh, int q, int r, int s, int t, u32 orig) {
    PixRec pr;
    pr.orig = orig;
    pr.bcdf = (u32)b | ((u32)c << 8) | ((u32)d << 16) | ((u32)f << 24);
    pr.ghqr = (u32)g | ((u32)hh << 8) | ((u32)q << 16) | ((u32)r << 24);
    const int act = abs(b - c) + abs(b - d) + abs(b - f) + abs(d - g);
    pr.st_act = (u32)s | ((u32)t << 8) | ((u32)act << 16);
    const int K0 = abs(c - q) + abs(b - c) + abs(d - b);
    const int K1 = abs(c - hh) + abs(b - f) + abs(d - g);
    const int K2 = abs(c - s) + abs(b - hh) + abs(d - f);
    const int K3 = abs(c - f) + abs(b - g) + abs(d - r);
    const int K4 = abs(2 * c - q - s) + abs(2 * b - c - hh) + abs(2 * d - b - f);
    const int K5 = abs(2 * c - s - hh) + abs(2 * b - hh - f) + abs(2 * d - f - g);
    const int K6 = abs(2 * c - hh - f) + abs(2 * b - f - g) + abs(2 * d - g - r);
    pr.k01 = (u32)K0 | ((u32)K1 << 16); pr.k23 = (u32)K2 | ((u32)K3 << 16); pr.k45 = (u32)K4 | ((u32)K5 << 16);
    pr.k6_lin = (u32)K6 | ((u32)(9 * b + 2 * d - 2 * c - f + 1024) << 16);
    return pr;
}
/* Phase P of the feedback modes for pixel (i, j), i >= 1: neighbours of the two rows above with the
 * reference's border fallbacks (R: NBLIC.c:288-303), read from global memory (row 1, where f..s fall back per column). */
NB_DEV PixRec make_pixrec(const uint8_t *row, int w, int i, int j, u32 orig) {
    const uint8_t *r1 = row - w, *r2 = r1 - w;
    const bool up2 = i >= 2, l1 = j >= 1, l2 = j >= 2, rt1 = j + 1 < w, rt2 = j + 2 < w;
    const int b = r1[j];
    const int c = l1 ? (int)r1[j - 1] : b;
    const int d = rt1 ? (int)r1[j + 1] : b;
    const int f = up2 ? (int)r2[j] : b;
    const int g = (up2 && rt1) ? (int)r2[j + 1] : f;
    const int hh = (up2 && l1) ? (int)r2[j - 1] : f;
    const int q = l2 ? (int)r1[j - 2] : c;
    const int r = (up2 && rt2) ? (int)r2[j + 2] : g;
    const int s = (up2 && l2) ? (int)r2[j - 2] : hh;
    const int t = rt2 ? (int)r1[j + 2] : d;
    return pixrec_from(b, c, d, f, g, hh, q, r, s, t, orig);
}
/* The same from rows staged in shared memory with materialised halo cells (row_stage.cuh), i >= 2: no predicates.
 * p1 / p2 -> pixel j of rows i-1 / i-2. */
NB_DEV PixRec make_pixrec_staged(const uint8_t *p1, const uint8_t *p2, u32 orig) {
    return pixrec_from(p1[0], p1[-1], p1[1], p2[0], p2[1], p2[-1], p1[-2], p2[2], p2[-2], p1[2], orig);
}
/* Finish the 7-direction predictor from a record (words ra, rb) once a and e are known.  nb must already
 * hold a, b, c, d, e, q.  R: NBLIC.c:307-364 / QNBLIC.c:94-143 */
NB_DEV Pred finish_predictor(const Nb &nb, const uint4 &ra, const uint4 &rb) {
    const int a = nb.a, e = nb.e;
    const int c0 = 2 * (abs(a - e) + (int)(ra.w & 0xffffu)), c1 = 2 * (abs(a - nb.c) + (int)(ra.w >> 16));
    const int c2 = 2 * (abs(a - nb.q) + (int)(rb.x & 0xffffu)), c3 = 2 * (abs(a - nb.b) + (int)(rb.x >> 16));
    const int c4 = abs(2 * a - e - nb.q) + (int)(rb.y & 0xffffu), c5 = abs(2 * a - nb.q - nb.c) + (int)(rb.y >> 16);
    const int c6 = abs(2 * a - nb.c - nb.b) + (int)(rb.z & 0xffffu);
    Pred pt;
    int best = c0;
    pt.ang2 = 2 * a;
    if (c1 < best) { best = c1; pt.ang2 = 2 * nb.b; }
    if (c2 < best) { best = c2; pt.ang2 = 2 * nb.c; }
    if (c3 < best) { best = c3; pt.ang2 = 2 * nb.d; }
    if (c4 < best) { best = c4; pt.ang2 = a + nb.c; }
    if (c5 < best) { best = c5; pt.ang2 = nb.c + nb.b; }
    if (c6 < best) { best = c6; pt.ang2 = nb.b + nb.d; }
    pt.spread = c0 + c1 + c2 + c3 + c4 + c5 + c6 - 7 * best;
    pt.lin16 = clampi(9 * a + (int)(rb.z >> 16) - 1024 - e, 0, 16 * 255);
    return pt;
}

/* shared-memory image of one stream's adaptive state (one warp = one CTA).  The counter forest follows
 * as a separate, compacted array: class u only owns the (256 >> top) << (u / k_step) nodes its Golomb
 * order can reach (1000 nodes = 4 KB for lossless streams instead of 16 x 256). */
struct CoopSmem {
    int16_t ctx[N_CTX_ENTRIES];        /*  4 KB: bias-cancel table                                    */
    uint16_t soft[208];                /* activity (clamped to 200) -> u | v << 4 | wv << 8             */
    uint16_t fbase[N_CLASSES];         /* first forest slot of class u                                 */
};
/* Dynamic shared memory of a coop_nblic_kernel CTA, in this order:
 *   CoopSmem | PixRec[32] (feedback modes) | staged row tiles (row_stage.cuh) | AvpSmem (efforts 2/3) | rank[512*20] bytes (effort 1) | forest[]
 * The rank tables (10 KB: encoder symbol -> rank, decoder rank -> symbol) stay in shared memory for
 * effort 1 while the batch fits the 11 streams/SM that allows; efforts 2/3 spend ~40k cycles per pixel in
 * the least-squares solve, so there (and for effort-1 batches larger than 11 x SMs images, RG = true) the
 * tables move to global memory (L2) and the freed space buys 16 (efforts 2/3) or 24 resident streams. */
template <int NAVP, int MODE, bool RG> struct CoopLayout {
    static constexpr bool kFeedback = MODE != 0;
    static constexpr bool kRankGlobal = RG;
    static constexpr int kStageRows = kFeedback ? 2 : 3; /* row_stage.cuh: the rows above (+ the current row when all pixels are known) */
    static constexpr size_t kRecOff = (sizeof(CoopSmem) + 15) & ~(size_t)15;
    static constexpr size_t kStageOff = kRecOff + (kFeedback ? sizeof(PixRec) * 32 : 0);
    static constexpr size_t kAvpOff = kStageOff + 2 * kStageRows * kStageLine;
    static constexpr size_t kRankOff = kAvpOff + (NAVP > 0 ? sizeof(AvpSmem) : 0);
    static constexpr size_t kForestOff = (kRankOff + (kRankGlobal ? 0 : N_RANK_ENTRIES) + 15) & ~(size_t)15;
};

/* number of forest nodes a stream with this k_step can touch (host and device) */
__host__ __device__ inline int forest_nodes(int k_step) {
    const int top = (N_CLASSES - 1) / k_step;
    int n = 0;
    for (int u = 0; u < N_CLASSES; u++) n += (256 >> top) << (u / k_step);
    return n;
}
/* slot of tree node `node` inside a class of order k: unary index and in-order suffix offset */
NB_DEV int compact_node(int node, int top, int k) { return ((node >> top) << k) + (node & ((1 << k) - 1)); }

/* ---- byte streams: a 128-byte line lives in the warp, one 32-bit word per lane ---------------- */

struct LineWriter { /* base must be 128-byte aligned and private to the stream */
    uint8_t *base;
    u32 cap, pos, word;
    int lane;
    bool overflow;
    NB_DEV void start(uint8_t *p, u32 capacity, int ln) { base = p; cap = capacity; pos = 0; word = 0; lane = ln; overflow = false; }
    NB_DEV void flush_line(u32 line_off) {
        const u32 at = line_off + 4u * (u32)lane;
        if (at + 4u <= cap) *reinterpret_cast<u32 *>(base + at) = word;
        word = 0;
    }
    NB_DEV void put(u32 byte) {
        if (pos >= cap) overflow = true;
        if ((int)((pos >> 2) & 31u) == lane) word |= (byte & 255u) << (8u * (pos & 3u));
        pos++;
        if ((pos & 127u) == 0) flush_line(pos - 128u);
    }
    NB_DEV void finish() { if (pos & 127u) flush_line(pos & ~127u); }
};

struct LineReader {
    const uint8_t *base;
    u32 len, pos, word;
    unsigned long long line; /* absolute address of the cached line, 0 = none */
    int lane;
    NB_DEV void start(const uint8_t *p, u32 n, u32 at, int ln) { base = p; len = n; pos = at; word = 0; line = 0; lane = ln; }
    NB_DEV u32 get() {
        if (pos >= len) { pos++; return 0u; } /* the reference reads on; we return zeros */
        const unsigned long long a = (unsigned long long)(base + pos);
        if ((a & ~127ull) != line) { /* a 128-byte aligned line that holds a valid byte is always mapped */
            line = a & ~127ull;
            word = *reinterpret_cast<const u32 *>(line + 4ull * (unsigned)lane);
        }
        const u32 wsel = __shfl_sync(FULL, word, (int)((a >> 2) & 31ull));
        pos++;
        return (wsel >> (8u * (u32)(a & 3ull))) & 255u;
    }
};

/* (span >> 12) * p + (((span & 0xfff) * p) >> 12) of the reference (R: NBLIC.c:556) equals floor(span * p / 4096):
 * span * p = (span >> 12) * 4096 * p + (span & 0xfff) * p, and the first term is a multiple of 4096.  One wide
 * multiply and a funnel shift instead of seven instructions. */
NB_DEV u32 split_point(u32 span, u32 p1) { return (u32)(((u64)span * p1) >> 12); }

/* range coder with warp-uniform registers */
template <bool DEC> struct CoopCoder;

template <> struct CoopCoder<false> {
    u32 lo, hi;
    LineWriter out;
    NB_DEV void start() { lo = 0; hi = 0xffffffffu; }
    NB_DEV int bit(int b, u32 p1) {
        const u32 mid = lo + split_point(hi - lo, p1);
        if (b) hi = mid; else lo = mid + 1;
        while (((lo ^ hi) & 0xff000000u) == 0) { out.put(hi >> 24); lo <<= 8; hi = (hi << 8) | 0xffu; }
        return b;
    }
    NB_DEV void finish() { for (int k = 0; k < 4; k++) { out.put(lo >> 24); lo <<= 8; } out.finish(); }
};

template <> struct CoopCoder<true> {
    u32 lo, hi, code;
    LineReader in;
    NB_DEV void start() { lo = 0; hi = 0xffffffffu; code = 0; for (int k = 0; k < 4; k++) code = (code << 8) | in.get(); }
    NB_DEV int bit(int, u32 p1) {
        const u32 mid = lo + split_point(hi - lo, p1);
        const int b = code <= mid;
        if (b) hi = mid; else lo = mid + 1;
        while (((lo ^ hi) & 0xff000000u) == 0) { code = (code << 8) | in.get(); lo <<= 8; hi = (hi << 8) | 0xffu; }
        return b;
    }
};

/* node_learn on the packed pair n0 | n1 << 16 whose sum s = n0 + n1 the caller already knows (R: NBLIC.c:606-617) */
NB_DEV u32 learn_packed(u32 c, u32 s, int bit, int weight) {
    c += (u32)weight << (bit ? 16 : 0);
    if (s + (u32)weight > (u32)(N_MIX * 256)) c = ((c + 0x00010001u) >> 1) & 0x7fff7fffu;
    return c;
}
NB_DEV u32 pair_sum(u32 c) { return (c & 0xffffu) + (c >> 16); }
/* floor(4096 * n1 / s) without the integer divide and without the conversion pipe: n1 and s (< 2^23) become
 * floats by the 2^23 mantissa trick, the quotient estimate (error < 2e-3) is rounded to an integer by the
 * 1.5 * 2^23 trick, then multiply back and correct by one.  Only the reciprocal itself uses the SFU. */
NB_DEV int node_p1_fast(u32 packed, u32 s) {
    const u32 n1 = packed >> 16;
    const float fn1 = __uint_as_float(0x4B000000u | n1) - 8388608.0f;
    const float fs = __uint_as_float(0x4B000000u | s) - 8388608.0f;
    float rcp;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rcp) : "f"(fs));
    int q = (int)(__float_as_uint(fn1 * 4096.0f * rcp + 12582912.0f) - 0x4B400000u); /* nearest integer to the estimate */
    const int r = (int)(n1 << 12) - q * (int)s;
    if (r < 0) q--; else if (r >= (int)s) q++;
    return q;
}
/* R: NBLIC.c:627-631.  Both p1 are <= 4095 (n0 >= 1), so only the lower clip can act. */
NB_DEV u32 mixed_p(u32 cu, u32 cv, u32 su, u32 sv, int wv) {
    const int p = (node_p1_fast(cu, su) * (N_MIX - wv) + node_p1_fast(cv, sv) * wv + N_MIX / 2) >> 5; /* operands >= 0: >> 5 == / 32 */
    return (u32)max(p, 1);
}
/* iu / iv: forest slots of the node in the main and the side class (equal when both are the same class) */
NB_DEV void learn_pair(u32 *forest, int iu, int iv, u32 cu, u32 cv, u32 su, u32 sv, int wv, int bit) {
    if (iu == iv) { const u32 c1 = learn_packed(cu, su, bit, N_MIX - wv); forest[iu] = learn_packed(c1, pair_sum(c1), bit, wv); }
    else { forest[iu] = learn_packed(cu, su, bit, N_MIX - wv); forest[iv] = learn_packed(cv, sv, bit, wv); }
}

/* k = u / k_step for u = 0..15 as a multiply-shift: (u * (65536 / k_step + 1)) >> 16 is exact for 3 <= k_step <= 16 */
NB_DEV u32 make_order_table(int k_step) { return 65536u / (u32)k_step + 1u; }
NB_DEV int order_of(u32 magic, int u) { return (int)(((u32)u * magic) >> 16); }

NB_DEV void coop_reset(CoopSmem &sm, uint8_t *rank, u32 *forest, int k_step, int *count, int lane) {
    const int n_nodes = forest_nodes(k_step), top = (N_CLASSES - 1) / k_step;
    for (int k = lane; k < n_nodes; k += 32) forest[k] = (u32)N_MIX | ((u32)N_MIX << 16);
    if (lane < N_CLASSES) { int b = 0; for (int u = 0; u < lane; u++) b += (256 >> top) << (u / k_step); sm.fbase[lane] = (uint16_t)b; }
    for (int k = lane; k < N_CTX_ENTRIES; k += 32) sm.ctx[k] = 0;
    for (int k = lane; k < N_RANK_ENTRIES; k += 32) { const int r = k % N_RANKS; rank[k] = (uint8_t)r; count[k] = 2 * (N_RANKS - 1 - r); }
    for (int d = lane; d <= 200; d += 32) { int u, v, wv; n_soft_class(d, u, v, wv); sm.soft[d] = (uint16_t)(u | (v << 4) | (wv << 8)); }
    __syncwarp();
}

NB_DEV void coop_put_header(CoopCoder<false> &rc, int h, int w, int near, int k_step, int effort) { /* R: NBLIC.c:682-694 */
    const char magic[8] = {'N', 'B', 'L', 'I', 'C', '0', '.', '3'};
    for (int k = 0; k < 8; k++) rc.out.put((u32)magic[k]);
    rc.out.put(1); rc.out.put((u32)h >> 8); rc.out.put((u32)h); rc.out.put((u32)w >> 8); rc.out.put((u32)w);
    rc.out.put((u32)near); rc.out.put((u32)k_step); rc.out.put((u32)effort);
}

/* ---- encoder side of one symbol: every decision of the pixel in its own lane ------------------- */
NB_DEV void coop_encode_symbol(CoopCoder<false> &rc, CoopSmem &sm, u32 *forest, int k_step, int top, u32 ktab, int u, int v, int wv,
                               int z, int lane) {
    const int k = order_of(ktab, u);
    if (order_of(ktab, v) != k) v = u;
    const int bu = sm.fbase[u], bv = sm.fbase[v];
    const int q = z >> k;
    const int D = q + 1 + k; /* decisions of this pixel */
    if (q < (256 >> top) && D <= 32) {
        int slot, bit; /* slot = compacted node index inside the class */
        if (lane <= q) { slot = lane << k; bit = lane < q; }             /* unary run, then its terminating 0 */
        else {
            const int t = lane - q - 1, kk = max(k - 1 - t, 0);          /* t-th suffix bit, weight 2^kk */
            const int hi_bits = (z & ((1 << k) - 1)) & ~((2 << kk) - 1);
            slot = (q << k) + ((1 + hi_bits + t - __popc(hi_bits)) & ((1 << k) - 1));
            bit = (z >> kk) & 1;
        }
        u32 coded = 0;
        if (lane < D) {
            const u32 cu = forest[bu + slot], cv = forest[bv + slot], su = pair_sum(cu), sv = pair_sum(cv);
            coded = mixed_p(cu, cv, su, sv, wv) | ((u32)bit << 12);
            learn_pair(forest, bu + slot, bv + slot, cu, cv, su, sv, wv, bit);
        }
        for (int d = 0; d < D; d++) {
            const u32 cd = __shfl_sync(FULL, coded, d);
            rc.bit((int)(cd >> 12), cd & 0xfffu);
        }
    } else { /* order escape (or an over-long unary run): the decisions one after the other, warp-uniform */
        int node = 0, kk = k, iu_base = bu, iv_base = bv, phase = 0; /* phase 0: unary, 1: suffix, 2: done */
        int k_tree = k;                                             /* order of the tree being walked */
        while (phase < 2) {
            int bit;
            if (phase == 0) bit = (node >> top) < (z >> kk);
            else bit = (z >> kk) & 1;
            const int slot = compact_node(node, top, k_tree);
            const u32 cu = forest[iu_base + slot], cv = forest[iv_base + slot], su = pair_sum(cu), sv = pair_sum(cv);
            rc.bit(bit, mixed_p(cu, cv, su, sv, wv));
            __syncwarp();
            if (lane == 0) learn_pair(forest, iu_base + slot, iv_base + slot, cu, cv, su, sv, wv, bit);
            __syncwarp();
            if (phase == 0) {
                if (!bit) { node++; kk--; phase = kk >= 0 ? 1 : 2; }
                else {
                    node += 1 << top;
                    if (node >= 256) { /* escape to the next order (R: NBLIC.c:658-662) */
                        node >>= 1;
                        const int uu = (kk + 1) * k_step;
                        if (uu >= N_CLASSES) { rc.out.overflow = true; phase = 2; }
                        else { kk = order_of(ktab, uu); k_tree = kk; iu_base = iv_base = sm.fbase[uu]; }
                    }
                }
            } else {
                node += bit ? (1 << kk) : 1;
                kk--;
                if (kk < 0) phase = 2;
            }
        }
    }
}

/* ---- decoder side of one symbol ----------------------------------------------------------------- */
/* Returns z, or -1 for a corrupt stream.
 * A class's compacted nodes are laid out so that slot = (unary index << k) + in-order suffix offset; the
 * walk of one symbol therefore stays inside the aligned 2^k-slot group of its unary index.  The lanes
 * evaluate a 32-slot window (slot = window + lane: one coalesced load per class), the coder walks the
 * unary run and the suffix inside it, and the visited lanes then update their counters.  For k <= 2 the whole
 * class is one window; for k = 3 a second window is needed only when the unary run exceeds 3. */
NB_DEV int coop_decode_symbol(CoopCoder<true> &rc, CoopSmem &sm, u32 *forest, int k_step, int top, u32 ktab, int u, int v, int wv, int lane) {
    int k = order_of(ktab, u);
    if (order_of(ktab, v) != k) v = u;
    const int bu = sm.fbase[u], bv = sm.fbase[v];
    const int n_unary = 256 >> top;  /* unary nodes of one tree before the order escape */
    const int n_slots = n_unary << k;
    int q = 0, z = 0;
    bool done = false;
    while (!done && q < n_unary) {
        const int w0 = (q << k) & ~31;
        const int slot = w0 + lane;
        u32 cu = 0, cv = 0, su = 64, sv = 64, p = 0;
        if (slot < n_slots) { cu = forest[bu + slot]; cv = forest[bv + slot]; su = pair_sum(cu); sv = pair_sum(cv); p = mixed_p(cu, cv, su, sv, wv); }
        /* unary run inside the window: nodes q << k */
        const int q_first = q;
        bool zero_seen = false;
        while (q < n_unary && ((q << k) & ~31) == w0) {
            if (!rc.bit(0, __shfl_sync(FULL, p, (q << k) & 31))) { zero_seen = true; break; }
            q++;
        }
        /* lane-side view of the run: my slot is unary node uq iff its low k bits are zero */
        const int uq = slot >> k;
        bool visited = (slot & ((1 << k) - 1)) == 0 && uq >= q_first && (zero_seen ? uq <= q : uq < q);
        int my_bit = uq < q;
        if (zero_seen) {
            z = q << k;
            if (k > 0) { /* suffix: stays inside the aligned 2^k-slot group of q, hence inside this window */
                int m = 1;
                unsigned path = 0, ones = 0;
                for (int kk = k - 1; kk >= 0; kk--) {
                    const int ln = ((q << k) + m) & 31;
                    const int b = rc.bit(0, __shfl_sync(FULL, p, ln));
                    path |= 1u << ln; ones |= (u32)b << ln;
                    z += b << kk;
                    m += b ? (1 << kk) : 1;
                }
                if ((path >> lane) & 1u) { visited = true; my_bit = (int)((ones >> lane) & 1u); }
            }
            done = true;
        }
        if (visited) learn_pair(forest, bu + slot, bv + slot, cu, cv, su, sv, wv, my_bit);
        __syncwarp();
    }
    if (done) return z;
    /* order escape: continue one decision at a time in the next order's tree (R: NBLIC.c:658-662) */
    int uu = (k + 1) * k_step, node = 128;
    for (;;) {
        if (uu >= N_CLASSES) return -1;
        k = order_of(ktab, uu);
        const int at = sm.fbase[uu] + compact_node(node, top, k);
        const u32 c = forest[at], sc = pair_sum(c);
        const int bit = rc.bit(0, mixed_p(c, c, sc, sc, wv));
        __syncwarp();
        if (lane == 0) learn_pair(forest, at, at, c, c, sc, sc, wv, bit);
        __syncwarp();
        if (!bit) break;
        node += 1 << top;
        if (node >= 256) { node >>= 1; uu = (k + 1) * k_step; }
    }
    z = (node >> top) << k;
    const int k_tree = k;
    for (node++, k--; k >= 0; k--) {
        const int at = sm.fbase[uu] + compact_node(node & 255, top, k_tree);
        const u32 c = forest[at], sc = pair_sum(c);
        const int bit = rc.bit(0, mixed_p(c, c, sc, sc, wv));
        __syncwarp();
        if (lane == 0) learn_pair(forest, at, at, c, c, sc, sc, wv, bit);
        __syncwarp();
        if (bit) z += 1 << k;
        node += bit ? (1 << k) : 1;
    }
    return z;
}

/* ---- rank mapper, lane s holds entry s of the key's table ---------------------------------------- */
/* Both directions first fetch the key's 20 table entries and frequencies (lane s < 20 holds entry s);
 * the frequencies live in global memory (L2), so the fetch is issued before the symbol is coded. */
template <bool RG>
NB_DEV void coop_rank_fetch(const uint8_t *rank, const int *count, int key, int lane, int &my_tab, int &my_count) {
    my_tab = lane < N_RANKS ? (int)(RG ? __ldcg(rank + key + lane) : rank[key + lane]) : 255;
    my_count = lane < N_RANKS ? __ldcg(count + key + lane) : 0;
}
/* encoder: table = symbol -> rank */
NB_DEV int coop_rank_encode(int y, int my_rank) {
    const int zr = __shfl_sync(FULL, my_rank, y & 31);
    return y < N_RANKS ? zr : y;
}
NB_DEV void coop_rank_touch_encode(uint8_t *rank, int *count, int key, int y, int z, int lane, int my_rank, int my_count) {
    if (y >= N_RANKS) return;
    const int cz = __shfl_sync(FULL, my_count, z) + 1;
    const int cp = __shfl_sync(FULL, my_count, max(z - 1, 0));
    const unsigned holders = __ballot_sync(FULL, my_rank == z - 1);
    if (lane == 0) {
        if (z > 0 && cp < cz) { /* one adjacent promotion (R: NBLIC.c:513-521) */
            const int other = __ffs(holders) - 1;
            count[key + z] = cp; count[key + z - 1] = cz;
            rank[key + y] = (uint8_t)(z - 1); rank[key + other] = (uint8_t)z;
        } else count[key + z] = cz;
    }
}
/* decoder: table = rank -> symbol.  Returns y and performs the update. */
NB_DEV int coop_rank_decode(uint8_t *rank, int *count, int key, int z, int lane, int my_sym, int my_count) {
    if (z >= N_RANKS) return z;
    const int y = __shfl_sync(FULL, my_sym, z);
    const int other = __shfl_sync(FULL, my_sym, max(z - 1, 0));
    const int cz = __shfl_sync(FULL, my_count, z) + 1;
    const int cp = __shfl_sync(FULL, my_count, max(z - 1, 0));
    if (lane == 0) {
        if (z > 0 && cp < cz) {
            count[key + z] = cp; count[key + z - 1] = cz;
            rank[key + z] = (uint8_t)other; rank[key + z - 1] = (uint8_t)y;
        } else count[key + z] = cz;
    }
    return y;
}

/*
 * Lossless effort-1 encode of one image by one warp (phase P covers the whole front end).
 * `count`: the stream's rank-mapper frequency table in global memory ([512][20] int, indexed by rank).
 * `stream` must be 128-byte aligned.  Returns the stream length or 0xffffffff on overflow.
 */
template <bool RG>
__device__ u32 coop_e1_encode_lossless(const uint8_t *img, int h, int w, uint8_t *stream, u32 cap, CoopSmem &sm, uint8_t *stage_buf, uint8_t *rank,
                                        u32 *forest, int *count, int lane) {
    const int k_step = 3, top = (N_CLASSES - 1) / k_step; /* near = 0 (R: NBLIC.c:769) */
    const u32 ktab = make_order_table(k_step);
    coop_reset(sm, rank, forest, k_step, count, lane);
    CoopCoder<false> rc;
    rc.out.start(stream, cap, lane);
    coop_put_header(rc, h, w, 0, k_step, 1);
    rc.start();
    RowStage<3> rows;
    rows.start(stage_buf, img, h, w);

    for (int i = 0; i < h; i++) {
        int carry_px0 = 0; /* px0 of the pixel left of this block */
        for (int j0 = 0; j0 < w; j0 += 32) {
            /* ---------------- phase P: lane = pixel j0 + lane ---------------- */
            const bool active = j0 + lane < w;
            const int j = min(j0 + lane, w - 1);
            Nb nb;
            int x;
            if (i >= 2) { /* rows i, i-1, i-2 come from shared memory, staged by cp.async one 64-pixel tile ahead */
                if ((j0 & (kStageTile - 1)) == 0) {
                    const bool more = j0 + kStageTile < w;
                    rows.advance(i, j0, more ? i : (i + 1 < h ? i + 1 : -1), more ? j0 + kStageTile : 0, lane, 32, lane == 0);
                }
                sample_staged3(rows, j - (j0 & ~(kStageTile - 1)), nb, x);
                if (j == 1) nb.e = nb.a;
            } else {
                sample_positional(img, w, i, j, nb);
                x = img[(size_t)i * w + j];
            }
            const Pred pt = predictor_terms(nb);
            const int px0 = blend_prediction(pt, n_weight(pt.spread));
            int px0_left = __shfl_up_sync(FULL, px0, 1);
            if (lane == 0) px0_left = carry_px0;
            const int err_in = j == 0 ? 0 : clampi(nb.a - px0_left, -127, 127); /* nb.a is the coded value of pixel j-1 */
            const u32 soft = sm.soft[min(activity(nb, err_in), 200)];
            const int adr = active ? (int)((((soft & 15u) >> 1) << 8) | (u32)texture_bits(nb, px0)) : 0x10000 + lane;
            carry_px0 = __shfl_sync(FULL, px0, 31);

            /* ---- bias table (R: NBLIC.c:413-428) and residual fold (:431-447): pixels that share a table
             * address go in raster order, one per round; different addresses proceed together ---- */
            int px = 0, sign = 0, y = 0;
            {
                const unsigned peers = __match_any_sync(FULL, adr);
                const int my_turn = __popc(peers & ((1u << lane) - 1u));
                const int rounds = __reduce_max_sync(FULL, active ? __popc(peers) : 0);
                for (int r = 0; r < rounds; r++) {
                    if (active && my_turn == r) {
                        const int c = sm.ctx[adr];
                        n_bias_apply(c, px0, px, sign);
                        const int room = min(px, 255 - px), mag = abs(x - px);
                        y = mag == 0 ? 0 : (mag <= room ? 2 * mag - ((x >= px) ^ sign) : mag + room);
                        sm.ctx[adr] = (int16_t)n_bias_learn(c, clampi(x - px0, -127, 127));
                    }
                    __syncwarp();
                }
            }
            /* ---- rank mapper (R: NBLIC.c:470-523), same scheme keyed by (px, sign) ---- */
            int z = y;
            {
                const int key = ((px << 1) | sign) * N_RANKS;
                const bool ranked = active && y < N_RANKS;
                const unsigned peers = __match_any_sync(FULL, ranked ? key : -1 - lane);
                const int my_turn = __popc(peers & ((1u << lane) - 1u));
                const int rounds = __reduce_max_sync(FULL, ranked ? __popc(peers) : 0);
                for (int r = 0; r < rounds; r++) {
                    if (ranked && my_turn == r) {
                        z = RG ? (int)__ldcg(rank + key + y) : (int)rank[key + y];
                        const int cz = __ldcg(count + key + z) + 1;
                        bool promoted = false;
                        if (z > 0) {
                            const int cp = __ldcg(count + key + z - 1);
                            if (cp < cz) { /* one adjacent promotion: find the symbol that holds rank z-1 */
                                const u32 want = (u32)(z - 1) * 0x01010101u;
                                int other = 0;
#pragma unroll
                                for (int q4 = 0; q4 < N_RANKS / 4; q4++) {
                                    const u32 *wp = reinterpret_cast<const u32 *>(rank + key) + q4;
                                    const u32 hit = __vcmpeq4(RG ? __ldcg(wp) : *wp, want);
                                    if (hit) other = 4 * q4 + ((__ffs((int)hit) - 1) >> 3);
                                }
                                count[key + z] = cp; count[key + z - 1] = cz;
                                rank[key + y] = (uint8_t)(z - 1); rank[key + other] = (uint8_t)z;
                                promoted = true;
                            }
                        }
                        if (!promoted) count[key + z] = cz;
                    }
                    __syncwarp();
                }
            }
            /* ---- binarisation (R: NBLIC.c:589-679).  Every decision of the block visits two counter nodes; a
             * node's visits must happen in decision order, visits to different nodes are independent.  The
             * decisions are flattened (prefix sum of the per-pixel counts), taken 16 at a time with lane
             * 2*dd + role = visit of decision dd to its main (role 0) / side (role 1) class node, and the
             * visits advance in match_any rounds like the tables above.  The range coder then walks the 16
             * (bit, p) pairs.  A block containing an order escape falls back to the per-pixel routine. ---- */
            const int cu_ = (int)(soft & 15u), wv_ = (int)((soft >> 8) & 31u), k_ = order_of(ktab, cu_);
            const int cv_ = order_of(ktab, (int)((soft >> 4) & 15u)) == k_ ? (int)((soft >> 4) & 15u) : cu_;
            const int q_ = z >> k_;
            const int D_ = active ? q_ + 1 + k_ : 0;
            const u32 rec0 = (u32)cu_ | ((u32)cv_ << 4) | ((u32)wv_ << 8) | ((u32)z << 13); /* u:4 v:4 wv:5 z:8 */
            const int n_here = min(32, w - j0);
            if (__any_sync(FULL, active && q_ >= (256 >> top))) {
                for (int jj = 0; jj < n_here; jj++) {
                    const u32 r0 = __shfl_sync(FULL, rec0, jj);
                    coop_encode_symbol(rc, sm, forest, k_step, top, ktab, (int)(r0 & 15u), (int)((r0 >> 4) & 15u), (int)((r0 >> 8) & 31u), (int)(r0 >> 13), lane);
                    __syncwarp();
                }
                continue;
            }
            int off = D_; /* exclusive prefix sum of the decision counts */
#pragma unroll
            for (int dlt = 1; dlt < 32; dlt <<= 1) { const int t = __shfl_up_sync(FULL, off, dlt); if (lane >= dlt) off += t; }
            const int total = __shfl_sync(FULL, off, 31);
            off -= D_;
            for (int base = 0; base < total; base += 16) {
                const int dd = lane >> 1, role = lane & 1, g = base + dd;
                const bool valid = g < total;
                int owner = 0; /* the pixel decision g belongs to: largest lane whose offset is <= g */
#pragma unroll
                for (int step = 16; step >= 1; step >>= 1) {
                    const int cand = owner + step;
                    const int o = __shfl_sync(FULL, off, cand & 31);
                    if (cand < 32 && o <= g) owner = cand;
                }
                const u32 r0 = __shfl_sync(FULL, rec0, owner);
                const int d = g - __shfl_sync(FULL, off, owner);
                const int u = (int)(r0 & 15u), v = (int)((r0 >> 4) & 15u), wv = (int)((r0 >> 8) & 31u), zz = (int)(r0 >> 13);
                const int k = order_of(ktab, u), q = zz >> k;
                int slot, bit;
                if (d <= q) { slot = d << k; bit = d < q; }
                else {
                    const int t = d - q - 1, kk = max(k - 1 - t, 0);
                    const int hi_bits = (zz & ((1 << k) - 1)) & ~((2 << kk) - 1);
                    slot = (q << k) + ((1 + hi_bits + t - __popc(hi_bits)) & ((1 << k) - 1));
                    bit = (zz >> kk) & 1;
                }
                const bool same = u == v;
                const bool visiting = valid && !(same && role == 1);
                const int idx = visiting ? (int)sm.fbase[role ? v : u] + slot : -1 - lane;
                const int weight = role ? wv : N_MIX - wv;
                const unsigned peers = __match_any_sync(FULL, idx);
                const int my_turn = __popc(peers & ((1u << lane) - 1u));
                const int rounds = __reduce_max_sync(FULL, visiting ? __popc(peers) : 0);
                int p1 = 0;
                for (int r = 0; r < rounds; r++) {
                    if (visiting && my_turn == r) {
                        const u32 c = forest[idx], sc = pair_sum(c);
                        p1 = node_p1_fast(c, sc);
                        u32 c2 = learn_packed(c, sc, bit, weight);
                        if (same) c2 = learn_packed(c2, pair_sum(c2), bit, wv);
                        forest[idx] = c2;
                    }
                    __syncwarp();
                }
                const int p_mate = __shfl_xor_sync(FULL, p1, 1);
                const int pu = role ? p_mate : p1, pv = same ? pu : (role ? p1 : p_mate);
                const u32 coded = (u32)max((pu * (N_MIX - wv) + pv * wv + N_MIX / 2) >> 5, 1) | ((u32)bit << 12);
                const int nd = min(16, total - base);
                for (int e = 0; e < nd; e++) {
                    const u32 cd = __shfl_sync(FULL, coded, 2 * e);
                    rc.bit((int)(cd >> 12), cd & 0xfffu);
                }
            }
        }
    }
    rc.finish();
    return rc.out.overflow ? 0xffffffffu : rc.out.pos;
}

/*
 * Stream with a per-pixel sequential front end: the decoder (any near), the near-lossless encoder, and
 * every effort-2/3 stream (NAVP = 6 / 10: the AVP accumulators form a pixel-to-pixel chain even when
 * the pixels are known).  `nbimg` is the raster the neighbours come from, `out_rec` (may be NULL when
 * nbimg is the source image) receives the decoded / reconstructed pixels.
 * Encoder: `stream` 128-byte aligned, cap = capacity; returns length or 0xffffffff.
 * Decoder: cap = valid bytes; returns 0, or 1 for a corrupt stream.
 */
template <int NAVP, bool DEC, bool RG>
__device__ u32 coop_feedback(const uint8_t *src, const uint8_t *nbimg, uint8_t *out_rec, int h, int w, int near, int k_step, uint8_t *stream,
                             u32 cap, CoopSmem &sm, PixRec *recs, uint8_t *stage_buf, AvpSmem *avp_sm, uint8_t *rank, u32 *forest, i64 *Brow, i64 *Frow,
                             int *count, int lane) {
    constexpr int AN = NAVP > 0 ? NAVP : 1;
    constexpr int AM = AvpGeom<AN>::M, ANS = AvpGeom<AN>::NS;
    const int top = (N_CLASSES - 1) / k_step;
    const u32 ktab = make_order_table(k_step);
    const int qn = 2 * near + 1;
    const u32 qmagic = 65536u / (u32)qn + 1u; /* n / qn == (n * qmagic) >> 16 for 0 <= n < 3400 */
    coop_reset(sm, rank, forest, k_step, count, lane);
    i64 E[ANS], ridge = 8;
    if constexpr (NAVP > 0) {
        for (size_t k = lane; k < (size_t)w * AM; k += 32) Brow[k] = 0;
        __syncwarp();
    }
    CoopCoder<DEC> rc;
    if constexpr (DEC) { rc.in.start(stream, cap, 16, lane); }
    else { rc.out.start(stream, cap, lane); coop_put_header(rc, h, w, near, k_step, NAVP == 0 ? 1 : (NAVP == 6 ? 2 : 3)); }
    rc.start();
    RowStage<2> rows; /* rows i-1, i-2 of the raster the neighbours come from (for a decoder: what it wrote itself) */
    rows.start(stage_buf, nbimg, h, w);

    for (int i = 0; i < h; i++) {
        int err = 0, x1 = 0, x2 = 0; /* previous two pixels of this row */
        const uint8_t *row = nbimg + (size_t)i * w;
        if constexpr (NAVP > 0) { avp_row_start<AN>(E, Brow, Frow, w, lane); __syncwarp(); }
        for (int j0 = 0; j0 < w; j0 += 32) {
            /* ---------------- phase P: the rows above, lane = pixel j0 + lane ---------------- */
            {
                const int j = min(j0 + lane, w - 1);
                PixRec pr;
                const u32 orig = DEC ? 0u : (u32)src[(size_t)i * w + j];
                if (i >= 2) { /* the tile after this one (same row: the next row's neighbours are not all written yet) is prefetched */
                    if ((j0 & (kStageTile - 1)) == 0) rows.advance(i, j0, j0 + kStageTile < w ? i : -1, j0 + kStageTile, lane, 32, lane == 0);
                    const int jr = j - (j0 & ~(kStageTile - 1));
                    pr = make_pixrec_staged(rows.at(0, jr), rows.at(1, jr), orig);
                } else if (i == 1) pr = make_pixrec(row, w, i, j, orig);
                else { pr.bcdf = pr.ghqr = pr.st_act = pr.k01 = pr.k23 = pr.k45 = pr.k6_lin = 0; pr.orig = orig; }
                recs[lane] = pr;
                __syncwarp();
            }

            /* ---------------- phase S: one pixel at a time ---------------- */
            const int n_here = min(32, w - j0);
            u32 my_x = 0;
            for (int jj = 0; jj < n_here; jj++) {
                const int j = j0 + jj;
                const uint4 ra = *reinterpret_cast<const uint4 *>(&recs[jj]);
                const uint4 rb = *(reinterpret_cast<const uint4 *>(&recs[jj]) + 1);
                Nb nb;
                if (i >= 1) {
                    nb.b = ra.x & 255; nb.c = (ra.x >> 8) & 255; nb.d = (ra.x >> 16) & 255; nb.f = ra.x >> 24;
                    nb.g = ra.y & 255; nb.h = (ra.y >> 8) & 255; nb.q = (ra.y >> 16) & 255; nb.r = ra.y >> 24;
                    nb.s = ra.z & 255; nb.t = (ra.z >> 8) & 255;
                    nb.a = j == 0 ? nb.b : x1;
                    nb.e = j >= 2 ? x2 : nb.a;
                } else { /* first row: every neighbour falls back to the pixel on the left (R: NBLIC.c:288-303) */
                    const int a = j >= 1 ? x1 : 128;
                    nb.a = nb.b = nb.c = nb.d = nb.f = nb.g = nb.h = nb.q = nb.r = nb.s = nb.t = a;
                    nb.e = j >= 2 ? x2 : a;
                }
                int px0 = 0, ok1 = 0, ok2 = 0;
                i64 p1 = 0, p2 = 0, ef0 = 0, r1 = 0, r2 = 0;
                if constexpr (NAVP > 0) { /* R: NBLIC.c:831-846 */
                    if (lane < NAVP) {
                        const int sel = lane == 0 ? nb.a : lane == 1 ? nb.b : lane == 2 ? nb.c : lane == 3 ? nb.d : lane == 4 ? nb.e : lane == 5 ? nb.f :
                                        lane == 6 ? nb.t : lane == 7 ? nb.h : lane == 8 ? nb.q : nb.g;
                        avp_sm->vec[lane] = sel - 128;
                    }
                    r1 = ridge * 21 / 22; r2 = ridge * 22 / 21;
                    r1 = clampl(r1, -1, ridge - 1); r2 = clampl(r2, ridge + 1, N_BIAS_MAX + 1);
                    r1 = clampl(r1, 0, N_BIAS_MAX); r2 = clampl(r2, 0, N_BIAS_MAX);
                    __syncwarp();
                    avp_predict_pair<AN>(*avp_sm, E, Frow + (size_t)j * AM, r1, r2, lane, ok1, ok2, p1, p2, ef0);
                    if (ok1) px0 = (int)((p1 + (1 << (N_FRAC - 1))) >> N_FRAC);
                }
                if (!ok1) { /* the 7-direction gradient predictor */
                    if (i >= 1) {
                        const Pred pt = finish_predictor(nb, ra, rb);
                        px0 = blend_prediction(pt, n_weight(pt.spread));
                    } else {
                        const Pred pt = predictor_terms(nb);
                        px0 = blend_prediction(pt, n_weight(pt.spread));
                    }
                    p1 = (i64)px0 << N_FRAC;
                }
                const int act = i >= 1 ? abs(nb.a - nb.e) + abs(nb.a - nb.c) + (int)(ra.z >> 16) + 2 * abs(err) : activity(nb, err);
                const u32 soft = sm.soft[min(act, 200)];
                const int u = soft & 15, v = (soft >> 4) & 15, wv = (soft >> 8) & 31;
                const int adr = ((u >> 1) << 8) | texture_bits(nb, px0);
                const int c = sm.ctx[adr];
                int px, sign;
                n_bias_apply(c, px0, px, sign);
                const int key = ((px << 1) | sign) * N_RANKS;
                const int room = (int)(((u32)(min(px, 255 - px) + near) * qmagic) >> 16);

                int y, my_tab, my_count;
                coop_rank_fetch<RG>(rank, count, key, lane, my_tab, my_count);
                if constexpr (DEC) {
                    const int z = coop_decode_symbol(rc, sm, forest, k_step, top, ktab, u, v, wv, lane);
                    if (z < 0) return 1u;
                    y = coop_rank_decode(rank, count, key, z, lane, my_tab, my_count);
                } else {
                    const int xo = (int)rb.w;
                    const int mag = (int)(((u32)(abs(xo - px) + near) * qmagic) >> 16);
                    y = mag <= 0 ? 0 : (mag <= room ? 2 * mag - ((xo >= px) ^ sign) : mag + room);
                    const int z = coop_rank_encode(y, my_tab);
                    coop_encode_symbol(rc, sm, forest, k_step, top, ktab, u, v, wv, z, lane);
                    coop_rank_touch_encode(rank, count, key, y, z, lane, my_tab, my_count);
                }

                /* reconstruction (R: NBLIC.c:449-466) */
                int mag, up;
                if (y <= 0) { mag = 0; up = 0; }
                else if (y <= 2 * room) { mag = (y + 1) >> 1; up = (y & 1) ^ sign; }
                else { mag = y - room; up = px < 128; }
                mag *= qn;
                const int x = clampi(up ? px + mag : px - mag, 0, 255);
                if (lane == jj) my_x = (u32)x;
                err = clampi(x - px0, -127, 127);
                if (lane == 0) sm.ctx[adr] = (int16_t)n_bias_learn(c, err);
                x2 = x1; x1 = x;
                if constexpr (NAVP > 0) { /* R: NBLIC.c:882-893 */
                    const i64 target = (i64)x << N_FRAC;
                    const i64 s_now = labs64(p1 - target);
                    avp_learn_coop<AN>(*avp_sm, E, Brow + (size_t)j * AM, x, s_now, ef0 + s_now * 3 / 2, lane);
                    if (ok1 && ok2) ridge = s_now > labs64(p2 - target) ? r2 : r1;
                }
                __syncwarp();
            }
            if (out_rec && lane < n_here) out_rec[(size_t)i * w + j0 + lane] = (uint8_t)my_x; /* one coalesced store per block */
            __syncwarp();
        }
    }
    if constexpr (DEC) return 0u;
    else {
        rc.finish();
        return rc.out.overflow ? 0xffffffffu : rc.out.pos;
    }
}

} /* namespace nblic */
