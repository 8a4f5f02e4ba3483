/*
 * subwarp_nblic.cuh -- NBLIC effort-1 decoder (any near) with SEVERAL coder streams per warp.
 *
 * Why.  A decoder is one dependent chain per image: predict -> bias -> decode symbol -> reconstruct
 * -> next pixel.  With one stream per warp (coop_nblic.cuh: coop_feedback) every lane recomputes the same
 * warp-uniform scalars: ncu counted 537 warp instructions per pixel with 29.8 active lanes each, i.e. a 27-fold lane
 * redundancy on an issue-bound kernel.  Here a warp carries SPW = 32 / LPS streams (LPS = 8 lanes per stream by
 * default), all of the same height x width so the row / column loops stay warp-uniform; the straight-line part of a
 * pixel then issues once for SPW pixels.  Data-dependent trip counts (unary run, suffix depth, renormalisation) are
 * made uniform: symbols are decoded in ROUNDS -- the lanes of a stream evaluate up to 8 candidate counter nodes (the
 * next 8 unary nodes, or the next three levels of the suffix sub-tree in heap order), the coder of every stream then
 * walks its own candidates, and a round lasts as long as its slowest stream; the range coder renormalises without a
 * loop (count of equal leading bytes, funnel shifts, a 64-bit look-ahead of stream bytes).
 *
 * Shared memory per stream is what bounds the resident streams, so the bias table (R: NBLIC.c:60-64, 4 KB) moves to
 * global memory / L2 and only the 256-entry row of the current activity class pair sits in shared memory: the row is
 * known before the predictor runs, is fetched with cp.async (16-byte LDGSTS) behind it, and entries are written
 * through.  The rank tables and their frequencies already live in L2 (coop_nblic.cuh, RG).  With FORESTG the compacted
 * counter forest (4 KB lossless) goes to L2 as well: the first round's candidates are requested before the predictor
 * runs, so only the suffix round pays an L2 round trip (~250 cycles of a ~4000-cycle pixel) -- and a warp then needs
 * 3 KB of shared memory instead of 20 KB, so every stream of a 10 000-image batch is resident at once (17 warps per SM
 * instead of 11, no second wave).
 * Every arithmetic step is the one coop_feedback<0, true, RG> performs (same helpers); the streams produced by the
 * encoders decode to the same pixels.  R: NBLIC.c:749-908 (decode branch), :640-679 (Zcodec), :527-586 (coder).
 */
#pragma once
#include "coop_nblic.cuh"

namespace nblic {

/* global scratch of one stream: rank-mapper frequencies [512][20] int, rank -> symbol tables [512][32] bytes (20 used: rows are
 * 16-byte aligned for cp.async), bias table [2048] int16, counter forest [16 * 256] u32 (used when FORESTG) */
constexpr int kSubRankStride = 32;
constexpr size_t kSubCountOff = 0, kSubRankOff = (size_t)N_RANK_ENTRIES * 4, kSubCtxOff = kSubRankOff + 512 * kSubRankStride;
constexpr size_t kSubForestOff = kSubCtxOff + (size_t)N_CTX_ENTRIES * 2;
constexpr size_t kSubScratchBytes = kSubForestOff + (size_t)N_FOREST_ENTRIES * 4; /* 77 824 */

template <int LPS, bool FORESTG> struct SubLayout {
    static constexpr int SPW = 32 / LPS;
    static constexpr size_t kSoftOff = 0;                                          /* uint16 soft[208], shared by the streams  */
    static constexpr size_t kFbaseOff = 416;                                       /* uint16 fbase[SPW][16]                    */
    static constexpr size_t kRankOff = (kFbaseOff + SPW * 32 + 15) & ~(size_t)15;  /* per stream: int count[20], u8 sym[32] = 112 B */
    static constexpr size_t kRecOff = kRankOff + SPW * 112;                        /* PixRec recs[LPS][SPW]                    */
    static constexpr size_t kCtxOff = kRecOff + sizeof(PixRec) * 32;               /* int16 ctx[SPW][256]: row cache           */
    static constexpr size_t kStageOff = kCtxOff + (size_t)SPW * 256 * 2;           /* per stream: staged tiles of rows i-1, i-2 (row_stage.cuh) */
    static constexpr size_t kStagePerStream = 2 * 2 * kStageLine;
    static constexpr size_t kForestOff = kStageOff + SPW * kStagePerStream;        /* u32 forest[SPW][max_nodes] unless FORESTG */
    static size_t bytes(int max_nodes) { return kForestOff + (FORESTG ? 0 : (size_t)SPW * max_nodes * 4); }
};

/* range decoder of one stream (R: NBLIC.c:541-572), registers uniform across the stream's lanes.  The look-ahead
 * `ahead` holds the next `navail` unread stream bytes left-aligned (first unread byte in bits 63..56); `wnext` is the
 * aligned 32-bit word after them, loaded one merge early so its latency never sits on the coder chain. */
struct SubDecoder {
    const uint8_t *base;
    u32 len, lo, hi, code, rpos, wnext;
    u64 ahead;
    int navail;
    NB_DEV u32 byte_at(u32 p) const { return p < len ? (u32)__ldg(base + p) : 0u; } /* the reference reads on; we read zeros */
    NB_DEV u32 word_at(u32 p) const { /* base + p is 4-byte aligned; big-endian; an aligned word that holds a valid byte is mapped */
        u32 w = 0;
        if (p < len) {
            w = __byte_perm(__ldg(reinterpret_cast<const u32 *>(base + p)), 0, 0x0123);
            if (p + 4 > len) w &= 0xffffffffu << (8u * (p + 4 - len));
        }
        return w;
    }
    NB_DEV void start(const uint8_t *b, u32 n) {
        base = b; len = n; lo = 0; hi = 0xffffffffu; code = 0; ahead = 0; navail = 0;
        u32 p = 16;
        for (int k = 0; k < 4; k++) code = (code << 8) | byte_at(p++);
        while ((unsigned long long)(base + p) & 3ull) { ahead |= (u64)byte_at(p++) << (56 - 8 * navail); navail++; }
        rpos = p;
        wnext = word_at(rpos);
        refill(); /* navail <= 3 here */
    }
    /* merge the next word: navail <= 4 before, >= 4 after */
    NB_DEV void refill() { ahead |= (u64)wnext << (32 - 8 * navail); navail += 4; rpos += 4; wnext = word_at(rpos); }
    /* Decision with P(1) = p1 / 4096; the registers only move when `commit`.  navail >= 4 on entry (a step shifts out at
     * most 4 bytes); about one step in forty needs the refill, so it is a real branch to an out-of-line body. */
    NB_DEV int step(u32 p1, bool commit) {
        if (navail < 4) refill();
        const u32 mid = lo + split_point(hi - lo, p1);
        const int b = code <= mid;
        const u32 nlo = (commit && !b) ? mid + 1 : lo, nhi = (commit && b) ? mid : hi;
        /* the reference shifts one byte at a time while the top bytes agree (R: NBLIC.c:563-572): that is the number of
         * equal leading bytes, at most 4 (then lo = 0, hi = ~0 and the loop stops).  Without commit the registers are
         * already normalised (top bytes differ) and the shift is 0. */
        const int sh = (__clz((int)(nlo ^ nhi)) >> 3) << 3;
        lo = __funnelshift_lc(0u, nlo, sh);
        hi = __funnelshift_lc(0xffffffffu, nhi, sh);
        code = __funnelshift_lc((u32)(ahead >> 32), code, sh);
        ahead <<= sh;
        navail -= sh >> 3;
        return b;
    }
};

/* learn_pair of coop_nblic.cuh with the stores left to the caller: the node of the main class after the decision, and
 * the node of the side class (when both classes are the same node, the main result learns the second weight too and
 * `cv2` is what has to be stored).  R: NBLIC.c:606-617,633-636 */
NB_DEV void learn_two(u32 cu, u32 cv, u32 su, u32 sv, int wv, int bit, bool same, u32 &cu2, u32 &cv2) {
    cu2 = learn_packed(cu, su, bit, N_MIX - wv);
    const u32 src = same ? cu2 : cv, ssrc = same ? pair_sum(cu2) : sv;
    cv2 = learn_packed(src, ssrc, bit, wv);
}

/*
 * Decode the pack of up to SPW streams `pack[0..SPW)` (task indices, -1 = empty; pack[0] >= 0; equal h x w).
 * smem: SubLayout<LPS, FORESTG>; scratch: SPW x kSubScratchBytes of global memory private to this warp.
 */
template <int LPS, bool FORESTG>
__device__ void subwarp_decode_pack(Task *tasks, const int *pack, uint8_t *smem, uint8_t *scratch, int max_nodes, int lane) {
    using L = SubLayout<LPS, FORESTG>;
    constexpr int SPW = L::SPW;
    const int sl = lane % LPS, sid = lane / LPS;
    const int ti = pack[sid];
    const bool live = ti >= 0;
    const Task &t0 = tasks[pack[0]];
    const Task &t = tasks[live ? ti : pack[0]];
    const int h = t0.h, w = t0.w;
    const int near = t.near, k_step = t.k_step;
    uint8_t *img = t.rec; /* an empty slot mirrors stream 0 for its loads and never stores */

    uint16_t *soft_tab = reinterpret_cast<uint16_t *>(smem + L::kSoftOff);
    uint16_t *fb = reinterpret_cast<uint16_t *>(smem + L::kFbaseOff) + sid * 16;
    int *rk_count = reinterpret_cast<int *>(smem + L::kRankOff + sid * 112); /* the current key's 20 frequencies ...   */
    uint8_t *rk_sym = reinterpret_cast<uint8_t *>(rk_count + N_RANKS);          /* ... and its rank -> symbol table        */
    PixRec *recs = reinterpret_cast<PixRec *>(smem + L::kRecOff);
    int16_t *ctx_s = reinterpret_cast<int16_t *>(smem + L::kCtxOff) + sid * 256;
    uint8_t *mine = scratch + (size_t)sid * kSubScratchBytes;
    int *count = reinterpret_cast<int *>(mine + kSubCountOff);
    uint8_t *rank = mine + kSubRankOff;
    int16_t *ctx_g = reinterpret_cast<int16_t *>(mine + kSubCtxOff);
    u32 *forest = FORESTG ? reinterpret_cast<u32 *>(mine + kSubForestOff) : reinterpret_cast<u32 *>(smem + L::kForestOff) + (size_t)sid * max_nodes;
    auto node = [&](int at) -> u32 { return FORESTG ? __ldcg(forest + at) : forest[at]; };

    const int top = (N_CLASSES - 1) / k_step, n_unary = 256 >> top;
    const u32 ktab = make_order_table(k_step);
    const int qn = 2 * near + 1;
    const u32 qmagic = 65536u / (u32)qn + 1u; /* n / qn == (n * qmagic) >> 16 for 0 <= n < 3400 */

    /* ---- reset (R: NBLIC.c:795-804) ---- */
    {
        const int n_nodes = forest_nodes(k_step);
        for (int k = sl; k < n_nodes; k += LPS) forest[k] = (u32)N_MIX | ((u32)N_MIX << 16);
        if (sl == 0) { int b = 0; for (int u = 0; u < N_CLASSES; u++) { fb[u] = (uint16_t)b; b += (256 >> top) << (u / k_step); } }
        for (int d = lane; d <= 200; d += 32) { int u, v, wv; n_soft_class(d, u, v, wv); soft_tab[d] = (uint16_t)(u | (v << 4) | (wv << 8)); }
        for (int k = sl; k < N_CTX_ENTRIES / 8; k += LPS) reinterpret_cast<int4 *>(ctx_g)[k] = make_int4(0, 0, 0, 0);
        /* rank -> symbol tables: identity (5 words of a 32-byte row); frequencies 2 * (19 - rank), 5 x int4 per key */
        for (int k = sl; k < 512 * kSubRankStride / 4; k += LPS) reinterpret_cast<u32 *>(rank)[k] = (k & 7) < 5 ? 0x03020100u + 0x04040404u * (u32)(k & 7) : 0u;
        for (int k = sl; k < N_RANK_ENTRIES / 4; k += LPS) {
            const int r = 4 * (k % 5);
            reinterpret_cast<int4 *>(count)[k] = make_int4(38 - 2 * r, 36 - 2 * r, 34 - 2 * r, 32 - 2 * r);
        }
        __syncwarp();
    }
    int cur_row = -1; /* class-pair row of the bias table held in ctx_s */
    RowStage<2> rows; /* the two rows above, of this stream's own output */
    rows.start(smem + L::kStageOff + sid * L::kStagePerStream, img, h, w);

    SubDecoder dec;
    dec.start(t.slot, live ? t.slot_cap : 0u);
    bool corrupt = false;

    for (int i = 0; i < h; i++) {
        int err = 0, x1 = 0, x2 = 0; /* previous two pixels of this row */
        uint8_t *row = img + (size_t)i * w;
        for (int j0 = 0; j0 < w; j0 += LPS) {
            /* ---------------- phase P: the rows above, sub-lane = pixel j0 + sl ---------------- */
            if (i >= 2) { /* cp.async tiles of 64 pixels, the next one of this row in flight while this one is decoded */
                if ((j0 & (kStageTile - 1)) == 0) rows.advance(i, j0, j0 + kStageTile < w ? i : -1, j0 + kStageTile, sl, LPS, sl == 0);
                const int jr = min(j0 + sl, w - 1) - (j0 & ~(kStageTile - 1));
                recs[sl * SPW + sid] = make_pixrec_staged(rows.at(0, jr), rows.at(1, jr), 0u);
            } else if (i == 1) recs[sl * SPW + sid] = make_pixrec(row, w, i, min(j0 + sl, w - 1), 0u);
            __syncwarp();

            /* ---------------- phase S: one pixel of every stream at a time ---------------- */
            const int n_here = min(LPS, w - j0);
            u32 my_x = 0;
            for (int jj = 0; jj < n_here; jj++) {
                const int j = j0 + jj;
                Nb nb;
                uint4 ra = make_uint4(0, 0, 0, 0), rb = ra;
                int act;
                if (i >= 1) {
                    const uint4 *rp = reinterpret_cast<const uint4 *>(&recs[jj * SPW + sid]);
                    ra = rp[0]; rb = rp[1];
                    nb.b = ra.x & 255; nb.c = (ra.x >> 8) & 255; nb.d = (ra.x >> 16) & 255; nb.f = ra.x >> 24;
                    nb.g = ra.y & 255; nb.h = (ra.y >> 8) & 255; nb.q = (ra.y >> 16) & 255; nb.r = ra.y >> 24;
                    nb.s = ra.z & 255; nb.t = (ra.z >> 8) & 255;
                    nb.a = j == 0 ? nb.b : x1;
                    nb.e = j >= 2 ? x2 : nb.a;
                    act = abs(nb.a - nb.e) + abs(nb.a - nb.c) + (int)(ra.z >> 16) + 2 * abs(err);
                } else { /* first row: every neighbour falls back to the pixel on the left (R: NBLIC.c:288-303) */
                    const int a = j >= 1 ? x1 : 128;
                    nb.a = nb.b = nb.c = nb.d = nb.f = nb.g = nb.h = nb.q = nb.r = nb.s = nb.t = a;
                    nb.e = j >= 2 ? x2 : a;
                    act = activity(nb, err);
                }
                /* class pair, Golomb order, forest bases: known before the predictor (R: NBLIC.c:373-395, :640-645) */
                const u32 soft = soft_tab[min(act, 200)];
                const int u = soft & 15, wv = (soft >> 8) & 31;
                int v = (soft >> 4) & 15;
                int k = order_of(ktab, u);
                if (order_of(ktab, v) != k) v = u;
                int bu = fb[u], bv = fb[v];
                /* first round's candidates = unary nodes 0..7 (slot = index << k): requested now, evaluated after the predictor */
                int slot = (sl < 8 && sl < n_unary) ? (sl << k) : -1;
                u32 cu = 0, cv = 0;
                if (slot >= 0) { cu = node(bu + slot); cv = node(bv + slot); }
                { /* bias-table row of this class pair: 512 bytes by cp.async behind the predictor */
                    const int want = u >> 1;
                    if (want != cur_row) {
                        const uint8_t *src = reinterpret_cast<const uint8_t *>(ctx_g + (want << 8));
                        uint8_t *dst = reinterpret_cast<uint8_t *>(ctx_s);
#pragma unroll
                        for (int m = 0; m < 32 / LPS; m++) cp_async16(dst + 16 * (sl + LPS * m), src + 16 * (sl + LPS * m));
                        cur_row = want;
                    }
                }
                int px0;
                if (i >= 1) { const Pred pt = finish_predictor(nb, ra, rb); px0 = blend_prediction(pt, n_weight(pt.spread)); }
                else { const Pred pt = predictor_terms(nb); px0 = blend_prediction(pt, n_weight(pt.spread)); }
                const int tex = texture_bits(nb, px0);
                cp_async_wait_all();
                __syncwarp();
                const int c = ctx_s[tex];
                int px, sign;
                n_bias_apply(c, px0, px, sign);
                const int kidx = (px << 1) | sign, key = kidx * N_RANKS, rkey = kidx * kSubRankStride;
                const int room = (int)(((u32)(min(px, 255 - px) + near) * qmagic) >> 16);

                /* the key's frequencies (5 x 16 bytes) and rank table (2 x 16 bytes) travel from L2 to the stream's staging
                 * area by cp.async: requested before the symbol is decoded, waited for after */
                if (sl < 5) cp_async16(reinterpret_cast<int4 *>(rk_count) + sl, reinterpret_cast<const int4 *>(count + key) + sl);
                else if (sl < 7) cp_async16(rk_sym + 16 * (sl - 5), rank + rkey + 16 * (sl - 5));

                /* ---- symbol: rounds of (evaluate <= 8 candidate nodes, walk them)  R: NBLIC.c:640-679 ---- */
                int mode = 0 /* 0 unary run, 1 suffix, 2 done */, qbase = 0, z = 0, o_cur = 0, r_cur = 0, qz = 0;
                for (;;) {
                    u32 su = 2 * N_MIX, sv = 2 * N_MIX, p = 0;
                    if (slot >= 0) { su = pair_sum(cu); sv = pair_sum(cv); p = mixed_p(cu, cv, su, sv, wv); }
                    /* walk: unary = candidates 0, 1, 2 ... until a 0 decision; suffix = heap positions 0 -> 1 + b -> 3 + 2 b + b' */
                    const int limit = mode == 0 ? min(8, n_unary - qbase) : min(3, r_cur);
                    bool walking = mode < 2;
                    int pos = 0, steps = 0;
                    while (__any_sync(FULL, walking)) {
                        const u32 pt = __shfl_sync(FULL, p, pos, LPS);
                        const int b = dec.step(pt, walking);
                        if (walking) {
                            steps++;
                            if (mode == 0) { pos += b; walking = b && pos < limit; }
                            else { pos = 2 * pos + 1 + b; walking = steps < limit; }
                        }
                    }
                    /* which candidates were consulted, and with which outcome */
                    bool seen = false;
                    int my_bit = 0;
                    if (mode == 0) { seen = sl < steps; my_bit = sl < pos; } /* `pos` ones, then (if steps > pos) the closing 0 */
                    else if (mode == 1) { /* heap node sl lies on the path iff it is an ancestor of the position reached */
                        const int lvl = (sl >= 1) + (sl >= 3), up = steps - lvl;
                        seen = up >= 1 && ((pos + 1) >> up) == sl + 1;
                        my_bit = ((pos + 1) >> max(up - 1, 0)) & 1;
                    }
                    if (slot >= 0 && seen) {
                        u32 cu2, cv2;
                        const bool same = bu == bv;
                        learn_two(cu, cv, su, sv, wv, my_bit, same, cu2, cv2);
                        forest[bv + slot] = cv2;
                        if (!same) forest[bu + slot] = cu2;
                    }
                    __syncwarp();
                    if (mode == 0) {
                        if (steps > pos) { /* a 0 closed the run at candidate pos */
                            const int q = qbase + pos;
                            z = q << k;
                            if (k > 0) { mode = 1; o_cur = 1; r_cur = k; qz = q; } else mode = 2;
                        } else {
                            qbase += limit;
                            if (qbase >= n_unary) { /* escape to the next order (R: NBLIC.c:658-662) */
                                const int uu = (k + 1) * k_step;
                                if (uu >= N_CLASSES) { corrupt = true; mode = 2; z = 0; } /* no encoder output gets here */
                                else { k = k + 1; bu = bv = fb[uu]; qbase = n_unary >> 1; }
                            }
                        }
                    } else if (mode == 1) { /* the `steps` bits taken, first one highest: pos + 1 = 1 b b' b'' in binary */
                        const int bits = pos + 1 - (1 << steps);
                        z += bits << (r_cur - steps);
                        o_cur += (bits << (r_cur - steps)) + steps - __popc(bits); /* a 1 skips 2^(levels below) nodes, a 0 one */
                        r_cur -= steps;
                        if (r_cur == 0) mode = 2;
                    }
                    if (!__any_sync(FULL, mode < 2)) break;
                    /* next round's candidates */
                    slot = -1;
                    if (mode == 0) { /* unary nodes qbase .. qbase + 7 */
                        if (sl < 8 && qbase + sl < n_unary) slot = (qbase + sl) << k;
                    } else if (mode == 1 && sl < 7) { /* heap position sl of the sub-tree rooted at pre-order offset o_cur, r_cur levels left */
                        const int lvl = (sl >= 1) + (sl >= 3);
                        if (lvl < r_cur) {
                            int off = o_cur;
                            if (lvl >= 1) off += ((lvl == 1 ? sl - 1 : (sl - 3) >> 1) ? (1 << (r_cur - 1)) : 1);
                            if (lvl == 2) off += (((sl - 3) & 1) ? (1 << (r_cur - 2)) : 1);
                            slot = (qz << k) + off;
                        }
                    }
                    if (slot >= 0) { cu = node(bu + slot); cv = node(bv + slot); }
                }

                /* ---- rank mapper, decoder direction (R: NBLIC.c:470-523) ---- */
                cp_async_wait_all();
                __syncwarp();
                int y = z;
                if (z < N_RANKS) {
                    y = rk_sym[z];
                    if (sl == 0) {
                        const int cz = rk_count[z] + 1;
                        if (z > 0 && rk_count[z - 1] < cz) { /* one adjacent promotion */
                            count[key + z] = rk_count[z - 1]; count[key + z - 1] = cz;
                            rank[rkey + z] = rk_sym[z - 1]; rank[rkey + z - 1] = (uint8_t)y;
                        } else count[key + z] = cz;
                    }
                }

                /* ---- reconstruction (R: NBLIC.c:449-466) ---- */
                int mag, up;
                if (y <= 0) { mag = 0; up = 0; }
                else if (y <= 2 * room) { mag = (y + 1) >> 1; up = (y & 1) ^ sign; }
                else { mag = y - room; up = px < 128; }
                mag *= qn;
                const int x = clampi(up ? px + mag : px - mag, 0, 255);
                if (sl == jj) my_x = (u32)x;
                err = clampi(x - px0, -127, 127);
                /* write through; the lane that fetches an entry's 16-byte piece also stores it (same-thread order) */
                if (sl == ((tex >> 3) % LPS)) { const int c_new = n_bias_learn(c, err); ctx_s[tex] = (int16_t)c_new; ctx_g[(cur_row << 8) | tex] = (int16_t)c_new; }
                x2 = x1; x1 = x;
                __syncwarp();
            }
            if (live && sl < n_here) row[j0 + sl] = (uint8_t)my_x; /* one store of LPS bytes per stream and block */
            __syncwarp();
        }
    }
    if (live && sl == 0 && corrupt) tasks[ti].status = NBLIC_B200_CORRUPT;
    __syncwarp();
}

} /* namespace nblic */
