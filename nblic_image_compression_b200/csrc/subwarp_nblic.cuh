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
 * through.  The rank tables and their frequencies already live in L2 (coop_nblic.cuh, RG).  What stays in shared
 * memory: the compacted counter forest (4 KB lossless), the row cache (512 B), the phase-P records (32 B x LPS).
 *
 * Every arithmetic step is the one coop_feedback<0, true, RG> performs (same helpers); the streams produced by the
 * encoders decode to the same pixels.  R: NBLIC.c:749-908 (decode branch), :640-679 (Zcodec), :527-586 (coder).
 */
#pragma once
#include "coop_nblic.cuh"

namespace nblic {

/* global scratch of one stream: rank-mapper frequencies [512][20] int, rank tables [512][20] bytes, bias table [2048] int16 */
constexpr size_t kSubCountOff = 0, kSubRankOff = (size_t)N_RANK_ENTRIES * 4, kSubCtxOff = kSubRankOff + N_RANK_ENTRIES;
constexpr size_t kSubScratchBytes = kSubCtxOff + (size_t)N_CTX_ENTRIES * 2; /* 55 296 */

template <int LPS, bool CTXG> struct SubLayout {
    static constexpr int SPW = 32 / LPS;
    static constexpr int kCtxEntries = CTXG ? 256 : N_CTX_ENTRIES; /* per stream, int16 */
    static constexpr size_t kSoftOff = 0;                                          /* uint16 soft[208], shared by the streams */
    static constexpr size_t kFbaseOff = 416;                                       /* uint16 fbase[SPW][16]                    */
    static constexpr size_t kRecOff = (kFbaseOff + SPW * 32 + 15) & ~(size_t)15;   /* PixRec recs[LPS][SPW]                    */
    static constexpr size_t kCtxOff = kRecOff + sizeof(PixRec) * 32;               /* int16 ctx[SPW][kCtxEntries]              */
    static constexpr size_t kForestOff = kCtxOff + (size_t)SPW * kCtxEntries * 2;  /* u32 forest[SPW][max_nodes]               */
    static size_t bytes(int max_nodes) { return kForestOff + (size_t)SPW * max_nodes * 4; }
};

NB_DEV void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
NB_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

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
    }
    /* keeps navail >= 4: a step shifts out at most 4 bytes */
    NB_DEV void top_up() {
        if (navail <= 4) { ahead |= (u64)wnext << (32 - 8 * navail); navail += 4; rpos += 4; wnext = word_at(rpos); }
    }
    /* decision with P(1) = p1 / 4096; the registers only move when `commit` */
    NB_DEV int step(u32 p1, bool commit) {
        const u32 mid = lo + split_point(hi - lo, p1);
        const int b = code <= mid;
        if (commit) {
            const u32 nlo = b ? lo : mid + 1, nhi = b ? mid : hi;
            /* the reference shifts one byte at a time while the top bytes agree (R: NBLIC.c:563-572): that is the number of
             * equal leading bytes, at most 4 (then lo = 0, hi = ~0 and the loop stops) */
            const int sh = (__clz((int)(nlo ^ nhi)) >> 3) << 3;
            lo = __funnelshift_lc(0u, nlo, sh);
            hi = __funnelshift_lc(0xffffffffu, nhi, sh);
            code = __funnelshift_lc((u32)(ahead >> 32), code, sh);
            ahead <<= sh;
            navail -= sh >> 3;
        }
        return b;
    }
};

NB_DEV int pick4(const int4 &v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w)); }

/*
 * Decode the pack of up to SPW streams `pack[0..SPW)` (task indices, -1 = empty; pack[0] >= 0; equal h x w).
 * smem: SubLayout<LPS, CTXG>; scratch: SPW x kSubScratchBytes of global memory private to this warp.
 */
template <int LPS, bool CTXG>
__device__ void subwarp_decode_pack(Task *tasks, const int *pack, uint8_t *smem, uint8_t *scratch, int max_nodes, int lane) {
    using L = SubLayout<LPS, CTXG>;
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
    PixRec *recs = reinterpret_cast<PixRec *>(smem + L::kRecOff);
    int16_t *ctx_s = reinterpret_cast<int16_t *>(smem + L::kCtxOff) + sid * L::kCtxEntries;
    u32 *forest = reinterpret_cast<u32 *>(smem + L::kForestOff) + (size_t)sid * max_nodes;
    uint8_t *mine = scratch + (size_t)sid * kSubScratchBytes;
    int *count = reinterpret_cast<int *>(mine + kSubCountOff);
    uint8_t *rank = mine + kSubRankOff;
    int16_t *ctx_g = reinterpret_cast<int16_t *>(mine + kSubCtxOff);

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
        if (CTXG) { for (int k = sl; k < N_CTX_ENTRIES / 8; k += LPS) reinterpret_cast<int4 *>(ctx_g)[k] = make_int4(0, 0, 0, 0); }
        else { for (int k = sl; k < N_CTX_ENTRIES; k += LPS) ctx_s[k] = 0; }
        /* rank -> symbol tables: identity, 20 bytes per key = 5 words with period 5; frequencies 2 * (19 - rank) */
        for (int k = sl; k < N_RANK_ENTRIES / 4; k += LPS) {
            const int r = 4 * (k % 5);
            reinterpret_cast<u32 *>(rank)[k] = 0x03020100u + 0x01010101u * (u32)r;
            reinterpret_cast<int4 *>(count)[k] = make_int4(38 - 2 * r, 36 - 2 * r, 34 - 2 * r, 32 - 2 * r);
        }
        __syncwarp();
    }
    int cur_row = -1; /* class-pair row of the bias table held in ctx_s (CTXG) */

    SubDecoder dec;
    dec.start(t.slot, live ? t.slot_cap : 0u);
    bool corrupt = false;

    for (int i = 0; i < h; i++) {
        int err = 0, x1 = 0, x2 = 0; /* previous two pixels of this row */
        uint8_t *row = img + (size_t)i * w;
        for (int j0 = 0; j0 < w; j0 += LPS) {
            /* ---------------- phase P: the rows above, sub-lane = pixel j0 + sl ---------------- */
            if (i >= 1) recs[sl * SPW + sid] = make_pixrec(row, w, i, min(j0 + sl, w - 1), 0u);
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
                if (CTXG) { /* bias-table row of this class pair: 512 bytes by cp.async behind the predictor */
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
                int c;
                if (CTXG) { cp_async_wait_all(); __syncwarp(); c = ctx_s[tex]; }
                else c = ctx_s[((u >> 1) << 8) | tex];
                int px, sign;
                n_bias_apply(c, px0, px, sign);
                const int key = ((px << 1) | sign) * N_RANKS;
                const int room = (int)(((u32)(min(px, 255 - px) + near) * qmagic) >> 16);

                /* the key's rank table (20 bytes = 5 words) and frequencies (5 x int4) ride in sub-lanes 0..4; issued
                 * before the symbol is decoded, used after */
                u32 symw = 0;
                int4 cnt = make_int4(0, 0, 0, 0);
                if (sl < 5) {
                    symw = __ldcg(reinterpret_cast<const u32 *>(rank + key) + sl);
                    cnt = __ldcg(reinterpret_cast<const int4 *>(count + key) + sl);
                }

                /* ---- symbol: rounds of (evaluate <= 8 candidate nodes, walk them)  R: NBLIC.c:640-679 ---- */
                int mode = 0 /* 0 unary run, 1 suffix, 2 done */, qbase = 0, z = 0, o_cur = 0, r_cur = 0, qz = 0;
                for (;;) {
                    int slot = -1; /* compacted slot of my candidate inside the class (coop_nblic.cuh: compact_node) */
                    if (mode == 0) { /* unary nodes qbase .. qbase + 7: slot = index << k */
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
                    u32 cu = 0, cv = 0, su = 2 * N_MIX, sv = 2 * N_MIX, p = 0;
                    if (slot >= 0) { cu = forest[bu + slot]; cv = forest[bv + slot]; su = pair_sum(cu); sv = pair_sum(cv); p = mixed_p(cu, cv, su, sv, wv); }

                    const int ncand = min(8, n_unary - qbase);
                    bool walking = mode < 2, zero = false;
                    int pos = 0, steps = 0;
                    u32 vis = 0, ones = 0; /* candidates consulted / consulted with outcome 1 */
                    while (__any_sync(FULL, walking)) {
                        dec.top_up();
                        const u32 pt = __shfl_sync(FULL, p, pos, LPS);
                        const int b = dec.step(pt, walking);
                        if (walking) {
                            vis |= 1u << pos; ones |= (u32)b << pos;
                            if (mode == 0) {
                                if (b) { pos++; walking = pos < ncand; } else { zero = true; walking = false; }
                            } else {
                                o_cur += b ? (1 << (r_cur - 1)) : 1;
                                r_cur--;
                                z += b << r_cur;
                                pos = 2 * pos + 1 + b;
                                steps++;
                                walking = steps < 3 && r_cur > 0;
                            }
                        }
                    }
                    if (slot >= 0 && ((vis >> sl) & 1u)) learn_pair(forest, bu + slot, bv + slot, cu, cv, su, sv, wv, (int)((ones >> sl) & 1u));
                    __syncwarp();
                    if (mode == 0) {
                        if (zero) {
                            const int q = qbase + pos;
                            z = q << k;
                            if (k > 0) { mode = 1; o_cur = 1; r_cur = k; qz = q; } else mode = 2;
                        } else {
                            qbase += ncand;
                            if (qbase >= n_unary) { /* escape to the next order (R: NBLIC.c:658-662) */
                                const int uu = (k + 1) * k_step;
                                if (uu >= N_CLASSES) { corrupt = true; mode = 2; z = 0; } /* no encoder output gets here */
                                else { k = k + 1; bu = bv = fb[uu]; qbase = n_unary >> 1; }
                            }
                        }
                    } else if (mode == 1 && r_cur == 0) mode = 2;
                    if (!__any_sync(FULL, mode < 2)) break;
                }

                /* ---- rank mapper, decoder direction (R: NBLIC.c:470-523) ---- */
                int y = z;
                {
                    const int zc = min(z, N_RANKS - 1), zp = max(zc - 1, 0);
                    const u32 wz = __shfl_sync(FULL, symw, zc >> 2, LPS), wp = __shfl_sync(FULL, symw, zp >> 2, LPS);
                    const int cz = __shfl_sync(FULL, pick4(cnt, zc & 3), zc >> 2, LPS) + 1;
                    const int cp = __shfl_sync(FULL, pick4(cnt, zp & 3), zp >> 2, LPS);
                    if (z < N_RANKS) {
                        y = (int)((wz >> (8 * (zc & 3))) & 255u);
                        if (sl == 0) {
                            if (z > 0 && cp < cz) { /* one adjacent promotion */
                                const int other = (int)((wp >> (8 * (zp & 3))) & 255u);
                                count[key + z] = cp; count[key + z - 1] = cz;
                                rank[key + z] = (uint8_t)other; rank[key + z - 1] = (uint8_t)y;
                            } else count[key + z] = cz;
                        }
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
                const int c_new = n_bias_learn(c, err);
                if (CTXG) { /* write through; the lane that fetches an entry's 16-byte piece also stores it (same-thread order) */
                    if (sl == ((tex >> 3) % LPS)) { ctx_s[tex] = (int16_t)c_new; ctx_g[(cur_row << 8) | tex] = (int16_t)c_new; }
                } else if (sl == 0) ctx_s[((u >> 1) << 8) | tex] = (int16_t)c_new;
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
