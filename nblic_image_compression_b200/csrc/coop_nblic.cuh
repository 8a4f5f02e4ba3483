/*
 * coop_nblic.cuh -- warp-cooperative NBLIC effort-1 coder: one coder stream per warp, the 32 lanes
 * share the work of every pixel.
 *
 *   phase P (lanes = 32 consecutive pixels)   everything that depends only on already-known pixels:
 *                                             neighbourhood sampling, the 7-direction predictor, activity,
 *                                             soft context class, texture bits            R: NBLIC.c:287-410
 *   phase S (one pixel at a time, warp-uniform registers)
 *       bias cancel + residual fold           ctx table in shared memory                  R: NBLIC.c:413-466
 *       rank mapper                           lane s holds rank/count of symbol / rank s  R: NBLIC.c:470-523
 *       binarisation                          lane d owns decision d of the pixel: node address,
 *                                             both counter pairs, mixed probability, counter
 *                                             update -- all decisions of a pixel at once  R: NBLIC.c:589-679
 *       range coder                           consumes (bit, p) of lane 0..D-1 via __shfl R: NBLIC.c:552-586
 *
 * The lossless encoder runs phase P on the input image 32 pixels ahead; the decoder and the
 * near-lossless encoder (reconstruction feedback) use the serial kernels of codec_core.cuh.
 * Pixels whose Golomb code escapes to the next order (0.1-0.5 %) take the sequential routine.
 */
#pragma once
#include "codec_core.cuh"

namespace nblic {

constexpr unsigned FULL = 0xffffffffu;

/* shared-memory image of one stream's adaptive state (one warp = one CTA) */
struct CoopSmem {
    u32 forest[N_FOREST_ENTRIES];      /* 16 KB: node counters n0 | n1 << 16                          */
    int16_t ctx[N_CTX_ENTRIES];        /*  4 KB: bias-cancel table                                    */
    uint8_t rank[N_RANK_ENTRIES];      /* 10 KB: encoder: symbol -> rank; decoder: rank -> symbol       */
    uint16_t soft[208];                /* activity (clamped to 200) -> u | v << 4 | wv << 8             */
};

/* range coder whose registers are warp-uniform; only the leader lane stores bytes */
struct CoopEncoder {
    u32 lo, hi;
    uint8_t *wr, *wr_end;
    bool overflow, leader;
    NB_DEV void start(uint8_t *p, uint8_t *end, bool lead) { lo = 0; hi = 0xffffffffu; wr = p; wr_end = end; overflow = false; leader = lead; }
    NB_DEV void put(u32 byte) { if (wr < wr_end) { if (leader) *wr = (uint8_t)byte; } else overflow = true; wr++; }
    NB_DEV void bit(int b, u32 p1) {
        const u32 span = hi - lo;
        const u32 mid = lo + (span >> 12) * p1 + (((span & 0xfffu) * p1) >> 12);
        if (b) hi = mid; else lo = mid + 1;
        while (((lo ^ hi) & 0xff000000u) == 0) { put(hi >> 24); lo <<= 8; hi = (hi << 8) | 0xffu; }
    }
    NB_DEV void finish() { for (int k = 0; k < 4; k++) { put(lo >> 24); lo <<= 8; } }
};

NB_DEV u32 learn_packed(u32 c, int bit, int weight) { /* node_learn on the packed pair */
    c += (u32)weight << (bit ? 16 : 0);
    if ((c & 0xffffu) + (c >> 16) > (u32)(N_MIX * 256)) c = ((c + 0x00010001u) >> 1) & 0x7fff7fffu;
    return c;
}

/* Order tables of one stream: k = u / k_step for u = 0..15, 4 bits each. */
NB_DEV unsigned long long make_order_table(int k_step) {
    unsigned long long t = 0;
    for (int u = 0; u < N_CLASSES; u++) t |= (unsigned long long)(u / k_step) << (4 * u);
    return t;
}
NB_DEV int order_of(unsigned long long tab, int u) { return (int)((tab >> (4 * u)) & 15u); }

/*
 * Lossless effort-1 encode of one image by one warp.  `count` is the stream's rank-mapper frequency
 * table in global memory ([512][20] int, indexed by rank).  Returns the stream length in bytes or
 * 0xffffffff on overflow (identical in all lanes).
 */
__device__ u32 coop_e1_encode_lossless(const uint8_t *img, int h, int w, uint8_t *stream, u32 cap, CoopSmem &sm, int *count, int lane) {
    const int k_step = 3, top = (N_CLASSES - 1) / k_step; /* near = 0 (R: NBLIC.c:769) */
    const unsigned long long ktab = make_order_table(k_step);

    /* ---- reset the adaptive state ---- */
    for (int k = lane; k < N_FOREST_ENTRIES; k += 32) sm.forest[k] = (u32)N_MIX | ((u32)N_MIX << 16);
    for (int k = lane; k < N_CTX_ENTRIES; k += 32) sm.ctx[k] = 0;
    for (int k = lane; k < N_RANK_ENTRIES; k += 32) { const int r = k % N_RANKS; sm.rank[k] = (uint8_t)r; count[k] = 2 * (N_RANKS - 1 - r); }
    for (int d = lane; d <= 200; d += 32) { int u, v, wv; n_soft_class(d, u, v, wv); sm.soft[d] = (uint16_t)(u | (v << 4) | (wv << 8)); }
    __syncwarp();

    CoopEncoder rc;
    if (lane == 0) {
        const char magic[8] = {'N', 'B', 'L', 'I', 'C', '0', '.', '3'};
        for (int k = 0; k < 8; k++) stream[k] = (uint8_t)magic[k];
        stream[8] = 1; stream[9] = (uint8_t)(h >> 8); stream[10] = (uint8_t)h; stream[11] = (uint8_t)(w >> 8); stream[12] = (uint8_t)w;
        stream[13] = 0; stream[14] = (uint8_t)k_step; stream[15] = 1;
    }
    rc.start(stream + 16, stream + cap, lane == 0);

    for (int i = 0; i < h; i++) {
        int carry_px0 = 0; /* px0 of the pixel left of this block */
        for (int j0 = 0; j0 < w; j0 += 32) {
            /* ---------------- phase P: lane = pixel j0 + lane ---------------- */
            const int j = min(j0 + lane, w - 1);
            Nb nb;
            sample_positional(img, w, i, j, nb);
            const Pred pt = predictor_terms(nb);
            const int px0 = blend_prediction(pt, n_weight(pt.spread));
            const int x = img[(size_t)i * w + j];
            int px0_left = __shfl_up_sync(FULL, px0, 1);
            if (lane == 0) px0_left = carry_px0;
            const int err_in = j == 0 ? 0 : clampi(nb.a - px0_left, -127, 127); /* nb.a is the coded value of pixel j-1 */
            const u32 soft = sm.soft[min(activity(nb, err_in), 200)];
            const u32 adr = (((soft & 15u) >> 1) << 8) | (u32)texture_bits(nb, px0);
            const u32 rec0 = (u32)px0 | ((u32)x << 8) | (soft << 16); /* px0:8 x:8 u:4 v:4 wv:5 */
            carry_px0 = __shfl_sync(FULL, px0, 31);

            /* ---------------- phase S: one pixel at a time ---------------- */
            const int n_here = min(32, w - j0);
            for (int jj = 0; jj < n_here; jj++) {
                const u32 r0 = __shfl_sync(FULL, rec0, jj);
                const int adr_s = (int)__shfl_sync(FULL, adr, jj);
                const int s_px0 = r0 & 255, s_x = (r0 >> 8) & 255;
                const int u = (r0 >> 16) & 15;
                int v = (r0 >> 20) & 15;
                const int wv = (r0 >> 24) & 31;

                /* bias cancel, residual fold, context update (R: NBLIC.c:413-466) */
                const int c = sm.ctx[adr_s];
                int px, sign;
                n_bias_apply(c, s_px0, px, sign);
                const int room = min(px, 255 - px), mag = abs(s_x - px);
                const int y = mag == 0 ? 0 : (mag <= room ? 2 * mag - ((s_x >= px) ^ sign) : mag + room);
                if (lane == 0) sm.ctx[adr_s] = (int16_t)n_bias_learn(c, clampi(s_x - s_px0, -127, 127));

                /* rank mapper: lane s < 20 holds rank_of[s] and count[s] of this (px, sign) key */
                const int key = ((px << 1) | sign) * N_RANKS;
                const int my_rank = lane < N_RANKS ? (int)sm.rank[key + lane] : 255;
                const int my_count = lane < N_RANKS ? __ldcg(count + key + lane) : 0;
                const int z_ranked = __shfl_sync(FULL, my_rank, y & 31);
                const int z = y < N_RANKS ? z_ranked : y;

                /* binarisation (R: NBLIC.c:640-679) */
                const int k = order_of(ktab, u);
                if (order_of(ktab, v) != k) v = u;
                const int q = z >> k;
                if (q < (256 >> top)) {
                    const int D = q + 1 + k; /* decisions of this pixel, one per lane */
                    int node, bit;
                    if (lane <= q) { node = lane << top; bit = lane < q; }
                    else {
                        const int t = lane - q - 1, kk = k - 1 - t;          /* t-th suffix bit, weight 2^kk */
                        const int hi_bits = (z & ((1 << k) - 1)) & ~((2 << max(kk, 0)) - 1);
                        node = (q << top) + 1 + hi_bits + t - __popc(hi_bits);
                        bit = (z >> max(kk, 0)) & 1;
                    }
                    u32 coded = 0;
                    if (lane < D) {
                        u32 *nu = sm.forest + u * 256 + node, *nv = sm.forest + v * 256 + node;
                        const u32 cu = *nu, cv = *nv;
                        const int p = (node_p1(cu) * (N_MIX - wv) + node_p1(cv) * wv + N_MIX / 2) >> 5;
                        coded = (u32)clampi(p, 1, N_PROB_ONE - 1) | ((u32)bit << 12);
                        if (u == v) *nu = learn_packed(learn_packed(cu, bit, N_MIX - wv), bit, wv);
                        else { *nu = learn_packed(cu, bit, N_MIX - wv); *nv = learn_packed(cv, bit, wv); }
                    }
                    for (int d = 0; d < D; d++) {
                        const u32 cd = __shfl_sync(FULL, coded, d);
                        rc.bit((int)(cd >> 12), cd & 0xfffu);
                    }
                } else { /* order escape: sequential routine on the leader, coder registers re-broadcast */
                    if (lane == 0) {
                        RangeCoder<false> seq;
                        seq.lo = rc.lo; seq.hi = rc.hi; seq.wr = rc.wr; seq.wr_end = rc.wr_end; seq.overflow = rc.overflow;
                        golomb_symbol<false>(seq, k_step, sm.forest, u, v, wv, z);
                        rc.lo = seq.lo; rc.hi = seq.hi; rc.wr = seq.wr; rc.overflow = seq.overflow;
                    }
                    rc.lo = __shfl_sync(FULL, rc.lo, 0); rc.hi = __shfl_sync(FULL, rc.hi, 0);
                    rc.wr = stream + __shfl_sync(FULL, (u32)(rc.wr - stream), 0);
                    rc.overflow = __shfl_sync(FULL, (int)rc.overflow, 0) != 0;
                }

                /* rank mapper update (R: NBLIC.c:500-523) */
                if (y < N_RANKS) {
                    const int cz = __shfl_sync(FULL, my_count, z) + 1;
                    const int cp = __shfl_sync(FULL, my_count, max(z - 1, 0));
                    const unsigned holders = __ballot_sync(FULL, my_rank == z - 1);
                    if (lane == 0) {
                        if (z > 0 && cp < cz) { /* one adjacent promotion */
                            const int other = __ffs(holders) - 1;
                            count[key + z] = cp; count[key + z - 1] = cz;
                            sm.rank[key + y] = (uint8_t)(z - 1); sm.rank[key + other] = (uint8_t)z;
                        } else count[key + z] = cz;
                    }
                }
                __syncwarp();
            }
        }
    }
    rc.finish();
    return rc.overflow ? 0xffffffffu : (u32)(rc.wr - stream);
}

} /* namespace nblic */
