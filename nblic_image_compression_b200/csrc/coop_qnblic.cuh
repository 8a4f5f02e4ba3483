/*
 * coop_qnblic.cuh -- warp-cooperative QNBLIC ("Q0.2", effort 0): one stream per warp.
 *
 * Encoder (R: QNBLIC.c:562-655).  Pass 1 is data parallel except for the bias table: lanes take 32
 * consecutive pixels, compute the whole front end, then apply / train the bias entries in rounds --
 * __match_any groups the lanes that hit the same table address, and round r lets the r-th member of
 * every group through, so each address sees its pixels in raster order while different addresses
 * proceed together.  Histogram counts are shared-memory atomics (order-free).  The 12 histograms are
 * normalised, described and accumulated by 12 lanes side by side.  Pass 2, the reverse rANS sweep, is
 * the only sequential chain: lanes prefetch 32 symbols' (freq, cumulative) and a 32-bit reciprocal of
 * freq, the state update then costs one multiply-high, one multiply-back and a correction.
 *
 * Decoder (R: QNBLIC.c:493-555): per pixel sequential (the prediction needs the previous pixel), with
 * the two rows above pre-reduced by 32 lanes (coop_nblic.cuh: make_pixrec) and the symbol found by two
 * ballots over the cumulative table (32 coarse + 8 fine probes) instead of the reference's 384 KB LUT.
 *
 * Rows 0 and 1 keep QNBLIC's literal shift-register neighbourhood (R: QNBLIC.c:67-79), which differs
 * from positional sampling there; from row 2 on the two agree except e at column 1 (= the old a).
 */
#pragma once
#include "codec_core.cuh"
#include "coop_nblic.cuh"

namespace nblic {

struct QCoopSmem { /* encoder: only the bias table is latency critical; the 12 x 256 counters (then freq | cumulative << 16)
                     * live in global memory / L2 -- pass 1 adds to them with fire-and-forget atomics, pass 2 fetches 32
                     * symbols' entries one block ahead -- which doubles the resident streams per SM (17 instead of 8) */
    int ctx[Q_CTX_ENTRIES]; /* 12 KB bias-cancel table */
    __align__(16) uint8_t stage[RowStage<3, 32>::kBytes]; /* rows i, i-1, i-2: cp.async tiles of 32 pixels (row_stage.cuh); 384 B keeps 17 streams per SM */
};
enum { Q_CUM_STRIDE = 258 }; /* 257 cumulative frequencies per class (freq = difference), padded to 4 bytes */
struct QDecSmem { /* decoder: 19.2 KB instead of 25 KB, 11 instead of 8 resident streams per SM */
    int ctx[Q_CTX_ENTRIES];                 /* 12 KB bias-cancel table                            */
    uint16_t cum[Q_CLASSES * Q_CUM_STRIDE]; /*  6 KB cumulative frequencies, cum[256] = 2^15      */
    PixRec rec[32];                         /*  1 KB phase-P records (and scratch while the histograms are parsed) */
    __align__(16) uint8_t stage[2 * 2 * kStageLine]; /* rows i-1, i-2: cp.async tiles (row_stage.cuh) */
};

/* rows 0 and 1 (and any row of a sequential fallback): the reference loop, warp-uniform, leader stores */
template <bool DEC, class Smem, class OnPixel>
NB_DEV void q_serial_row(const uint8_t *img, int w, int i, Smem &sm, int lane, OnPixel on_pixel) {
    Nb nb;
    int err = 0;
    sample_positional(img, w, i, 0, nb);
    for (int j = 0; j < w; j++) {
        const Pred pt = predictor_terms(nb);
        const int px0 = blend_prediction(pt, q_weight(pt.spread));
        const int cls = q_class(activity(nb, err));
        const int adr = q_ctx_address(nb, px0, cls);
        const int c = sm.ctx[adr];
        int px, sign;
        q_bias_apply(c, px0, px, sign);
        const int x = on_pixel(j, cls, px, sign); /* encoder: reads x and records y; decoder: decodes x */
        err = x - px0;
        __syncwarp();
        if (lane == 0) sm.ctx[adr] = q_bias_learn(c, err);
        __syncwarp();
        q_window_shift(img, w, i, j, x, nb);
    }
}

/* writes the 12 histogram descriptions; lanes 0..11 work side by side.  Returns the advanced word index. */
NB_DEV u32 q_finish_histograms(u32 *tab, uint16_t *out, u32 o, u32 cap, int lane) {
    u32 mine = 0;
    if (lane < Q_CLASSES) {
        q_normalise(tab + lane * 256);
        mine = q_put_hist(out, 0, 0, tab + lane * 256); /* dry run: cap 0 suppresses the stores */
    }
    u32 before = mine; /* exclusive prefix over lanes */
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) { const u32 t = __shfl_up_sync(FULL, before, d); if (lane >= d) before += t; }
    before -= mine;
    if (lane < Q_CLASSES) {
        q_put_hist(out + o + before, 0, o + before < cap ? cap - (o + before) : 0, tab + lane * 256);
        q_pack_cumulative(tab + lane * 256);
    }
    const u32 total = __shfl_sync(FULL, before + mine, Q_CLASSES - 1);
    __syncwarp();
    return o + total;
}

__device__ bool coop_q_finish(const uint16_t *sym, int h, int w, uint16_t *out, u32 cap, u32 *tab, int lane, u32 &head_words, u32 &tail_words);

__device__ bool coop_q_encode(const uint8_t *img, int h, int w, uint16_t *out, u32 cap, uint8_t *sym8, QCoopSmem &sm, u32 *tab, int lane,
                              u32 &head_words, u32 &tail_words) {
    uint16_t *sym = reinterpret_cast<uint16_t *>(sym8); /* cls | y << 8 per pixel, raster order */
    for (int k = lane; k < Q_CTX_ENTRIES; k += 32) { sm.ctx[k] = 0; tab[k] = 0; }
    __syncwarp();
    RowStage<3, 32> rows;
    rows.start(sm.stage, img, h, w);

    for (int i = 0; i < h; i++) { /* pass 1: model every pixel.  R: QNBLIC.c:586-623 */
        const uint8_t *row = img + (size_t)i * w;
        if (i < 2) {
            q_serial_row<false>(img, w, i, sm, lane, [&](int j, int cls, int px, int sign) {
                const int x = row[j];
                const int y = q_fold(x, px, sign);
                if (lane == 0) { sym[(size_t)i * w + j] = (uint16_t)(cls | (y << 8)); atomicAdd(&tab[cls * 256 + y], 1u); }
                return x;
            });
            continue;
        }
        int carry_px0 = 0;
        for (int j0 = 0; j0 < w; j0 += 32) {
            const bool active = j0 + lane < w;
            const int j = min(j0 + lane, w - 1);
            /* rows i, i-1, i-2 from shared memory, staged by cp.async one 32-pixel tile ahead.  At column 1 QNBLIC's shift
             * register still holds the old a = img[i-1][0] in e: exactly the halo cell r0[-1], so no lane is patched. */
            {
                const bool more = j0 + 32 < w;
                rows.advance(i, j0, more ? i : (i + 1 < h ? i + 1 : -1), more ? j0 + 32 : 0, lane, 32, lane == 0);
            }
            Nb nb;
            int x;
            sample_staged3(rows, j - j0, nb, x);
            const Pred pt = predictor_terms(nb);
            const int px0 = blend_prediction(pt, q_weight(pt.spread));
            int px0_left = __shfl_up_sync(FULL, px0, 1);
            if (lane == 0) px0_left = carry_px0;
            const int err_in = j == 0 ? 0 : nb.a - px0_left;
            const int cls = q_class(activity(nb, err_in));
            const int adr = active ? q_ctx_address(nb, px0, cls) : 0x10000 + lane;
            carry_px0 = __shfl_sync(FULL, px0, 31);

            /* bias table: same-address pixels in raster order, different addresses together */
            const unsigned peers = __match_any_sync(FULL, adr);
            const int my_turn = __popc(peers & ((1u << lane) - 1u));
            const int rounds = __reduce_max_sync(FULL, active ? __popc(peers) : 0);
            int y = 0;
            for (int r = 0; r < rounds; r++) {
                if (active && my_turn == r) {
                    const int c = sm.ctx[adr];
                    int px, sign;
                    q_bias_apply(c, px0, px, sign);
                    y = q_fold(x, px, sign);
                    sm.ctx[adr] = q_bias_learn(c, x - px0);
                }
                __syncwarp();
            }
            if (active) {
                atomicAdd(&tab[cls * 256 + y], 1u); /* result unused: a fire-and-forget reduction in L2 */
                sym[(size_t)i * w + j] = (uint16_t)(cls | (y << 8));
            }
        }
    }
    __syncwarp();
    return coop_q_finish(sym, h, w, out, cap, tab, lane, head_words, tail_words);
}

/* Per-stream symbol statistics in global memory: Q_TAB_ENTRIES counters (then freq | cumulative << 16), followed by the
 * Q_TAB_ENTRIES division magics of pass 2. */
constexpr int Q_TAB_STRIDE = 2 * Q_TAB_ENTRIES;

/* Exact floor(n / f) for every 32-bit n by one multiply-high (round-up method, Granlund & Montgomery): with
 * l = ceil(log2 f) and magic = floor(2^32 (2^l - f) / f) + 1,  t = mulhi(magic, n),  q = (t + ((n - t) >> min(l, 1))) >> max(l - 1, 0).
 * A power of two (f = 1 included) gives magic = 1, t = 0 and the plain shift. */
NB_DEV u32 q_div_magic(u32 f) {
    const int l = f > 1 ? 32 - __clz((int)(f - 1)) : 0;
    return (u32)(((u64)((1u << l) - f) << 32) / f) + 1u;
}

/* Second half of the encoder: header, the 12 histogram descriptions, and pass 2 (the reverse rANS sweep) over the
 * (class | y << 8) symbols of pass 1 and their counts in `tab`.  One warp; also the final stage of the
 * whole-GPU single-image pipeline (pipe_qnblic.cuh). */
__device__ bool coop_q_finish(const uint16_t *sym, int h, int w, uint16_t *out, u32 cap, u32 *tab, int lane, u32 &head_words, u32 &tail_words) {
    if (cap < 8) return false;
    if (lane == 0) { out[0] = 0x3051; out[1] = 0x322e; out[2] = (uint16_t)h; out[3] = (uint16_t)w; } /* R: QNBLIC.c:463-473 */
    const u32 o = q_finish_histograms(tab, out, 4, cap, lane);
    if (o >= cap) return false;
    u32 *magic = tab + Q_TAB_ENTRIES;
    for (int k = lane; k < Q_TAB_ENTRIES; k += 32) { const u32 f = tab[k] & 0xffffu; magic[k] = f ? q_div_magic(f) : 0u; }
    __syncwarp();

    /* pass 2: rANS, last pixel first, words written downwards from the end of the slot.  R: QNBLIC.c:238-253,635-650.
     * This sweep is the one serial chain of the encoder and its warp issues an instruction every ~5 cycles, so what
     * counts is the number of DEPENDENT operations per symbol: state / freq is one multiply-high and two shift-adds with
     * the entry's magic (exact: no correction steps), everything that does not depend on the state -- the block's table
     * entries (fetched one block ahead), shuffles (eight symbols ahead), shift counts -- sits off the chain. */
    u32 state = 1u << 16, p = cap;
    bool ok = true;
    const long long n = (long long)h * w;
    auto fetch = [&](long long base, u32 &e, u32 &m) { /* (freq | cumulative << 16, magic) of the block's 32 symbols */
        const long long idx = base + lane;
        e = 1u; m = 1u;
        if (base >= 0 && idx < n) {
            const u32 pair = (u32)sym[idx];
            const u32 at = (pair & 255u) * 256 + (pair >> 8);
            e = __ldcg(tab + at);
            m = __ldcg(magic + at);
        }
    };
    auto step = [&](u32 ej, u32 mj) {
        const u32 f = ej & 0xffffu, cum = ej >> 16;
        const int l = 32 - __clz((int)(f - 1)), sh1 = min(l, 1), sh2 = l - sh1; /* f >= 1; clz(0) = 32 gives l = 0 */
        u32 t = __umulhi(state, mj);
        u32 q = (t + ((state - t) >> sh1)) >> sh2;
        if (q > 0x1ffffu) {
            if (p > o) { p--; if (lane == 0) out[p] = (uint16_t)state; } else ok = false;
            state >>= 16;
            t = __umulhi(state, mj);
            q = (t + ((state - t) >> sh1)) >> sh2;
        }
        state = (state - q * f) + (q << Q_NORM_BITS) + cum;
    };
    u32 e, m, e_next, m_next;
    long long base = ((n - 1) / 32) * 32;
    fetch(base, e, m);
    for (; base >= 0; base -= 32) {
        fetch(base - 32, e_next, m_next); /* one block ahead: the L2 latency hides behind the 32 coder steps */
        const int last = (int)min(31ll, n - 1 - base);
        if (last == 31) {
            for (int g = 24; g >= 0; g -= 8) {
                u32 ee[8], mm[8];
#pragma unroll
                for (int k = 0; k < 8; k++) { ee[k] = __shfl_sync(FULL, e, g + 7 - k); mm[k] = __shfl_sync(FULL, m, g + 7 - k); }
#pragma unroll
                for (int k = 0; k < 8; k++) step(ee[k], mm[k]);
            }
        } else {
            for (int jj = last; jj >= 0; jj--) step(__shfl_sync(FULL, e, jj), __shfl_sync(FULL, m, jj));
        }
        e = e_next; m = m_next;
    }
    if (p >= o + 2) { p -= 2; if (lane == 0) { out[p + 1] = (uint16_t)state; out[p] = (uint16_t)(state >> 16); } } else ok = false;
    head_words = o; tail_words = cap - p;
    return ok;
}

/* 16-bit word reader over 128-byte lines (one 32-bit word per lane) */
struct WordReader {
    const uint16_t *base;
    u32 len, pos, word;
    unsigned long long line;
    int lane;
    NB_DEV void start(const uint16_t *p, u32 n_words, u32 at, int ln) { base = p; len = n_words; pos = at; word = 0; line = 0; lane = ln; }
    NB_DEV u32 get() {
        if (pos >= len) { pos++; return 0u; }
        const unsigned long long a = (unsigned long long)(base + pos);
        if ((a & ~127ull) != line) { line = a & ~127ull; word = *reinterpret_cast<const u32 *>(line + 4ull * (unsigned)lane); }
        const u32 wsel = __shfl_sync(FULL, word, (int)((a >> 2) & 31ull));
        pos++;
        return (a & 2ull) ? (wsel >> 16) : (wsel & 0xffffu);
    }
};

__device__ void coop_q_decode(const uint16_t *in, u32 avail, uint8_t *img, int h, int w, QDecSmem &sm, int lane) {
    for (int k = lane; k < Q_CTX_ENTRIES; k += 32) sm.ctx[k] = 0;
    u32 rd = 4;
    u32 *hist = reinterpret_cast<u32 *>(sm.rec); /* 256 words of scratch */
    for (int c = 0; c < Q_CLASSES; c++) { /* the 12 histogram descriptions are one sequential code stream.  R: QNBLIC.c:415-459 */
        if (lane == 0) {
            for (int k = 0; k < 256; k++) hist[k] = 0;
            u32 pos = 0, sum = 0;
#define Q_NEXT_() (rd < avail ? (u32)in[rd++] : (rd++, 0u))
#define Q_PUSH_(val) do { const u32 v_ = (val); if (pos < 256) { hist[pos] = v_; sum += v_; } pos++; } while (0)
            while (pos < 256 && sum < Q_NORM_SUM) {
                const u32 code = Q_NEXT_();
                if ((code >> 15) == 0) { Q_PUSH_(code); }
                else if ((code >> 14) == 2) { Q_PUSH_((code >> 7) & 0x7f); Q_PUSH_(code & 0x7f); }
                else if ((code >> 12) == 12) { Q_PUSH_((code >> 8) & 15); Q_PUSH_((code >> 4) & 15); Q_PUSH_(code & 15); }
                else if ((code >> 12) == 13) { Q_PUSH_((code >> 9) & 7); Q_PUSH_((code >> 6) & 7); Q_PUSH_((code >> 3) & 7); Q_PUSH_(code & 7); }
                else {
                    u32 run = (code & 0xff) + 4;
                    const u32 closer = (code >> 8) & 15, bit = (code >> 12) & 1;
                    while (run--) Q_PUSH_(bit);
                    if (closer != bit) Q_PUSH_(closer);
                }
            }
#undef Q_PUSH_
#undef Q_NEXT_
            u32 acc = 0; /* R: QNBLIC.c:290-295; clipped so a malformed table cannot leave 16 bits */
            for (int k = 0; k < 256; k++) { sm.cum[c * Q_CUM_STRIDE + k] = (uint16_t)min(acc, (u32)Q_NORM_SUM); acc += hist[k]; }
            sm.cum[c * Q_CUM_STRIDE + 256] = (uint16_t)min(acc, (u32)Q_NORM_SUM);
        }
        __syncwarp();
    }
    rd = __shfl_sync(FULL, rd, 0);
    __syncwarp();

    WordReader words;
    words.start(in, avail, rd, lane);
    u32 state = words.get() << 16; state |= words.get(); /* R: QNBLIC.c:256-260 */

    auto decode_symbol = [&](int cls) -> int { /* R: QNBLIC.c:262-274 with the LUT replaced by two ballots */
        const u32 slot = state & (Q_NORM_SUM - 1);
        const uint16_t *cum = sm.cum + cls * Q_CUM_STRIDE;
        const unsigned coarse = __ballot_sync(FULL, (u32)cum[8 * lane] <= slot);
        const int c8 = 8 * (__popc(coarse) - 1);
        const unsigned fine = __ballot_sync(FULL, lane < 8 && (u32)cum[c8 + (lane & 7)] <= slot);
        const int y = c8 + __popc(fine) - 1;
        const u32 base = cum[y], freq = (u32)cum[y + 1] - base;
        state = (state >> Q_NORM_BITS) * freq + slot - base;
        if (state < (1u << 16)) state = (state << 16) | words.get();
        return y;
    };

    RowStage<2> rows; /* rows i-1, i-2 of the decoder's own output */
    rows.start(sm.stage, img, h, w);
    for (int i = 0; i < h; i++) { /* R: QNBLIC.c:520-552 */
        uint8_t *row = img + (size_t)i * w;
        if (i < 2) {
            q_serial_row<true>(img, w, i, sm, lane, [&](int j, int cls, int px, int sign) {
                const int x = q_unfold(decode_symbol(cls), px, sign);
                if (lane == 0) row[j] = (uint8_t)x;
                __syncwarp();
                return x;
            });
            continue;
        }
        int err = 0, x1 = 0, x2 = 0;
        for (int j0 = 0; j0 < w; j0 += 32) {
            if ((j0 & (kStageTile - 1)) == 0) rows.advance(i, j0, j0 + kStageTile < w ? i : -1, j0 + kStageTile, lane, 32, lane == 0);
            {
                const int jr = min(j0 + lane, w - 1) - (j0 & ~(kStageTile - 1));
                sm.rec[lane] = make_pixrec_staged(rows.at(0, jr), rows.at(1, jr), 0u);
            }
            __syncwarp();
            const int n_here = min(32, w - j0);
            u32 my_x = 0;
            for (int jj = 0; jj < n_here; jj++) {
                const int j = j0 + jj;
                const uint4 ra = *reinterpret_cast<const uint4 *>(&sm.rec[jj]);
                const uint4 rb = *(reinterpret_cast<const uint4 *>(&sm.rec[jj]) + 1);
                Nb nb;
                nb.b = ra.x & 255; nb.c = (ra.x >> 8) & 255; nb.d = (ra.x >> 16) & 255; nb.f = ra.x >> 24;
                nb.g = ra.y & 255; nb.q = (ra.y >> 16) & 255;
                nb.a = j == 0 ? nb.b : x1;
                nb.e = j >= 2 ? x2 : (j == 1 ? nb.c : nb.a); /* column 1: the shift register still holds img[i-1][0] */
                const Pred pt = finish_predictor(nb, ra, rb);
                const int px0 = blend_prediction(pt, q_weight(pt.spread));
                const int act = abs(nb.a - nb.e) + abs(nb.a - nb.c) + (int)(ra.z >> 16) + 2 * abs(err);
                const int cls = q_class(act);
                const int adr = q_ctx_address(nb, px0, cls);
                const int c = sm.ctx[adr];
                int px, sign;
                q_bias_apply(c, px0, px, sign);
                const int x = q_unfold(decode_symbol(cls), px, sign);
                if (lane == jj) my_x = (u32)x;
                err = x - px0;
                if (lane == 0) sm.ctx[adr] = q_bias_learn(c, err);
                x2 = x1; x1 = x;
                __syncwarp();
            }
            if (lane < n_here) row[j0 + lane] = (uint8_t)my_x;
            __syncwarp();
        }
    }
}

} /* namespace nblic */
