/*
 * pipe_nblic.cuh -- lossless NBLIC effort-1 encode of ONE image by the WHOLE GPU: what the
 * reference's -t pipeline (R: QNBLIC.c:660-866: predictor workers :683-739, serial consumer :802-861) becomes for the
 * effort-1 coder, bit-exact with NBLICcodec (R: NBLIC.c:749-908).
 *
 * in a LOSSLESS encode every pixel is known, so the front end is a pure function of the input, and each
 * adaptive table is a set of independent chains -- one per entry -- whose only ordering constraint is raster order
 * among the pixels (decisions) that touch the same entry:
 *
 *   e1p_front_kernel     lane per pixel: neighbourhood, 7-direction predictor, activity -> soft class pair, texture
 *                        bits -> bias address                                                     R: NBLIC.c:287-410
 *   [stable partition by bias address: 2048 keys]
 *   e1p_bias_kernel      lane per bias-table entry, its pixels in raster order: apply, fold, learn  R: NBLIC.c:413-447
 *   [stable partition by rank-mapper key (px, sign): 512 keys, pixels with y < 20 only]
 *   e1p_rank_kernel      lane per key: symbol -> rank, frequency count, adjacent promotion        R: NBLIC.c:470-523
 *   e1p_decisions_kernel lane per pixel: the binary decisions of its adaptive-Golomb code, first counted (-> exclusive
 *                        scan = position of every decision in the stream), then emitted as two counter-node VISITS
 *                        each (main / side class)                                                 R: NBLIC.c:640-679
 *   [stable partition of the visits by counter node: 16 x 256 keys]
 *   e1p_node_kernel      warp per counter node, its visits in decision order: P(1), learn         R: NBLIC.c:589-617
 *   e1p_mix_kernel       lane per decision: mixed probability | bit << 12                         R: NBLIC.c:620-637
 *   e1p_coder_kernel     one warp per image: the range coder, the only serial stage               R: NBLIC.c:527-586
 *
 * The stable partition is four small kernels (psort_*): per-chunk key histogram, a scan over the chunks of every key
 * and one over the keys, and a scatter that walks each chunk in order, 32 items a step, ranking equal keys of a step
 * with __match_any.  It yields the PERMUTATION (sorted position -> item index) and the items' records in sorted order, so
 * a chain kernel streams its input sequentially (16-byte loads, one batch ahead of the chain) and only scatters results.
 *
 * What is left of the latency is the range coder (~40 cycles per decision, ~4.6 decisions per pixel): an image encodes
 * at about the speed of one CPU core instead of 6 times slower, and the images of a small batch overlap completely.
 */
#pragma once
#include "coop_nblic.cuh"

namespace nblic {

constexpr u32 kSortSkip = 0xffffffffu; /* key of an item that takes no part in the partition */

/* ---- stable partition by key -------------------------------------------------------------------- */

/* counts[key * n_chunks + chunk] = items of `chunk` with that key */
template <int NKEYS>
__global__ void __launch_bounds__(256) psort_count_kernel(const u32 *keys, long long n, int chunk_items, int n_chunks, u32 *counts) {
    __shared__ u32 hist[NKEYS];
    for (int k = threadIdx.x; k < NKEYS; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    const long long lo = (long long)blockIdx.x * chunk_items, hi = min(n, lo + chunk_items);
    for (long long p = lo + threadIdx.x; p < hi; p += blockDim.x) { const u32 k = keys[p]; if (k != kSortSkip) atomicAdd(&hist[k], 1u); }
    __syncthreads();
    for (int k = threadIdx.x; k < NKEYS; k += blockDim.x) counts[(size_t)k * n_chunks + blockIdx.x] = hist[k];
}

/* Scan, step 1: one warp per key turns the key's per-chunk counts (contiguous) into exclusive prefixes in place and
 * leaves the key's total in key_total[key]. */
__global__ void __launch_bounds__(256) psort_scan_keys_kernel(u32 *counts, int n_keys, int n_chunks, u32 *key_total) {
    const int lane = threadIdx.x & 31, key = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (key >= n_keys) return;
    u32 *row = counts + (size_t)key * n_chunks;
    u32 carry = 0;
    for (int base = 0; base < n_chunks; base += 32) {
        const int c = base + lane;
        const u32 v = c < n_chunks ? row[c] : 0u;
        u32 incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += o; }
        if (c < n_chunks) row[c] = carry + incl - v;
        carry += __shfl_sync(FULL, incl, 31);
    }
    if (lane == 0) key_total[key] = carry;
}
/* Scan, step 2: key_start[key] = items with a smaller key (exclusive scan of key_total, n_keys <= 4096), key_start[n_keys]
 * = items that took part.  One CTA of 1024 threads, four keys each. */
__global__ void __launch_bounds__(1024) psort_scan_totals_kernel(const u32 *key_total, int n_keys, u32 *key_start) {
    __shared__ u32 warp_sum[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    u32 v[4], mine = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) { const int key = 4 * threadIdx.x + k; v[k] = key < n_keys ? key_total[key] : 0u; mine += v[k]; }
    u32 incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        u32 s = warp_sum[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(FULL, s, d); if (lane >= d) s += o; }
        warp_sum[lane] = s;
    }
    __syncthreads();
    u32 run = (wid ? warp_sum[wid - 1] : 0u) + incl - mine;
#pragma unroll
    for (int k = 0; k < 4; k++) { const int key = 4 * threadIdx.x + k; if (key < n_keys) key_start[key] = run; run += v[k]; }
    if (threadIdx.x == 1023) key_start[n_keys] = warp_sum[31];
}

/* perm[slot] = item index and sorted[slot] = payload[item]; one warp per chunk, in item order inside the chunk => stable.
 * offsets = the per-key exclusive prefixes of step 1, key_start the key bases of step 2. */
template <int NKEYS, class T>
__global__ void __launch_bounds__(32) psort_scatter_kernel(const u32 *keys, const T *payload, long long n, int chunk_items, int n_chunks, const u32 *offsets,
                                                           const u32 *key_start, u32 *perm, T *sorted) {
    __shared__ u32 next[NKEYS];
    const int lane = threadIdx.x;
    for (int k = lane; k < NKEYS; k += 32) next[k] = key_start[k] + offsets[(size_t)k * n_chunks + blockIdx.x];
    __syncwarp();
    const long long lo = (long long)blockIdx.x * chunk_items, hi = min(n, lo + chunk_items);
    for (long long base = lo; base < hi; base += 32) {
        const long long p = base + lane;
        const u32 key = p < hi ? keys[p] : kSortSkip;
        const bool active = key != kSortSkip;
        const unsigned peers = __match_any_sync(FULL, active ? key : 0x80000000u + (u32)lane);
        const int before = __popc(peers & ((1u << lane) - 1u));
        if (active) { const u32 slot = next[key] + (u32)before; perm[slot] = (u32)p; sorted[slot] = payload[p]; }
        __syncwarp();
        if (active && before == __popc(peers) - 1) next[key] += (u32)__popc(peers); /* the last member of a group advances its cursor */
        __syncwarp();
    }
}

/* Exclusive scan of n u32 values (in place), total to *total_out.  One CTA; same scheme as psort_scan_kernel. */
__global__ void __launch_bounds__(1024) e1p_scan_kernel(u32 *vals, long long n, unsigned long long *total_out) {
    __shared__ unsigned long long part[1024];
    const long long per = (n + 1023) / 1024;
    const long long lo = min(n, per * threadIdx.x), hi = min(n, lo + per);
    unsigned long long s = 0;
    for (long long k = lo; k < hi; k++) s += vals[k];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        unsigned long long mine = 0;
        for (int k = 0; k < 32; k++) mine += part[threadIdx.x * 32 + k];
        unsigned long long incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xffffffffu, incl, d); if ((int)threadIdx.x >= d) incl += o; }
        unsigned long long run = incl - mine;
        for (int k = 0; k < 32; k++) { const unsigned long long v = part[threadIdx.x * 32 + k]; part[threadIdx.x * 32 + k] = run; run += v; }
    }
    __syncthreads();
    unsigned long long run = part[threadIdx.x];
    for (long long k = lo; k < hi; k++) { const u32 v = vals[k]; vals[k] = (u32)run; run += v; }
    if (threadIdx.x == 1023) *total_out = run;
}

/* ---- stage 1: front end ----------------------------------------------------------------------------- */

constexpr int kE1BiasKeys = N_CTX_ENTRIES;   /* 2048 bias-table entries          */
constexpr int kE1RankKeys = 512;             /* (px, sign)                       */
constexpr int kE1NodeKeys = N_FOREST_ENTRIES; /* 16 classes x 256 tree positions */

NB_DEV int e1p_px0(const uint8_t *img, int w, int i, int j, Nb &nb) {
    sample_positional(img, w, i, j, nb);
    const Pred pt = predictor_terms(nb);
    return blend_prediction(pt, n_weight(pt.spread));
}

/* rec[p] = px0 | x << 8 | soft << 16 (soft = u | v << 4 | wv << 8, 13 bits); key[p] = bias address */
__global__ void __launch_bounds__(256) e1p_front_kernel(const uint8_t *img, int h, int w, u32 *rec, u32 *key) {
    __shared__ uint16_t soft_tab[208];
    for (int d = threadIdx.x; d <= 200; d += blockDim.x) { int u, v, wv; n_soft_class(d, u, v, wv); soft_tab[d] = (uint16_t)(u | (v << 4) | (wv << 8)); }
    __syncthreads();
    const long long n = (long long)h * w;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(p / w), j = (int)(p - (long long)i * w);
        Nb nb, left;
        const int px0 = e1p_px0(img, w, i, j, nb);
        const int x = img[p];
        /* err of the pixel to the left: in a lossless stream its coded value is the pixel itself (= nb.a) */
        const int err_in = j == 0 ? 0 : clampi(nb.a - e1p_px0(img, w, i, j - 1, left), -127, 127);
        const u32 soft = soft_tab[min(activity(nb, err_in), 200)];
        rec[p] = (u32)px0 | ((u32)x << 8) | (soft << 16);
        key[p] = (((soft & 15u) >> 1) << 8) | (u32)texture_bits(nb, px0);
    }
}

/* ---- chain input: a lane walks [lo, hi) of two sorted arrays, eight items a batch, the next batch in flight ---- */
template <class T> struct ChainBatch { u32 idx[8]; T rec[8]; };
template <class T>
NB_DEV void chain_load(ChainBatch<T> &b, const u32 *perm, const T *sorted, u32 base, u32 hi) {
#pragma unroll
    for (int u = 0; u < 8; u++) { const bool in = base + u < hi; b.idx[u] = in ? __ldg(perm + base + u) : 0u; b.rec[u] = in ? __ldg(sorted + base + u) : T(0); }
}

/* ---- stage 2: bias chains ------------------------------------------------------------------------------ */
/* One lane per bias-table entry; out: yz[p] = y | px << 8 | sign << 16 | soft << 17; rank key[p] = (px << 1 | sign) or skip. */
__global__ void __launch_bounds__(128) e1p_bias_kernel(const u32 *perm, const u32 *sorted_rec, const u32 *key_start, u32 *yz, u32 *rank_key) {
    const int adr = blockIdx.x * blockDim.x + threadIdx.x;
    if (adr >= kE1BiasKeys) return;
    const u32 lo = key_start[adr], hi = key_start[adr + 1];
    int c = 0;
    ChainBatch<u32> cur, nxt;
    chain_load(cur, perm, sorted_rec, lo, hi);
    for (u32 base = lo; base < hi; base += 8) {
        chain_load(nxt, perm, sorted_rec, base + 8, hi);
        int before[8]; /* the table entry each pixel sees */
#pragma unroll
        for (int u = 0; u < 8; u++) { /* the chain proper: the update uses x - px0, not the corrected prediction */
            before[u] = c;
            if (base + u < hi) c = n_bias_learn(c, clampi((int)((cur.rec[u] >> 8) & 255u) - (int)(cur.rec[u] & 255u), -127, 127));
        }
#pragma unroll
        for (int u = 0; u < 8; u++) { /* independent of each other */
            if (base + u < hi) {
                const u32 r = cur.rec[u];
                const int px0 = (int)(r & 255u), x = (int)((r >> 8) & 255u);
                int px, sign;
                n_bias_apply(before[u], px0, px, sign);
                const int room = min(px, 255 - px), mag = abs(x - px);
                const int y = mag == 0 ? 0 : (mag <= room ? 2 * mag - ((x >= px) ^ sign) : mag + room);
                yz[cur.idx[u]] = (u32)y | ((u32)px << 8) | ((u32)sign << 16) | ((r >> 16) << 17);
                rank_key[cur.idx[u]] = y < N_RANKS ? (u32)((px << 1) | sign) : kSortSkip;
            }
        }
        cur = nxt;
    }
}

/* ---- stage 3: rank-mapper chains ------------------------------------------------------------------------ */
/* One lane per key, tables interleaved in shared memory ([entry][lane]: conflict free).  Rewrites the y field of yz[p]
 * with the rank z for the pixels that go through the mapper (sorted_yz = their yz words in chain order). */
__global__ void __launch_bounds__(64) e1p_rank_kernel(const u32 *perm, const u32 *sorted_yz, const u32 *key_start, u32 *yz) {
    __shared__ int cnt[N_RANKS][64];
    __shared__ uint8_t rank_of[N_RANKS][64], sym_at[N_RANKS][64];
    const int t = threadIdx.x, key = blockIdx.x * 64 + t;
    for (int r = 0; r < N_RANKS; r++) { cnt[r][t] = 2 * (N_RANKS - 1 - r); rank_of[r][t] = (uint8_t)r; sym_at[r][t] = (uint8_t)r; }
    if (key >= kE1RankKeys) return;
    const u32 lo = key_start[key], hi = key_start[key + 1];
    ChainBatch<u32> cur, nxt;
    chain_load(cur, perm, sorted_yz, lo, hi);
    for (u32 base = lo; base < hi; base += 8) {
        chain_load(nxt, perm, sorted_yz, base + 8, hi);
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (base + u < hi) {
                const u32 v = cur.rec[u];
                const int y = (int)(v & 255u);
                const int z = rank_of[y][t];
                const int cz = cnt[z][t] + 1;
                bool promoted = false;
                if (z > 0) {
                    const int cp = cnt[z - 1][t];
                    if (cp < cz) { /* one adjacent promotion (R: NBLIC.c:513-521) */
                        const int other = sym_at[z - 1][t];
                        cnt[z][t] = cp; cnt[z - 1][t] = cz;
                        sym_at[z][t] = (uint8_t)other; sym_at[z - 1][t] = (uint8_t)y;
                        rank_of[y][t] = (uint8_t)(z - 1); rank_of[other][t] = (uint8_t)z;
                        promoted = true;
                    }
                }
                if (!promoted) cnt[z][t] = cz;
                yz[cur.idx[u]] = (v & ~255u) | (u32)z;
            }
        }
        cur = nxt;
    }
}

/* ---- stage 4: decisions ----------------------------------------------------------------------------------- */
/* The decisions of one symbol in coding order (R: NBLIC.c:640-679, encoder side): emit(d, node_u, node_v, bit) with
 * node = class * 256 + tree position.  Returns their number, or -1 when the code escapes past the last order (no
 * valid 8-bit residual does). */
template <class Emit>
NB_DEV int e1p_walk(int k_step, int top, int u, int v, int z, Emit emit) {
    if (v / k_step != u / k_step) v = u;
    int node = 0, k, n = 0;
    for (;;) {
        k = u / k_step;
        const int bit = (node >> top) < (z >> k);
        emit(n++, u * 256 + node, v * 256 + node, bit);
        if (!bit) break;
        node += 1 << top;
        if (node >= 256) {
            node >>= 1; u = v = (k + 1) * k_step;
            if (u >= N_CLASSES) return -1;
        }
    }
    for (node++, k--; k >= 0; k--) {
        const int bit = (z >> k) & 1;
        emit(n++, u * 256 + node, v * 256 + node, bit);
        node += bit ? (1 << k) : 1;
    }
    return n;
}

/* doff[p] = decisions of the pixels before p inside p's block of kE1ScanBlock pixels; block_sums[b] = decisions of block
 * b; *bad is raised for an escape past the last order.  Thread t of a CTA owns 16 consecutive pixels. */
constexpr int kE1ScanBlock = 4096;
__global__ void __launch_bounds__(256) e1p_count_kernel(const u32 *yz, long long n, int k_step, u32 *doff, u32 *block_sums, int *bad) {
    __shared__ u32 warp_sum[8];
    const int top = (N_CLASSES - 1) / k_step, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long first = (long long)blockIdx.x * kE1ScanBlock + 16 * threadIdx.x;
    u32 cnt[16], mine = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const long long p = first + i;
        cnt[i] = 0;
        if (p < n) {
            const u32 v = yz[p];
            const int soft = (int)(v >> 17);
            const int d = e1p_walk(k_step, top, soft & 15, (soft >> 4) & 15, (int)(v & 255u), [](int, int, int, int) {});
            if (d < 0) atomicExch(bad, 1); else cnt[i] = (u32)d;
        }
        mine += cnt[i];
    }
    u32 incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const u32 o = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += o; }
    if (lane == 31) warp_sum[wid] = incl;
    __syncthreads();
    u32 run = incl - mine;
    for (int k = 0; k < wid; k++) run += warp_sum[k];
#pragma unroll
    for (int i = 0; i < 16; i++) { const long long p = first + i; if (p < n) doff[p] = run; run += cnt[i]; }
    if (threadIdx.x == 255) block_sums[blockIdx.x] = run;
}

/* visit 2g + role of decision g: vis_key = node (side class: skip when both classes share the node),
 * vis_rec = bit | weight << 1 | same << 7; dec_rec[g] = wv | bit << 5 */
__global__ void __launch_bounds__(256) e1p_emit_kernel(const u32 *yz, long long n, int k_step, const u32 *doff, const u32 *block_off, u32 *vis_key,
                                                       uint8_t *vis_rec, uint8_t *dec_rec) {
    const int top = (N_CLASSES - 1) / k_step;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const u32 v = yz[p];
        const int soft = (int)(v >> 17), wv = (soft >> 8) & 31;
        const size_t g0 = (size_t)doff[p] + block_off[p / kE1ScanBlock];
        e1p_walk(k_step, top, soft & 15, (soft >> 4) & 15, (int)(v & 255u), [&](int d, int nu, int nv, int bit) {
            const size_t g = g0 + (size_t)d;
            const bool same = nu == nv;
            vis_key[2 * g] = (u32)nu;
            vis_key[2 * g + 1] = same ? kSortSkip : (u32)nv;
            vis_rec[2 * g] = (uint8_t)(bit | ((N_MIX - wv) << 1) | (same ? 0x80 : 0)); /* weight 32 - wv <= 32 needs 6 bits */
            vis_rec[2 * g + 1] = (uint8_t)(bit | (wv << 1));
            dec_rec[g] = (uint8_t)(wv | (bit << 5));
        });
    }
}

/* ---- stage 5: counter-node chains ------------------------------------------------------------------------- */
/* One WARP per node (a few nodes -- the first unary decision of the common classes -- own a tenth of all visits each, so
 * the longest chain sets the stage's time): the lanes load 32 visits coalesced, every lane runs the 32-step counter chain
 * (uniform: six dependent operations a visit, the visits' records arrive by shuffles issued eight ahead), lane u keeps the
 * counter pair visit u sees, and the 32 probabilities floor(4096 n1 / (n0 + n1)) are then computed and stored side by side. */
__global__ void __launch_bounds__(128) e1p_node_kernel(const u32 *perm, const uint8_t *sorted_rec, const u32 *key_start, uint16_t *p1) {
    const int lane = threadIdx.x & 31, node = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (node >= kE1NodeKeys) return;
    const u32 lo = key_start[node], hi = key_start[node + 1];
    u32 c = (u32)N_MIX | ((u32)N_MIX << 16);
    for (u32 base = lo; base < hi; base += 32) {
        const bool mine = base + lane < hi;
        const u32 my_idx = mine ? __ldg(perm + base + lane) : 0u;
        const u32 my_rec = mine ? (u32)__ldg(sorted_rec + base + lane) : 0u; /* a record of 0 (weight 0) leaves the pair alone */
        u32 seen = c;
#pragma unroll
        for (int g = 0; g < 32; g += 8) {
            u32 r[8];
#pragma unroll
            for (int k = 0; k < 8; k++) r[k] = __shfl_sync(FULL, my_rec, g + k);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (lane == g + k) seen = c;
                const int bit = (int)(r[k] & 1u), weight = (int)((r[k] >> 1) & 63u);
                c = learn_packed(c, pair_sum(c), bit, weight);
                if (r[k] & 0x80u) c = learn_packed(c, pair_sum(c), bit, N_MIX - weight); /* both classes are this node: it learns the side weight too (R: NBLIC.c:633-636) */
            }
        }
        if (mine) {
            const uint16_t p = (uint16_t)node_p1_fast(seen, pair_sum(seen));
            p1[my_idx] = p;
            if (my_rec & 0x80u) p1[my_idx + 1] = p;
        }
    }
}

/* ---- stage 6: mix ------------------------------------------------------------------------------------------ */
__global__ void __launch_bounds__(256) e1p_mix_kernel(const uint16_t *p1, const uint8_t *dec_rec, unsigned long long n_dec, uint16_t *coded) {
    for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_dec; g += (unsigned long long)gridDim.x * blockDim.x) {
        const int wv = dec_rec[g] & 31, bit = dec_rec[g] >> 5;
        const int pu = p1[2 * g], pv = p1[2 * g + 1];
        coded[g] = (uint16_t)(max((pu * (N_MIX - wv) + pv * wv + N_MIX / 2) >> 5, 1) | (bit << 12));
    }
}

/* ---- stage 7: the range coder, one warp per image ------------------------------------------------------------ */
__device__ u32 e1p_code_stream(const uint16_t *coded, unsigned long long n_dec, int h, int w, int k_step, uint8_t *stream, u32 cap, int lane) {
    CoopCoder<false> rc;
    rc.out.start(stream, cap, lane);
    coop_put_header(rc, h, w, 0, k_step, 1);
    rc.start();
    u32 cur = lane < (long long)n_dec ? coded[lane] : 0u;
    for (unsigned long long base = 0; base < n_dec; base += 32) {
        const unsigned long long nb = base + 32 + lane;
        const u32 nxt = nb < n_dec ? coded[nb] : 0u; /* one block ahead */
        const int cnt = (int)min(32ull, n_dec - base);
        if (cnt == 32) { /* eight decisions' shuffles are issued ahead of the eight coder steps: a shuffle in front of every step
                          * (~25 cycles) was two thirds of the chain of this lone warp */
            for (int g = 0; g < 32; g += 8) {
                u32 cd[8];
#pragma unroll
                for (int k = 0; k < 8; k++) cd[k] = __shfl_sync(FULL, cur, g + k);
#pragma unroll
                for (int k = 0; k < 8; k++) rc.bit((int)(cd[k] >> 12), cd[k] & 0xfffu);
            }
        } else {
            for (int e = 0; e < cnt; e++) {
                const u32 cd = __shfl_sync(FULL, cur, e);
                rc.bit((int)(cd >> 12), cd & 0xfffu);
            }
        }
        cur = nxt;
    }
    rc.finish();
    return rc.out.overflow ? 0xffffffffu : rc.out.pos;
}

} /* namespace nblic */
