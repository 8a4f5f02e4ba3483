/*
 * pipe_qnblic.cuh -- QNBLIC ("Q0.2", effort 0) encode of ONE image by the WHOLE GPU: the bit-exact generalisation of the
 * reference's -t pipeline (R: QNBLIC.c:660-866: four predictor threads produce {x, px, adr} per pixel, R: :683-739, a
 * serial consumer applies the bias table, builds the histograms and runs the rANS pass, R: :802-861).
 *
 * in the lossless encoder the whole front end is a pure function of the input, the bias table is one
 * independent chain per ADDRESS (its update uses x - px0, not the corrected prediction), the histograms are an order-free
 * reduction, and only the rANS sweep is serial.  So:
 *
 *   qpipe_front_kernel    one lane per pixel: neighbourhood, 7-direction predictor, activity class, bias address
 *                         (rows 0 and 1 keep QNBLIC's literal shift register, one lane per row).  R: QNBLIC.c:586-605
 *   qpipe_count_kernel    per chunk of pixels: how many pixels hit each of the 3072 addresses
 *   qpipe_scan_kernel     exclusive scan of those counts over (address, chunk): where every chunk writes every address
 *   qpipe_scatter_kernel  STABLE partition by address: a chunk is walked in raster order, 32 pixels at a time, and
 *                         __match_any ranks the pixels that share an address inside a step
 *   qpipe_chain_kernel    one lane per address: its pixels in raster order through bias apply / fold / learn
 *                         (R: QNBLIC.c:607-619), symbols scattered back to raster order, symbol counts by atomics
 *   qpipe_finish_kernel   one warp per image: histogram normalisation + descriptions + the reverse rANS sweep
 *                         (coop_qnblic.cuh: coop_q_finish).  R: QNBLIC.c:625-650
 *
 * The bytes are those of coop_q_encode / QNBLICcompress.  The serial sweep (one symbol per ~25 cycles) is what is left
 * of the latency: Kodak-size images encode in about half the time of one CPU core instead of four times longer.
 */
#pragma once
#include "coop_qnblic.cuh"

namespace nblic {

constexpr int kPipeKeys = Q_CTX_ENTRIES; /* 3072 bias-table addresses */

/* per-pixel record of the front end: adr (12 bits) | px0 << 12 | x << 20 */
NB_DEV u32 qpipe_pack(int adr, int px0, int x) { return (u32)adr | ((u32)px0 << 12) | ((u32)x << 20); }

/* px0 of pixel (i, j), i >= 2, positional sampling with the e-at-column-1 exception (coop_qnblic.cuh) */
NB_DEV int qpipe_px0(const uint8_t *img, int w, int i, int j, Nb &nb) {
    sample_positional(img, w, i, j, nb);
    if (j == 1) nb.e = img[(size_t)(i - 1) * w];
    const Pred pt = predictor_terms(nb);
    return blend_prediction(pt, q_weight(pt.spread));
}

__global__ void __launch_bounds__(256) qpipe_front_kernel(const uint8_t *img, int h, int w, u32 *meta) {
    const long long n = (long long)h * w;
    if (blockIdx.x == 0) { /* rows 0 and 1: the reference loop, one lane per row (R: QNBLIC.c:67-79) */
        const int i = threadIdx.x;
        if (i < 2 && i < h) {
            Nb nb;
            int err = 0;
            sample_positional(img, w, i, 0, nb);
            for (int j = 0; j < w; j++) {
                const int x = img[(size_t)i * w + j];
                const Pred pt = predictor_terms(nb);
                const int px0 = blend_prediction(pt, q_weight(pt.spread));
                const int cls = q_class(activity(nb, err));
                meta[(size_t)i * w + j] = qpipe_pack(q_ctx_address(nb, px0, cls), px0, x);
                err = x - px0;
                q_window_shift(img, w, i, j, x, nb);
            }
        }
        return;
    }
    const long long first = 2ll * w;
    for (long long p = first + (long long)(blockIdx.x - 1) * blockDim.x + threadIdx.x; p < n; p += (long long)(gridDim.x - 1) * blockDim.x) {
        const int i = (int)(p / w), j = (int)(p - (long long)i * w);
        Nb nb, left;
        const int px0 = qpipe_px0(img, w, i, j, nb);
        const int x = img[p];
        const int err = j == 0 ? 0 : (int)img[p - 1] - qpipe_px0(img, w, i, j - 1, left);
        const int cls = q_class(activity(nb, err));
        meta[p] = qpipe_pack(q_ctx_address(nb, px0, cls), px0, x);
    }
}

/* counts[key * n_chunks + chunk] = pixels of `chunk` with bias address `key` */
__global__ void __launch_bounds__(256) qpipe_count_kernel(const u32 *meta, long long n, int chunk_px, int n_chunks, u32 *counts) {
    __shared__ u32 hist[kPipeKeys];
    for (int k = threadIdx.x; k < kPipeKeys; k += blockDim.x) hist[k] = 0;
    __syncthreads();
    const long long lo = (long long)blockIdx.x * chunk_px, hi = min(n, lo + chunk_px);
    for (long long p = lo + threadIdx.x; p < hi; p += blockDim.x) atomicAdd(&hist[meta[p] & 0xfffu], 1u);
    __syncthreads();
    for (int k = threadIdx.x; k < kPipeKeys; k += blockDim.x) counts[(size_t)k * n_chunks + blockIdx.x] = hist[k];
}

/* In place: counts -> exclusive prefix in (key-major, chunk-minor) order; key_start[key] = first slot of the key, key_start[3072] = n.
 * One CTA of 1024 threads; thread t owns the contiguous piece [t * per, (t + 1) * per) of the flattened array. */
__global__ void __launch_bounds__(1024) qpipe_scan_kernel(u32 *counts, int n_chunks, u32 *key_start) {
    __shared__ unsigned long long part[1024];
    const size_t total = (size_t)kPipeKeys * n_chunks;
    const size_t per = (total + 1023) / 1024;
    const size_t lo = min(total, per * threadIdx.x), hi = min(total, lo + per);
    unsigned long long s = 0;
    for (size_t k = lo; k < hi; k++) s += counts[k];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < 32) { /* scan of the 1024 partial sums by one warp: 32 per lane, then a shuffle scan */
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
    for (size_t k = lo; k < hi; k++) { const u32 v = counts[k]; counts[k] = (u32)run; run += v; }
    if (threadIdx.x == 1023) key_start[kPipeKeys] = (u32)run; /* every entry lies in front of the last piece's end */
    __syncthreads();
    for (int key = threadIdx.x; key < kPipeKeys; key += 1024) key_start[key] = counts[(size_t)key * n_chunks];
}

/* sorted[slot] = (pixel index, meta); one warp per chunk, raster order inside the chunk => stable */
__global__ void __launch_bounds__(32) qpipe_scatter_kernel(const u32 *meta, long long n, int chunk_px, int n_chunks, const u32 *offsets, uint2 *sorted) {
    __shared__ u32 next[kPipeKeys];
    const int lane = threadIdx.x;
    for (int k = lane; k < kPipeKeys; k += 32) next[k] = offsets[(size_t)k * n_chunks + blockIdx.x];
    __syncwarp();
    const long long lo = (long long)blockIdx.x * chunk_px, hi = min(n, lo + chunk_px);
    for (long long base = lo; base < hi; base += 32) {
        const long long p = base + lane;
        const bool active = p < hi;
        const u32 m = active ? meta[p] : 0u;
        const int key = active ? (int)(m & 0xfffu) : 0x10000 + lane;
        const unsigned peers = __match_any_sync(FULL, key);
        const int before = __popc(peers & ((1u << lane) - 1u));
        if (active) {
            sorted[next[key] + (u32)before] = make_uint2((u32)p, m);
        }
        __syncwarp();
        if (active && before == __popc(peers) - 1) next[key] += (u32)__popc(peers); /* the last member of a group advances its cursor */
        __syncwarp();
    }
}

/* one lane per bias-table address: R: QNBLIC.c:176-217 on the address's pixels in raster order */
__global__ void __launch_bounds__(128) qpipe_chain_kernel(const uint2 *sorted, const u32 *key_start, uint16_t *sym, u32 *tab) {
    const int key = blockIdx.x * blockDim.x + threadIdx.x;
    if (key >= kPipeKeys) return;
    const u32 lo = key_start[key], hi = key_start[key + 1];
    const int cls = key >> 8;
    int c = 0;
    for (u32 base = lo; base < hi; base += 8) { /* eight independent loads in flight, then the chain */
        uint2 e[8];
#pragma unroll
        for (int u = 0; u < 8; u++) e[u] = base + u < hi ? __ldg(sorted + base + u) : make_uint2(0u, 0u);
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (base + u < hi) {
                const int px0 = (int)((e[u].y >> 12) & 255u), x = (int)(e[u].y >> 20);
                int px, sign;
                q_bias_apply(c, px0, px, sign);
                const int y = q_fold(x, px, sign);
                c = q_bias_learn(c, x - px0);
                sym[e[u].x] = (uint16_t)(cls | (y << 8));
                atomicAdd(&tab[cls * 256 + y], 1u);
            }
        }
    }
}

} /* namespace nblic */
