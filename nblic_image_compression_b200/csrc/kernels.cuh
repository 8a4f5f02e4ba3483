/*
 * kernels.cuh -- every __global__ entry point of libnblic_b200.so and the task record they share with the
 * host side (stream_kernels.cu).  Single translation unit: included once, by stream_kernels.cu.
 *
 *   coop_nblic_kernel<NAVP, MODE, RG>  warp-cooperative NBLIC, efforts 1-3 (coop_nblic.cuh, coop_avp.cuh)
 *   coop_q_kernel<DEC>                 warp-cooperative QNBLIC, effort 0 (coop_qnblic.cuh)
 *   coder_kernel<KIND, DEC, MAP>       sequential formulation, one agent per stream (codec_core.cuh); test build only (NBLIC_B200_SEQUENTIAL)
 *   scan_lengths_kernel / gather_streams_kernel   compaction of the variable-length streams
 *   peek_headers_kernel, synth_gray_kernel, divcheck_kernel
 */
#pragma once
#include <cuda_runtime.h>

#include <type_traits>

#include "../../include/nblic_b200.h"
#include "codec_core.cuh"
#include "coop_nblic.cuh"
#include "coop_qnblic.cuh"

using namespace nblic;

namespace {

enum { KIND_Q = 0, KIND_N = 1 };
enum { MAP_WARP = 1, MAP_LANE = 2 };

/* one entry per image of the batch, device resident */
struct Task {
    const uint8_t *src;  /* pixels in (encode)                                   */
    uint8_t *rec;        /* reconstruction (near > 0 encode) / decoded raster    */
    uint8_t *slot;       /* encode: private worst-case output slot; decode: stream bytes */
    uint8_t *sym;        /* QNBLIC encode scratch (2 B / pixel)                   */
    u32 slot_cap;        /* bytes (encode: capacity, decode: valid)               */
    u32 head_len;        /* out: bytes at the start of the slot                   */
    u32 tail_len;        /* out: bytes at the end of the slot (QNBLIC rANS words) */
    int status;          /* out                                                   */
    int h, w, near, k_step, effort;
};

} /* namespace */
#include "subwarp_nblic.cuh" /* needs Task */
#include "pipe_qnblic.cuh"
#include "pipe_nblic.cuh"
namespace {

constexpr int kChunkImagesPerSm = 192; /* host-buffer API: images per pipeline chunk and SM (multiple of 8, 16 and 24) */
constexpr int N_STATE_BYTES = N_CTX_ENTRIES * 2 + N_FOREST_ENTRIES * 4 + N_RANK_ENTRIES * 2 + N_RANK_ENTRIES * 4; /* 81920 */
constexpr int N_SMEM_BYTES = N_CTX_ENTRIES * 2 + N_FOREST_ENTRIES * 4 + N_RANK_ENTRIES * 2;                       /* 40960 */
constexpr int N_COUNT_BYTES = N_RANK_ENTRIES * 4;
constexpr int Q_STATE_BYTES = Q_CTX_ENTRIES * 4 + Q_TAB_ENTRIES * 4; /* 24576 */

#ifdef NBLIC_B200_SEQUENTIAL /* the sequential formulation ships only in the test build (libnblic_b200_seq.so), as the parity tests' second opinion */
__device__ __forceinline__ NState carve_nstate(uint8_t *hot, uint8_t *counts, i64 *avp, size_t avp_half) {
    NState s;
    s.forest = reinterpret_cast<u32 *>(hot);
    s.ctx = reinterpret_cast<int16_t *>(hot + N_FOREST_ENTRIES * 4);
    s.rank_of = hot + N_FOREST_ENTRIES * 4 + N_CTX_ENTRIES * 2;
    s.sym_at = s.rank_of + N_RANK_ENTRIES;
    s.count = reinterpret_cast<int *>(counts);
    s.Brow = avp;
    s.Frow = avp ? avp + avp_half : nullptr;
    return s;
}

template <bool DEC>
__device__ void run_nblic(Task &t, const NState &st, int lane, int nl) {
    nstate_reset(st, lane, nl);
    const int n = t.effort == 1 ? 0 : (t.effort == 2 ? 6 : 10);
    if (n > 0) {
        const size_t cells = (size_t)t.w * (1 + n + n * n);
        for (size_t k = lane; k < cells; k += nl) st.Brow[k] = 0;
    }
    if (nl > 1) __syncwarp();
    if (lane == 0) {
        NJob job;
        job.src = t.src; job.rec = t.rec; job.stream = t.slot; job.stream_cap = t.slot_cap;
        job.h = t.h; job.w = t.w; job.near = t.near; job.k_step = t.k_step;
        u32 len;
        if (n == 0) len = nblic_stream<0, DEC>(job, st);
        else if (n == 6) len = nblic_stream<6, DEC>(job, st);
        else len = nblic_stream<10, DEC>(job, st);
        if (!DEC) {
            if (len == 0xffffffffu) { t.status = NBLIC_B200_OVERFLOW; t.head_len = 0; }
            else t.head_len = len;
            t.tail_len = 0;
        } else if (len != 0) t.status = NBLIC_B200_CORRUPT;
    }
    if (nl > 1) __syncwarp();
}

template <bool DEC>
__device__ void run_qnblic(Task &t, const QState &st, int lane, int nl) {
    if (nl > 1) __syncwarp();
    if (lane == 0) {
        QJob job;
        job.src = t.src; job.rec = t.rec; job.stream = reinterpret_cast<uint16_t *>(t.slot);
        job.stream_cap_words = t.slot_cap / 2; job.sym = t.sym; job.h = t.h; job.w = t.w;
        if (DEC) qnblic_decode_stream(job, st);
        else {
            u32 head = 0, tail = 0;
            if (qnblic_encode_stream(job, st, head, tail)) { t.head_len = head * 2; t.tail_len = tail * 2; }
            else { t.status = NBLIC_B200_OVERFLOW; t.head_len = t.tail_len = 0; }
        }
    }
    if (nl > 1) __syncwarp();
}

/*
 * Persistent coder kernel.  `order` lists the task indices of this launch, largest image first;
 * `queue` is the shared cursor.  MAP_WARP: blockDim = 32, dynamic shared memory holds the hot
 * adaptive state of the warp's current stream; `cold` holds per-slot state that does not fit
 * (rank-mapper frequencies).  MAP_LANE: every thread is a slot and all state is in `cold`.
 */
template <int KIND, bool DEC, int MAP>
__global__ void __launch_bounds__(32) coder_kernel(Task *tasks, const int *order, int n_order, int *queue, uint8_t *cold,
                                                   size_t cold_stride, i64 *avp, size_t avp_stride) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31;
    const size_t slot = MAP == MAP_WARP ? blockIdx.x : (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint8_t *my_cold = cold + slot * cold_stride;
    i64 *my_avp = avp ? avp + slot * avp_stride : nullptr;
    for (;;) {
        int pos;
        if (MAP == MAP_WARP) {
            pos = lane == 0 ? atomicAdd(queue, 1) : 0;
            pos = __shfl_sync(0xffffffffu, pos, 0);
        } else {
            pos = atomicAdd(queue, 1);
        }
        if (pos >= n_order) break;
        Task &t = tasks[order[pos]];
        if (KIND == KIND_N) {
            const NState st = MAP == MAP_WARP ? carve_nstate(smem, my_cold, my_avp, avp_stride / 2)
                                              : carve_nstate(my_cold, my_cold + N_SMEM_BYTES, my_avp, avp_stride / 2);
            run_nblic<DEC>(t, st, MAP == MAP_WARP ? lane : 0, MAP == MAP_WARP ? 32 : 1);
        } else {
            QState st;
            uint8_t *base = MAP == MAP_WARP ? smem : my_cold;
            st.ctx = reinterpret_cast<int *>(base);
            st.tab = reinterpret_cast<u32 *>(base + Q_CTX_ENTRIES * 4);
            run_qnblic<DEC>(t, st, MAP == MAP_WARP ? lane : 0, MAP == MAP_WARP ? 32 : 1);
        }
    }
}

#endif /* NBLIC_B200_SEQUENTIAL */

/* Warp-cooperative NBLIC kernels (coop_nblic.cuh, coop_avp.cuh): one warp per CTA, adaptive state in
 * shared memory, rank-mapper frequencies in `counts` (one [512][20] int table per CTA), AVP column
 * accumulators in `avp` (efforts 2 / 3: 2 * avp_half int64 per CTA).
 * MODE 0: lossless effort-1 encode (phase P = whole front end); 1: encode with a per-pixel front end
 * (near-lossless, and every effort-2/3 encode); 2: decode.  NAVP = 0 / 6 / 10 for effort 1 / 2 / 3. */
/* Two independent one-warp streams share a CTA (kCoopWarps): shared memory is granted in 256-byte steps plus 1 KB per
 * CTA, and a pair of streams wastes half of that -- 24 resident lossless effort-1 encoders per SM with the staged row
 * tiles in place (22 as one-warp CTAs), 22 decoders (20).  The warps never synchronise with each other. */
constexpr int kCoopWarps = 2;
template <int NAVP, int MODE, bool RG>
__global__ void __launch_bounds__(32 * kCoopWarps) coop_nblic_kernel(Task *tasks, const int *order, int n_order, int *queue, int *counts, i64 *avp,
                                                                     size_t avp_half, u32 smem_per_warp) {
    extern __shared__ __align__(16) uint8_t smem_all[];
    using L = CoopLayout<NAVP, MODE, RG>;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const size_t slot = (size_t)blockIdx.x * kCoopWarps + wid; /* this warp's index among the resident streams */
    uint8_t *smem = smem_all + (size_t)wid * smem_per_warp;
    /* per-stream global scratch: [512][20] int frequencies, then (efforts 2/3) the [512][20] byte rank tables */
    uint8_t *my_scratch = reinterpret_cast<uint8_t *>(counts) + slot * (N_RANK_ENTRIES * 5);
    int *my_counts = reinterpret_cast<int *>(my_scratch);
    CoopSmem &sm = *reinterpret_cast<CoopSmem *>(smem);
    PixRec *recs = reinterpret_cast<PixRec *>(smem + L::kRecOff);
    uint8_t *stage = smem + L::kStageOff;
    AvpSmem *asm_ = NAVP > 0 ? reinterpret_cast<AvpSmem *>(smem + L::kAvpOff) : nullptr;
    uint8_t *rank = L::kRankGlobal ? my_scratch + N_RANK_ENTRIES * 4 : smem + L::kRankOff;
    u32 *forest = reinterpret_cast<u32 *>(smem + L::kForestOff); /* sized by the host for the largest k_step of the launch */
    i64 *my_b = NAVP > 0 ? avp + slot * 2 * avp_half : nullptr;
    for (;;) {
        int pos = lane == 0 ? atomicAdd(queue, 1) : 0;
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (pos >= n_order) break;
        Task &t = tasks[order[pos]];
        u32 len;
        if constexpr (MODE == 0) len = coop_e1_encode_lossless<RG>(t.src, t.h, t.w, t.slot, t.slot_cap, sm, stage, rank, forest, my_counts, lane);
        else {
            const bool lossless_enc = MODE == 1 && t.near == 0; /* neighbours are the source pixels themselves */
            len = coop_feedback<NAVP, MODE == 2, RG>(t.src, lossless_enc ? t.src : t.rec, lossless_enc ? nullptr : t.rec, t.h, t.w, t.near, t.k_step,
                                                 t.slot, t.slot_cap, sm, recs, stage, asm_, rank, forest, my_b, my_b ? my_b + avp_half : nullptr, my_counts,
                                                 lane);
        }
        if (lane == 0) {
            if (MODE == 2) { if (len != 0) t.status = NBLIC_B200_CORRUPT; }
            else {
                if (len == 0xffffffffu) { t.status = NBLIC_B200_OVERFLOW; t.head_len = 0; }
                else t.head_len = len;
                t.tail_len = 0;
            }
        }
        __syncwarp();
    }
}

/* Effort-1 decoder with 32 / LPS streams per warp (subwarp_nblic.cuh).  `packs`: 32 / LPS task indices per entry
 * (-1 = empty), all of one height x width; `scratch`: per CTA 32 / LPS x kSubScratchBytes. */
template <int LPS, bool FORESTG>
__global__ void __launch_bounds__(32) subwarp_decode_kernel(Task *tasks, const int *packs, int n_packs, int *queue, uint8_t *scratch, int max_nodes) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x;
    uint8_t *mine = scratch + (size_t)blockIdx.x * (32 / LPS) * kSubScratchBytes;
    for (;;) {
        int pos = lane == 0 ? atomicAdd(queue, 1) : 0;
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (pos >= n_packs) break;
        subwarp_decode_pack<LPS, FORESTG>(tasks, packs + (size_t)pos * (32 / LPS), smem, mine, max_nodes, lane);
        __syncwarp();
    }
}

/* Warp-cooperative QNBLIC kernel (coop_qnblic.cuh): one warp per CTA, bias table and the 12 histograms in
 * shared memory. */
template <bool DEC>
__global__ void __launch_bounds__(32) coop_q_kernel(Task *tasks, const int *order, int n_order, int *queue, u32 *tabs) {
    __shared__ typename std::conditional<DEC, QDecSmem, QCoopSmem>::type sm;
    const int lane = threadIdx.x;
    for (;;) {
        int pos = lane == 0 ? atomicAdd(queue, 1) : 0;
        pos = __shfl_sync(0xffffffffu, pos, 0);
        if (pos >= n_order) break;
        Task &t = tasks[order[pos]];
        if constexpr (DEC) coop_q_decode(reinterpret_cast<const uint16_t *>(t.slot), t.slot_cap / 2, t.rec, t.h, t.w, sm, lane);
        else {
            u32 head = 0, tail = 0;
            const bool ok = coop_q_encode(t.src, t.h, t.w, reinterpret_cast<uint16_t *>(t.slot), t.slot_cap / 2, t.sym, sm,
                                          tabs + (size_t)blockIdx.x * Q_TAB_STRIDE, lane, head, tail);
            if (lane == 0) {
                if (ok) { t.head_len = head * 2; t.tail_len = tail * 2; }
                else { t.status = NBLIC_B200_OVERFLOW; t.head_len = t.tail_len = 0; }
            }
        }
        __syncwarp();
    }
}

/* Last stage of the whole-GPU QNBLIC encode (pipe_qnblic.cuh): one warp per image finishes the stream from the
 * symbols in t.sym and the counts in tabs[image]. */
__global__ void __launch_bounds__(32) qpipe_finish_kernel(Task *tasks, const int *order, int n_order, u32 *tabs) {
    const int lane = threadIdx.x;
    if ((int)blockIdx.x >= n_order) return;
    Task &t = tasks[order[blockIdx.x]];
    u32 head = 0, tail = 0;
    const bool ok = coop_q_finish(reinterpret_cast<const uint16_t *>(t.sym), t.h, t.w, reinterpret_cast<uint16_t *>(t.slot), t.slot_cap / 2,
                                  tabs + (size_t)blockIdx.x * Q_TAB_STRIDE, lane, head, tail);
    if (lane == 0) {
        if (ok) { t.head_len = head * 2; t.tail_len = tail * 2; }
        else { t.status = NBLIC_B200_OVERFLOW; t.head_len = t.tail_len = 0; }
    }
}

/* Last stage of the whole-GPU effort-1 encode (pipe_nblic.cuh): one warp codes the image's decisions. */
__global__ void __launch_bounds__(32) e1p_coder_kernel(Task *tasks, int task_index, const uint16_t *coded, const unsigned long long *n_dec, const int *bad) {
    Task &t = tasks[task_index];
    const u32 len = *bad ? 0xffffffffu : e1p_code_stream(coded, *n_dec, t.h, t.w, t.k_step, t.slot, t.slot_cap, threadIdx.x);
    if (threadIdx.x == 0) {
        if (len == 0xffffffffu) { t.status = NBLIC_B200_OVERFLOW; t.head_len = 0; }
        else t.head_len = len;
        t.tail_len = 0;
    }
}

/* test hook: the reciprocal-based exact division of coop_avp.cuh on arbitrary operands */
__global__ void divcheck_kernel(const i64 *num, const i64 *den, int n, i64 *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = den[i] == 0 ? 0 : div_rcp(num[i], make_rcp(den[i]));
}

/* Exclusive scan of head_len + tail_len over the batch; one CTA of 1024 threads, warp shuffles. */
__global__ void __launch_bounds__(1024) scan_lengths_kernel(const Task *tasks, int n, unsigned long long *offsets) {
    __shared__ unsigned long long warp_sum[32];
    __shared__ unsigned long long carry;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned long long len = i < n ? (unsigned long long)tasks[i].head_len + tasks[i].tail_len : 0ull;
        unsigned long long v = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xffffffffu, v, d); if (lane >= d) v += o; }
        if (lane == 31) warp_sum[wid] = v;
        __syncthreads();
        if (wid == 0) {
            unsigned long long s = warp_sum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xffffffffu, s, d); if (lane >= d) s += o; }
            warp_sum[lane] = s;
        }
        __syncthreads();
        const unsigned long long before = carry + (wid ? warp_sum[wid - 1] : 0ull) + v - len;
        if (i < n) offsets[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + len;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[n] = carry;
}

/* Pack the streams: CTA (image, piece) copies up to `chunk` bytes of image's stream.  The head comes
 * from the start of the slot, the tail from its end.  The image index rides gridDim.x (no 65535 limit). */
__global__ void __launch_bounds__(256) gather_streams_kernel(const Task *tasks, const unsigned long long *offsets, uint8_t *out,
                                                             unsigned long long out_cap, u32 chunk, int *overflow) {
    const Task &t = tasks[blockIdx.x];
    const u32 total = t.head_len + t.tail_len;
    const u32 begin = blockIdx.y * chunk;
    if (begin >= total) return;
    const u32 end = min(total, begin + chunk);
    const unsigned long long dst0 = offsets[blockIdx.x];
    if (dst0 + total > out_cap) { if (threadIdx.x == 0 && blockIdx.y == 0) atomicExch(overflow, 1); return; }
    const uint8_t *tail = t.slot + (t.slot_cap & ~1u) - t.tail_len;
    for (u32 k = begin + threadIdx.x; k < end; k += blockDim.x)
        out[dst0 + k] = k < t.head_len ? t.slot[k] : tail[k - t.head_len];
}

struct Peek { int h, w, near, k_step, effort, ok; };

__host__ __device__ inline Peek peek_bytes(const uint8_t *p, size_t len) { /* R: NBLIC.c:698-745, QNBLIC.c:475-486 */
    Peek r = {0, 0, 0, 0, 0, 0};
    if (len >= 8 && p[0] == 0x51 && p[1] == 0x30 && p[2] == 0x2e && p[3] == 0x32) { /* "Q0.2" as LE words */
        r.h = p[4] | (p[5] << 8); r.w = p[6] | (p[7] << 8); r.effort = 0;
        r.ok = r.h > 0 && r.w > 0 && (long long)r.h * r.w <= 100000000LL;
        return r;
    }
    const char magic[9] = "NBLIC0.3";
    if (len < 16) return r;
    for (int k = 0; k < 8; k++) if (p[k] != (uint8_t)magic[k]) return r;
    const int channels = p[8];
    r.h = (p[9] << 8) | p[10]; r.w = (p[11] << 8) | p[12]; r.near = p[13]; r.k_step = p[14]; r.effort = p[15];
    r.ok = r.h > 0 && r.w > 0 && (long long)r.h * r.w <= 100000000LL && channels <= 1 && r.near <= 9 && r.k_step >= 3 && r.k_step <= 16 &&
           r.effort >= 1 && r.effort <= 3;
    return r;
}

__global__ void peek_headers_kernel(const uint8_t *streams, const unsigned long long *starts, const unsigned long long *lens, int n, Peek *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = peek_bytes(streams + starts[i], (size_t)lens[i]);
}

/* ---- synthetic photographic-like generator (nblic_image_compression_b200/synth.py) ---- */
__device__ __forceinline__ u32 h32(u32 v) { v ^= v >> 16; v *= 0x7feb352du; v ^= v >> 15; v *= 0x846ca68bu; v ^= v >> 16; return v; }
__device__ __forceinline__ int lattice(u32 key, int ix, int iy) { return (int)(h32(((u32)ix * 0x9E3779B1u) ^ h32(((u32)iy * 0x85EBCA77u) ^ key)) & 255u); }

struct Occluders { int v[12][5]; };

/* one image (gridDim.y == 1, occluders by value) or a batch of equal-size images (blockIdx.y = image, seed + blockIdx.y * seed_stride,
 * occluder tables in `occ_batch`, rasters packed back to back) */
__global__ void __launch_bounds__(256) synth_gray_kernel(uint8_t *out, int h, int w, u32 seed, u32 seed_stride, Occluders occ, const Occluders *occ_batch) {
    const long long n = (long long)h * w;
    if (occ_batch) { occ = occ_batch[blockIdx.y]; seed += blockIdx.y * seed_stride; out += (size_t)blockIdx.y * (size_t)n; }
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(p / w), x = (int)(p % w);
        const int shifts[7] = {8, 7, 6, 5, 4, 3, 2}, amps[7] = {64, 48, 32, 20, 12, 7, 4};
        int acc = 0;
#pragma unroll
        for (int o = 0; o < 7; o++) {
            const int sh = shifts[o], S = 1 << sh;
            const u32 key = seed * 131u + (u32)o * 7919u;
            const int ix = x >> sh, iy = y >> sh, fx = x & (S - 1), fy = y & (S - 1);
            const int v00 = lattice(key, ix, iy), v10 = lattice(key, ix + 1, iy), v01 = lattice(key, ix, iy + 1), v11 = lattice(key, ix + 1, iy + 1);
            const int v = ((v00 * (S - fx) + v10 * fx) * (S - fy) + (v01 * (S - fx) + v11 * fx) * fy) >> (2 * sh);
            acc += amps[o] * v;
        }
        int img = acc / 187;
        img = 128 + (((img - 128) * 3) >> 1);
#pragma unroll
        for (int k = 0; k < 12; k++) {
            const long long dx = x - occ.v[k][0], dy = y - occ.v[k][1], rad = occ.v[k][2];
            const bool in = occ.v[k][4] == 0 ? (dx * dx + dy * dy < rad * rad) : ((dx < 0 ? -dx : dx) < rad && (dy < 0 ? -dy : dy) < rad / 2 + 1);
            if (in) img += occ.v[k][3];
        }
        const u32 hn = h32(((u32)x * 0x27d4eb2du) ^ h32((u32)y ^ (seed * 977u + 12345u)));
        const int nz = (int)((hn & 3) + ((hn >> 4) & 3) + ((hn >> 8) & 3) + ((hn >> 12) & 3)) - 6;
        out[p] = (uint8_t)min(max(img + nz, 0), 255);
    }
}

} /* namespace */
