/*
 * stream_kernels.cu -- host side of libnblic_b200.so: contexts, scratch management, launch planning
 * (which kernel family, how many resident streams, where the rank tables live) and the batch C ABI of
 * include/nblic_b200.h.  The kernels themselves are in kernels.cuh and the headers it includes.
 *
 * One launch per kernel family and batch: images are grouped (QNBLIC / lossless effort-1 encode /
 * per-pixel front end by effort / sequential kernels), sorted largest first and pulled from an atomic
 * queue by a persistent grid sized to the residency of the kernel (cudaOccupancyMaxActiveBlocksPerMultiprocessor).
 * Encode results are compacted by scan_lengths_kernel + gather_streams_kernel.
 *
 * No CPU fallback anywhere: without a usable CUDA device every entry point fails.
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nblic_b200.h"
#include "kernels.cuh"

namespace {

/* ---- host-side helpers ----------------------------------------------------------------------- */

/* Device scratch that only grows.  Stream-ordered allocation (cudaMallocAsync / cudaFreeAsync on the owning context's
 * stream): cudaFree would wait for EVERY stream of the device, i.e. for the kernels of other contexts -- a context that
 * grows a buffer, or is destroyed, next to another context's minutes-long single-image launch must not stall on it. */
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaStream_t st = nullptr; /* set by nblic_b200_create */
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaMallocAsync(&p, want, st);
        if (e != cudaSuccess) { (void)cudaGetLastError(); want = bytes; e = cudaMallocAsync(&p, want, st); }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFreeAsync(p, st); p = nullptr; cap = 0; }
};

/* Page-locked host staging for the small per-call tables (task records, offsets, flags).  Copies between pageable host
 * memory and the device make the driver wait for the stream INSIDE the copy call, which serialises the host threads of
 * concurrent contexts (the lanes of split_over_lanes): measured, four lanes ran one after the other.  With pinned
 * staging every copy is a plain asynchronous DMA and only cudaStreamSynchronize waits. */
struct HostBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        release();
        const size_t want = bytes + bytes / 4 + 4096;
        if (pool_take(want, &p, &cap)) return cudaSuccess;
        const cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocPortable);
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    void release() { if (p) pool_give(p, cap); p = nullptr; cap = 0; }

    /* Released staging goes to a process-wide free list instead of cudaFreeHost (which, like cudaFree, waits for the
     * whole device); the list is bounded by the peak staging ever in use -- a few hundred bytes per image. */
    struct Spare { void *p; size_t cap; };
    static std::mutex &pool_lock() { static std::mutex m; return m; }
    static std::vector<Spare> &pool() { static std::vector<Spare> v; return v; }
    static bool pool_take(size_t want, void **out, size_t *cap_out) {
        std::lock_guard<std::mutex> g(pool_lock());
        std::vector<Spare> &v = pool();
        size_t best = v.size();
        for (size_t k = 0; k < v.size(); k++) if (v[k].cap >= want && (best == v.size() || v[k].cap < v[best].cap)) best = k;
        if (best == v.size()) return false;
        *out = v[best].p; *cap_out = v[best].cap;
        v.erase(v.begin() + (long)best);
        return true;
    }
    static void pool_give(void *q, size_t c) { std::lock_guard<std::mutex> g(pool_lock()); pool().push_back({q, c}); }
};

thread_local std::string g_create_error; /* message of a failed nblic_b200_create on this thread */

} /* namespace */

struct nblic_b200_ctx {
    int device = 0;
    int sm_count = 0;
    int mapping = NBLIC_B200_MAP_AUTO;
#ifdef NBLIC_B200_SEQUENTIAL
    bool serial_only = getenv("NBLIC_B200_SERIAL") != nullptr; /* debugging aid: force the sequential kernels */
#else
    bool serial_only = false;
#endif
    cudaStream_t stream = nullptr, copy = nullptr; /* compute / host<->device copies of the host-buffer calls */
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_copy[2] = {nullptr, nullptr};
    std::string error;
    uint64_t launches = 0;
    float coder_ms = 0.f, coder_ms_total = 0.f; /* most recent batch call / accumulated by split_over_lanes */
    const char *last_map = "none";
    int last_slots = 0; /* resident streams the most recent cooperative launch could hold */
    int co_streams = 0; /* lanes: streams of the whole batch, which share the GPU with this lane's piece (0 = this call is alone) */
    int sub_lps = getenv("NBLIC_B200_LPS") ? atoi(getenv("NBLIC_B200_LPS")) : 0; /* experiments: lanes per stream of the effort-1 decoder (32 = one stream per warp) */
    bool sub_forest_smem = getenv("NBLIC_B200_FOREST_SMEM") != nullptr;            /* experiments: counter forest in shared memory instead of L2 */
    int qpipe = getenv("NBLIC_B200_QPIPE") ? atoi(getenv("NBLIC_B200_QPIPE")) : -1; /* experiments: force (1) / forbid (0) the whole-GPU QNBLIC encode */                   /* experiments: whole bias table in shared memory */
    int e1pipe = getenv("NBLIC_B200_E1PIPE") ? atoi(getenv("NBLIC_B200_E1PIPE")) : -1; /* experiments: force (1) / forbid (0) the whole-GPU effort-1 encode */
    DevBuf tasks, order, queue, slots, sym, cold, coop_counts, sub_scratch, pipe_meta, pipe_sorted, pipe_counts, avp, offsets, flags, pixels, streams, recon, peeks;
    HostBuf h_tasks, h_order, h_small; /* pinned staging: task records both ways, launch order, offsets / flags / header peeks */
    /* whole-GPU effort-1 encode (pipe_nblic.cuh): images in flight side by side, each on its own stream and scratch set */
    struct E1Set {
        DevBuf rec, key, perm, sorted, counts, yz, doff, sums, visrec, decrec, p1, coded, totals;
        cudaStream_t st = nullptr;
        cudaEvent_t done = nullptr;
        DevBuf *all[13] = {&rec, &key, &perm, &sorted, &counts, &yz, &doff, &sums, &visrec, &decrec, &p1, &coded, &totals};
    };
    enum { kE1Sets = 32 };
    std::vector<E1Set> e1p;
    cudaEvent_t ev_e1 = nullptr; /* marks the task-table upload for the sets' streams */
    int occ_warp[2] = {0, 0};
    nblic_b200_ctx *lane[4] = {}; /* host-buffer calls: sub-contexts coding the pieces of a batch side by side (see split_over_lanes) */
};

namespace {

std::vector<DevBuf *> all_buffers(nblic_b200_ctx *c) {
    return {&c->tasks, &c->order, &c->queue, &c->slots, &c->sym, &c->cold, &c->coop_counts, &c->sub_scratch, &c->pipe_meta, &c->pipe_sorted, &c->pipe_counts,
            &c->avp, &c->offsets, &c->flags, &c->pixels, &c->streams, &c->recon, &c->peeks};
}

bool fail(nblic_b200_ctx *c, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->error = buf; else g_create_error = buf;
    return false;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { fail(c, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); return -1; } } while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct LaunchPlan { int map; int grid; size_t cold_stride; size_t smem; };

/* Persistent grid for n queued streams when `slots` can be resident.  Measured (round 2): spreading 10 000 streams evenly
 * over 4 waves of 2 500 instead of filling 3 108 slots made both effort-1 kernels 6 % SLOWER -- throughput grows with
 * the resident warps (issue slots are only 62-66 % used), and the dynamic queue already packs the tail.  So: fill. */
int balanced_grid(int n, int slots) { return std::max(1, std::min(n, slots)); }

#ifdef NBLIC_B200_SEQUENTIAL
template <int KIND, bool DEC, int MAP>
int launch_coder_t(nblic_b200_ctx *c, int n_order, const int *d_order, int *d_queue, const LaunchPlan &plan, size_t avp_stride) {
    auto kern = coder_kernel<KIND, DEC, MAP>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
    kern<<<plan.grid, 32, plan.smem, c->stream>>>((Task *)c->tasks.p, d_order, n_order, d_queue, (uint8_t *)c->cold.p, plan.cold_stride,
                                                   avp_stride ? (i64 *)c->avp.p : nullptr, avp_stride);
    c->launches++;
    CK(cudaGetLastError());
    return 0;
}

template <int KIND, bool DEC>
int launch_coder(nblic_b200_ctx *c, int n_order, const int *d_order, int *d_queue, int max_w, int max_effort) {
    int map = c->mapping;
    if (map == NBLIC_B200_MAP_AUTO || map == NBLIC_B200_MAP_WARP4) map = MAP_WARP;
    LaunchPlan plan;
    plan.map = map;
    if (map == MAP_WARP) {
        plan.smem = KIND == KIND_N ? N_SMEM_BYTES : Q_STATE_BYTES;
        plan.cold_stride = KIND == KIND_N ? N_COUNT_BYTES : 0;
        int per_sm = 0;
        auto kern = coder_kernel<KIND, DEC, MAP_WARP>;
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)plan.smem));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, plan.smem));
        plan.grid = std::min(n_order, c->sm_count * std::max(per_sm, 1));
    } else {
        plan.smem = 0;
        plan.cold_stride = KIND == KIND_N ? N_STATE_BYTES : Q_STATE_BYTES;
        const int lanes = std::min(n_order, c->sm_count * 32 * 8);
        plan.grid = (lanes + 31) / 32;
    }
    size_t avp_stride = 0;
    if (KIND == KIND_N && max_effort >= 2) { /* AVP column accumulators: bound the slot count by a 24 GB budget */
        const int n = max_effort == 2 ? 6 : 10;
        avp_stride = 2 * (size_t)max_w * (1 + n + n * n);
        const size_t budget_slots = std::max<size_t>(((size_t)24 << 30) / (avp_stride * sizeof(i64)), 1);
        if (map == MAP_WARP) plan.grid = (int)std::min<size_t>((size_t)plan.grid, budget_slots);
        else plan.grid = (int)std::max<size_t>(std::min<size_t>((size_t)plan.grid, budget_slots / 32), 1);
    }
    const size_t slots = map == MAP_WARP ? (size_t)plan.grid : (size_t)plan.grid * 32;
    CK(c->cold.reserve(std::max<size_t>(slots * plan.cold_stride, 16)));
    if (avp_stride) CK(c->avp.reserve(slots * avp_stride * sizeof(i64)));
    c->last_map = map == MAP_WARP ? "warp" : "lane";
    if (map == MAP_WARP) return launch_coder_t<KIND, DEC, MAP_WARP>(c, n_order, d_order, d_queue, plan, avp_stride);
    return launch_coder_t<KIND, DEC, MAP_LANE>(c, n_order, d_order, d_queue, plan, avp_stride);
}

#else
/* product build: no sequential kernels (they ship in the test build libnblic_b200_seq.so only) */
template <int KIND, bool DEC>
int launch_coder(nblic_b200_ctx *c, int, const int *, int *, int, int) { fail(c, "this build carries no sequential kernels (NBLIC_B200_SEQUENTIAL)"); return -1; }
#endif

template <bool DEC>
int launch_coop_q(nblic_b200_ctx *c, int n_order, const int *d_order, int *d_queue) {
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, coop_q_kernel<DEC>, 32, 0));
    const int grid = balanced_grid(n_order, c->sm_count * std::max(per_sm, 1));
    if (!DEC) CK(c->coop_counts.reserve((size_t)grid * Q_TAB_STRIDE * sizeof(u32))); /* encoder: per-CTA symbol statistics in L2 */
    coop_q_kernel<DEC><<<grid, 32, 0, c->stream>>>((Task *)c->tasks.p, d_order, n_order, d_queue, (u32 *)c->coop_counts.p);
    c->launches++;
    c->last_slots = c->sm_count * std::max(per_sm, 1);
    CK(cudaGetLastError());
    return 0;
}

template <int NAVP, int MODE, bool RG>
int launch_coop_t(nblic_b200_ctx *c, int n_order, const int *d_order, int *d_queue, int max_w, int max_nodes, bool probe_only, int *slots_out) {
    /* per stream: fixed tables + the compacted counter forest; kCoopWarps streams per CTA */
    const size_t per_warp = align_up(CoopLayout<NAVP, MODE, RG>::kForestOff + sizeof(u32) * (size_t)max_nodes, 16), smem = per_warp * kCoopWarps;
    auto kern = coop_nblic_kernel<NAVP, MODE, RG>;
    int per_sm = 0;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * kCoopWarps, smem));
    const int slots = c->sm_count * std::max(per_sm, 1) * kCoopWarps;
    if (slots_out) *slots_out = slots;
    if (probe_only) return 0;
    int streams = balanced_grid(n_order, slots);
    size_t avp_half = 0;
    if (NAVP > 0) { /* B and F: one m-vector per image column each; bound the resident streams by a 32 GB budget */
        avp_half = (size_t)max_w * (1 + NAVP + NAVP * NAVP);
        streams = (int)std::min<size_t>((size_t)streams, std::max<size_t>(((size_t)32 << 30) / (2 * avp_half * sizeof(i64)), 1));
    }
    const int grid = (streams + kCoopWarps - 1) / kCoopWarps;
    if (NAVP > 0) CK(c->avp.reserve((size_t)grid * kCoopWarps * 2 * avp_half * sizeof(i64)));
    CK(c->coop_counts.reserve((size_t)grid * kCoopWarps * N_RANK_ENTRIES * 5)); /* frequencies (int) + rank tables (bytes) per stream */
    kern<<<grid, 32 * kCoopWarps, smem, c->stream>>>((Task *)c->tasks.p, d_order, n_order, d_queue, (int *)c->coop_counts.p, (i64 *)c->avp.p, avp_half, (u32)per_warp);
    c->launches++;
    c->last_slots = slots;
    CK(cudaGetLastError());
    return 0;
}

/* QNBLIC encode of few (large) images: every image's front end, bias chains and symbol counts use the whole GPU, one
 * launch sequence per image on the context's stream; a single launch then finishes all streams, one warp per image
 * (pipe_qnblic.cuh).  `idx` are task indices in launch order, `d_order` the same list on the device. */
int launch_qpipe_encode(nblic_b200_ctx *c, const std::vector<Task> &tasks, const std::vector<int> &idx, const int *d_order) {
    const int n = (int)idx.size();
    size_t max_px = 1, max_counts = 1;
    auto chunking = [](size_t px, int &chunk_px, int &n_chunks) {
        size_t cp = std::max<size_t>(4096, (px + 4095) / 4096);
        cp = (cp + 31) / 32 * 32;
        chunk_px = (int)cp;
        n_chunks = (int)((px + cp - 1) / cp);
    };
    for (int i : idx) {
        const size_t px = (size_t)tasks[(size_t)i].h * tasks[(size_t)i].w;
        int chunk_px, n_chunks;
        chunking(px, chunk_px, n_chunks);
        max_px = std::max(max_px, px);
        max_counts = std::max(max_counts, (size_t)kPipeKeys * n_chunks);
    }
    CK(c->pipe_meta.reserve(max_px * sizeof(u32)));
    CK(c->pipe_sorted.reserve(max_px * sizeof(uint2)));
    CK(c->pipe_counts.reserve((max_counts + kPipeKeys + 1) * sizeof(u32)));
    CK(c->coop_counts.reserve((size_t)n * Q_TAB_STRIDE * sizeof(u32)));
    CK(cudaMemsetAsync(c->coop_counts.p, 0, (size_t)n * Q_TAB_STRIDE * sizeof(u32), c->stream));
    u32 *meta = (u32 *)c->pipe_meta.p, *counts = (u32 *)c->pipe_counts.p, *key_start = counts + max_counts;
    uint2 *sorted = (uint2 *)c->pipe_sorted.p;
    const bool timing = getenv("NBLIC_B200_E1PIPE_TIMING") != nullptr; /* development aid: stage times of the first image on stderr */
    std::vector<std::pair<const char *, cudaEvent_t>> marks;
    auto mark = [&](int k, const char *name) {
        if (!timing || k != 0) return;
        cudaEvent_t ev;
        if (cudaEventCreate(&ev) == cudaSuccess) { cudaEventRecord(ev, c->stream); marks.push_back({name, ev}); }
    };
    for (int k = 0; k < n; k++) {
        const Task &t = tasks[(size_t)idx[(size_t)k]];
        const long long px = (long long)t.h * t.w;
        int chunk_px, n_chunks;
        chunking((size_t)px, chunk_px, n_chunks);
        const int front_blocks = 1 + (int)std::max<long long>(1, std::min<long long>((px + 255) / 256, (long long)c->sm_count * 8));
        mark(k, "start");
        qpipe_front_kernel<<<front_blocks, 256, 0, c->stream>>>(t.src, t.h, t.w, meta);
        mark(k, "front");
        qpipe_count_kernel<<<n_chunks, 256, 0, c->stream>>>(meta, px, chunk_px, n_chunks, counts);
        qpipe_scan_kernel<<<1, 1024, 0, c->stream>>>(counts, n_chunks, key_start);
        qpipe_scatter_kernel<<<n_chunks, 32, 0, c->stream>>>(meta, px, chunk_px, n_chunks, counts, sorted);
        mark(k, "sort by bias address");
        qpipe_chain_kernel<<<(kPipeKeys + 127) / 128, 128, 0, c->stream>>>(sorted, key_start, reinterpret_cast<uint16_t *>(t.sym),
                                                                         (u32 *)c->coop_counts.p + (size_t)k * Q_TAB_STRIDE);
        mark(k, "bias chains");
        c->launches += 5;
    }
    mark(0, "other images");
    qpipe_finish_kernel<<<n, 32, 0, c->stream>>>((Task *)c->tasks.p, d_order, n, (u32 *)c->coop_counts.p);
    mark(0, "histograms + rANS sweep");
    c->launches++;
    CK(cudaGetLastError());
    if (timing && !marks.empty()) {
        cudaDeviceSynchronize();
        for (size_t m = 1; m < marks.size(); m++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks[m - 1].second, marks[m].second);
            fprintf(stderr, "qpipe %-24s %8.3f ms\n", marks[m].first, ms);
        }
        for (auto &m : marks) cudaEventDestroy(m.second);
    }
    return 0;
}

/* The launches of one image's pipeline, in stream order.  Several images are in flight on their own streams; their
 * launch lists are issued BREADTH FIRST (step s of every image, then step s + 1): the streams share the device's few
 * hardware work queues, and depth-first issue would park image k + 8's short stage kernels in the queue behind image k's
 * chain of dependent launches (measured: three images' stages ran one after the other). */
typedef std::vector<std::function<void()>> StepList;

void issue_breadth_first(std::vector<StepList> &lists) {
    size_t longest = 0;
    for (const StepList &l : lists) longest = std::max(longest, l.size());
    for (size_t s = 0; s < longest; s++)
        for (StepList &l : lists) if (s < l.size()) l[s]();
}

/* Stable partition of n items by keys[] on scratch set e (pipe_nblic.cuh): e.perm = sorted position -> item, e.sorted =
 * the items' payload in sorted order; returns where the first position of every key (NKEYS + 1 entries) will be. */
template <int NKEYS, class T>
u32 *e1p_sort_steps(nblic_b200_ctx *c, nblic_b200_ctx::E1Set &e, const u32 *keys, const T *payload, long long n, StepList &steps) {
    const int chunk_items = (int)std::max<long long>(4096, ((n + 255) / 256 + 31) / 32 * 32); /* at most 256 chunks */
    const int n_chunks = (int)std::max<long long>(1, (n + chunk_items - 1) / chunk_items);
    u32 *counts = (u32 *)e.counts.p, *key_total = counts + (size_t)kE1NodeKeys * 256, *key_start = key_total + kE1NodeKeys;
    u32 *perm = (u32 *)e.perm.p;
    T *sorted = (T *)e.sorted.p;
    cudaStream_t st = e.st;
    steps.push_back([=] { psort_count_kernel<NKEYS><<<n_chunks, 256, 0, st>>>(keys, n, chunk_items, n_chunks, counts); });
    steps.push_back([=] { psort_scan_keys_kernel<<<(NKEYS + 7) / 8, 256, 0, st>>>(counts, NKEYS, n_chunks, key_total); });
    steps.push_back([=] { psort_scan_totals_kernel<<<1, 1024, 0, st>>>(key_total, NKEYS, key_start); });
    steps.push_back([=] { psort_scatter_kernel<NKEYS, T><<<n_chunks, 32, 0, st>>>(keys, payload, n, chunk_items, n_chunks, counts, key_start, perm, sorted); });
    c->launches += 4;
    return key_start;
}

/* Lossless effort-1 encode of few (large) images: every stage but the range coder is spread over the whole GPU
 * (pipe_nblic.cuh).  Up to kE1Sets images are in flight side by side, each on its own stream and scratch set: the stage
 * kernels of one Kodak-size image occupy a fraction of the machine for a few hundred microseconds each, and the range
 * coder of an image is a single warp.  Two passes over the group because the number of binary decisions of an image,
 * which sizes the second half's buffers, comes back from the device in between.  `idx` = task indices. */
int launch_e1pipe_encode(nblic_b200_ctx *c, const std::vector<Task> &tasks, const std::vector<int> &idx) {
    using Set = nblic_b200_ctx::E1Set;
    const int n = (int)idx.size();
    size_t max_px = 1;
    for (int i : idx) max_px = std::max(max_px, (size_t)tasks[(size_t)i].h * tasks[(size_t)i].w);
    /* scratch per image in flight: ~70 bytes per pixel; keep the sets inside a 24 GB budget */
    const int sets = (int)std::max<size_t>(1, std::min<size_t>({(size_t)n, (size_t)nblic_b200_ctx::kE1Sets, ((size_t)24 << 30) / (max_px * 70)}));
    if ((int)c->e1p.size() < sets) c->e1p.resize((size_t)sets);
    CK(c->h_small.reserve(sizeof(unsigned long long) * (size_t)sets));
    unsigned long long *h_total = (unsigned long long *)c->h_small.p;
    if (!c->ev_e1) CK(cudaEventCreateWithFlags(&c->ev_e1, cudaEventDisableTiming));
    CK(cudaEventRecord(c->ev_e1, c->stream));
    for (int k = 0; k < sets; k++) {
        Set &e = c->e1p[(size_t)k];
        if (!e.st) CK(cudaStreamCreateWithFlags(&e.st, cudaStreamNonBlocking));
        if (!e.done) CK(cudaEventCreateWithFlags(&e.done, cudaEventDisableTiming));
        for (DevBuf *b : e.all) b->st = e.st;
        CK(cudaStreamWaitEvent(e.st, c->ev_e1, 0)); /* the task table upload */
        CK(e.counts.reserve(((size_t)kE1NodeKeys * 256 + 2 * kE1NodeKeys + 1) * sizeof(u32)));
        CK(e.totals.reserve(16));
    }
    const int wide = c->sm_count * 8;
    /* NBLIC_B200_E1PIPE_TIMING: per-stage CUDA-event times of the first image on stderr (development aid) */
    const bool timing = getenv("NBLIC_B200_E1PIPE_TIMING") != nullptr;
    std::vector<std::pair<const char *, cudaEvent_t>> marks;
    auto mark = [&](int k, const char *name, cudaStream_t st, StepList &steps) {
        if (!timing || k != 0) return;
        steps.push_back([&marks, name, st] {
            cudaEvent_t ev;
            if (cudaEventCreate(&ev) == cudaSuccess) { cudaEventRecord(ev, st); marks.push_back({name, ev}); }
        });
    };
    Task *d_tasks = (Task *)c->tasks.p;
    for (int g0 = 0; g0 < n; g0 += sets) {
        const int cnt = std::min(sets, n - g0);
        std::vector<StepList> lists((size_t)cnt);
        for (int k = 0; k < cnt; k++) { /* first half: up to the decision count */
            Set &e = c->e1p[(size_t)k];
            StepList &steps = lists[(size_t)k];
            const Task &t = tasks[(size_t)idx[(size_t)(g0 + k)]];
            const long long px = (long long)t.h * t.w;
            const int blocks = (int)std::max<long long>(1, std::min<long long>((px + 255) / 256, wide));
            const int scan_blocks = (int)((px + kE1ScanBlock - 1) / kE1ScanBlock);
            CK(e.rec.reserve((size_t)px * 4)); CK(e.yz.reserve((size_t)px * 4)); CK(e.doff.reserve((size_t)px * 4));
            CK(e.key.reserve((size_t)px * 4)); CK(e.perm.reserve((size_t)px * 4)); CK(e.sorted.reserve((size_t)px * 4));
            CK(e.sums.reserve((size_t)scan_blocks * 4));
            u32 *rec = (u32 *)e.rec.p, *yz = (u32 *)e.yz.p, *key = (u32 *)e.key.p, *doff = (u32 *)e.doff.p, *sums = (u32 *)e.sums.p;
            const u32 *perm = (const u32 *)e.perm.p, *sorted = (const u32 *)e.sorted.p;
            unsigned long long *d_total = (unsigned long long *)e.totals.p, *h_dst = h_total + k;
            cudaStream_t st = e.st;
            const uint8_t *src = t.src;
            const int h = t.h, w = t.w, k_step = t.k_step;
            steps.push_back([=] { cudaMemsetAsync(d_total, 0, 16, st); });
            mark(g0 + k, "start", st, steps);
            steps.push_back([=] { e1p_front_kernel<<<blocks, 256, 0, st>>>(src, h, w, rec, key); });
            mark(g0 + k, "front", st, steps);
            const u32 *ks_a = e1p_sort_steps<kE1BiasKeys, u32>(c, e, key, rec, px, steps);
            mark(g0 + k, "sort by bias address", st, steps);
            steps.push_back([=] { e1p_bias_kernel<<<kE1BiasKeys / 128, 128, 0, st>>>(perm, sorted, ks_a, yz, key); });
            mark(g0 + k, "bias chains", st, steps);
            const u32 *ks_b = e1p_sort_steps<kE1RankKeys, u32>(c, e, key, yz, px, steps);
            mark(g0 + k, "sort by rank key", st, steps);
            steps.push_back([=] { e1p_rank_kernel<<<kE1RankKeys / 64, 64, 0, st>>>(perm, sorted, ks_b, yz); });
            mark(g0 + k, "rank chains", st, steps);
            steps.push_back([=] { e1p_count_kernel<<<scan_blocks, 256, 0, st>>>(yz, px, k_step, doff, sums, (int *)(d_total + 1)); });
            steps.push_back([=] { e1p_scan_kernel<<<1, 1024, 0, st>>>(sums, scan_blocks, d_total); });
            mark(g0 + k, "count + scan decisions", st, steps);
            steps.push_back([=] { cudaMemcpyAsync(h_dst, d_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st); });
            c->launches += 5;
        }
        issue_breadth_first(lists);
        CK(cudaGetLastError());
        for (StepList &l : lists) l.clear();
        for (int k = 0; k < cnt; k++) { /* second half: the decision count sizes the remaining buffers */
            Set &e = c->e1p[(size_t)k];
            StepList &steps = lists[(size_t)k];
            const int ti = idx[(size_t)(g0 + k)];
            const Task &t = tasks[(size_t)ti];
            const long long px = (long long)t.h * t.w;
            const int blocks = (int)std::max<long long>(1, std::min<long long>((px + 255) / 256, wide));
            CK(cudaStreamSynchronize(e.st));
            const unsigned long long n_dec = h_total[k];
            if (n_dec >= (1ull << 31)) { fail(c, "effort-1 pipeline: %llu decisions in one image", n_dec); return -1; }
            const size_t nv = std::max<size_t>(2 * (size_t)n_dec, 1);
            CK(e.key.reserve(nv * 4)); CK(e.perm.reserve(nv * 4)); CK(e.sorted.reserve(nv));
            CK(e.visrec.reserve(nv)); CK(e.decrec.reserve(nv / 2 + 1)); CK(e.p1.reserve(nv * 2)); CK(e.coded.reserve(nv + 2));
            const u32 *yz = (const u32 *)e.yz.p, *doff = (const u32 *)e.doff.p, *sums = (const u32 *)e.sums.p, *perm = (const u32 *)e.perm.p;
            u32 *key = (u32 *)e.key.p;
            uint8_t *visrec = (uint8_t *)e.visrec.p, *decrec = (uint8_t *)e.decrec.p;
            const uint8_t *sorted = (const uint8_t *)e.sorted.p;
            uint16_t *p1 = (uint16_t *)e.p1.p, *coded = (uint16_t *)e.coded.p;
            const unsigned long long *d_total = (const unsigned long long *)e.totals.p;
            cudaStream_t st = e.st;
            const int k_step = t.k_step;
            steps.push_back([=] { e1p_emit_kernel<<<blocks, 256, 0, st>>>(yz, px, k_step, doff, sums, key, visrec, decrec); });
            mark(g0 + k, "emit visits", st, steps);
            const u32 *ks = e1p_sort_steps<kE1NodeKeys, uint8_t>(c, e, key, visrec, 2 * (long long)n_dec, steps);
            mark(g0 + k, "sort visits by node", st, steps);
            steps.push_back([=] { e1p_node_kernel<<<kE1NodeKeys / 4, 128, 0, st>>>(perm, sorted, ks, p1); });
            mark(g0 + k, "node chains", st, steps);
            const int mix_blocks = (int)std::max<unsigned long long>(1, std::min<unsigned long long>((n_dec + 255) / 256, (unsigned long long)wide));
            steps.push_back([=] { e1p_mix_kernel<<<mix_blocks, 256, 0, st>>>(p1, decrec, n_dec, coded); });
            mark(g0 + k, "mix", st, steps);
            steps.push_back([=] { e1p_coder_kernel<<<1, 32, 0, st>>>(d_tasks, ti, coded, d_total, (const int *)(d_total + 1)); });
            mark(g0 + k, "range coder", st, steps);
            c->launches += 4;
        }
        issue_breadth_first(lists);
        CK(cudaGetLastError());
    }
    if (timing && !marks.empty()) {
        cudaDeviceSynchronize();
        for (size_t m = 1; m < marks.size(); m++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks[m - 1].second, marks[m].second);
            fprintf(stderr, "e1pipe %-24s %8.3f ms\n", marks[m].first, ms);
        }
        for (auto &m : marks) cudaEventDestroy(m.second);
    }
    for (int k = 0; k < sets; k++) { /* the context's stream continues when every image is coded */
        CK(cudaEventRecord(c->e1p[(size_t)k].done, c->e1p[(size_t)k].st));
        CK(cudaStreamWaitEvent(c->stream, c->e1p[(size_t)k].done, 0));
    }
    return 0;
}

/* Effort-1 decode with 32 / LPS streams per warp; FORESTG: counter forest in L2 instead of shared memory. */
template <int LPS, bool FORESTG>
int launch_subwarp_t(nblic_b200_ctx *c, int n_packs, const int *d_packs, int *d_queue, int max_nodes) {
    const size_t smem = SubLayout<LPS, FORESTG>::bytes(max_nodes);
    auto kern = subwarp_decode_kernel<LPS, FORESTG>;
    int per_sm = 0;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32, smem));
    const int warps = c->sm_count * std::max(per_sm, 1);
    const int grid = balanced_grid(n_packs, warps);
    CK(c->sub_scratch.reserve((size_t)grid * (32 / LPS) * kSubScratchBytes));
    kern<<<grid, 32, smem, c->stream>>>((Task *)c->tasks.p, d_packs, n_packs, d_queue, (uint8_t *)c->sub_scratch.p, max_nodes);
    c->launches++;
    c->last_slots = warps * (32 / LPS);
    CK(cudaGetLastError());
    return 0;
}

int launch_subwarp(nblic_b200_ctx *c, int lps, int n_packs, const int *d_packs, int *d_queue, int max_nodes) {
    (void)lps;
    if (c->sub_forest_smem) return launch_subwarp_t<8, false>(c, n_packs, d_packs, d_queue, max_nodes);
    return launch_subwarp_t<8, true>(c, n_packs, d_packs, d_queue, max_nodes);
}

/* Lanes per stream for an effort-1 decode of n streams.  Both kernels are latency-bound, so their rate grows with the
 * streams they hold: four streams per warp run at ~0.27 MPix/s per stream however many there are (10 000 fit), one stream
 * per warp at 0.53 (21 per SM, full) to 0.85 MPix/s (alone) but at most 21 x SMs at a time -- 1.65 GPix/s.  The packed
 * kernel wins from about 6 000 streams on (40 per SM). */
int choose_sub_lps(nblic_b200_ctx *c, int n_streams) {
    if (c->sub_lps == 8 || c->sub_lps == 32) return c->sub_lps;
    if (c->mapping == NBLIC_B200_MAP_WARP4) return 8;
    if (c->mapping == NBLIC_B200_MAP_WARP) return 32;
    return std::max(n_streams, c->co_streams) >= c->sm_count * 40 ? 8 : 32;
}

/* Efforts 2/3 always keep the rank tables in L2.  Effort 1 keeps them in shared memory (lowest latency per
 * pixel) unless the batch has more images than that layout can hold resident; then the L2 layout more than
 * doubles the resident streams. */
template <int NAVP, int MODE>
int launch_coop(nblic_b200_ctx *c, int n_order, const int *d_order, int *d_queue, int max_w, int max_nodes) {
    if (NAVP > 0) return launch_coop_t<NAVP, MODE, true>(c, n_order, d_order, d_queue, max_w, max_nodes, false, nullptr);
    int slots_smem = 0;
    if (launch_coop_t<NAVP, MODE, false>(c, n_order, d_order, d_queue, max_w, max_nodes, true, &slots_smem)) return -1;
    const char *force = getenv("NBLIC_B200_RANK");  /* "smem" / "l2": override for experiments */
    const bool use_l2 = force ? force[0] == 'l' : std::max(n_order, c->co_streams) > slots_smem;
    if (use_l2) return launch_coop_t<NAVP, MODE, true>(c, n_order, d_order, d_queue, max_w, max_nodes, false, nullptr);
    return launch_coop_t<NAVP, MODE, false>(c, n_order, d_order, d_queue, max_w, max_nodes, false, nullptr);
}

/* Upload tasks, run the coder kernels for the Q and N groups, download the task results. */
template <bool DEC>
int run_tasks(nblic_b200_ctx *c, std::vector<Task> &tasks) {
    const int n = (int)tasks.size();
    /* launch groups: 0 QNBLIC, 1 NBLIC sequential kernels, 2 lossless effort-1 encode, 3..5 per-pixel front end, effort 1..3 */
    enum { G_Q, G_SEQ, G_E1_LOSSLESS, G_FB1, G_FB2, G_FB3, N_GROUPS };
    std::vector<int> group[N_GROUPS];
    int max_w[N_GROUPS] = {1, 1, 1, 1, 1, 1}, max_nodes[N_GROUPS] = {0, 0, 0, 0, 0, 0}, seq_effort = 0;
    const bool coop_ok = c->mapping != NBLIC_B200_MAP_LANE && !c->serial_only;
    for (int i = 0; i < n; i++) {
        const Task &t = tasks[(size_t)i];
        if (t.status != NBLIC_B200_OK) continue;
        int g;
        if (t.effort == 0) g = G_Q;
        else if (!coop_ok) { g = G_SEQ; seq_effort = std::max(seq_effort, t.effort); }
        else if (!DEC && t.effort == 1 && t.near == 0) g = G_E1_LOSSLESS;
        else g = G_FB1 + t.effort - 1;
        group[g].push_back(i);
        max_w[g] = std::max(max_w[g], t.w);
        if (t.effort > 0) max_nodes[g] = std::max(max_nodes[g], forest_nodes(t.k_step));
    }
    auto by_size = [&](int a, int b) {
        const long long pa = (long long)tasks[(size_t)a].h * tasks[(size_t)a].w, pb = (long long)tasks[(size_t)b].h * tasks[(size_t)b].w;
        return pa != pb ? pa > pb : a < b;
    };
    std::vector<int> order;
    size_t start[N_GROUPS];
    for (int g = 0; g < N_GROUPS; g++) {
        std::sort(group[g].begin(), group[g].end(), by_size);
        start[g] = order.size();
        order.insert(order.end(), group[g].begin(), group[g].end());
    }
    /* effort-1 decode with several streams per warp: packs of equal-size streams (the list is sorted by size) */
    const int sub_lps = DEC && !group[G_FB1].empty() ? choose_sub_lps(c, (int)group[G_FB1].size()) : 32;
    const size_t packs_at = order.size();
    int n_packs = 0;
    if (sub_lps < 32) {
        const int spw = 32 / sub_lps;
        const std::vector<int> &g = group[G_FB1];
        for (size_t at = 0; at < g.size();) {
            const Task &first = tasks[(size_t)g[at]];
            int k = 0;
            for (; k < spw && at + k < g.size(); k++) {
                const Task &t = tasks[(size_t)g[at + k]];
                if (t.h != first.h || t.w != first.w) break;
                order.push_back(g[at + k]);
            }
            for (int pad = k; pad < spw; pad++) order.push_back(-1);
            at += (size_t)k;
            n_packs++;
        }
    }
    CK(c->tasks.reserve(sizeof(Task) * (size_t)std::max(n, 1)));
    CK(c->order.reserve(sizeof(int) * std::max<size_t>(order.size(), 1)));
    CK(c->queue.reserve(N_GROUPS * sizeof(int)));
    CK(c->h_tasks.reserve(sizeof(Task) * (size_t)std::max(n, 1)));
    CK(c->h_order.reserve(sizeof(int) * std::max<size_t>(order.size(), 1)));
    memcpy(c->h_tasks.p, tasks.data(), sizeof(Task) * (size_t)n);
    memcpy(c->h_order.p, order.data(), sizeof(int) * order.size());
    CK(cudaMemcpyAsync(c->tasks.p, c->h_tasks.p, sizeof(Task) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    if (!order.empty()) CK(cudaMemcpyAsync(c->order.p, c->h_order.p, sizeof(int) * order.size(), cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->queue.p, 0, N_GROUPS * sizeof(int), c->stream));

    CK(cudaEventRecord(c->ev0, c->stream));
    bool pipelined = false;
    for (int g = 0; g < N_GROUPS; g++) {
        if (group[g].empty()) continue;
        const int cnt = (int)group[g].size();
        const int *d_ord = (const int *)c->order.p + start[g];
        int *d_q = (int *)c->queue.p + g;
        int rc = 0;
        switch (g) {
            case G_Q:
                /* few images: one warp per image would leave the GPU empty and the front end on the critical path */
                if (!DEC && coop_ok && (c->qpipe == 1 || (c->qpipe != 0 && cnt <= 8 && c->co_streams == 0))) /* measured cross-over on Kodak-size images: ~10 */ { rc = launch_qpipe_encode(c, tasks, group[g], d_ord); pipelined = true; }
                else rc = coop_ok ? launch_coop_q<DEC>(c, cnt, d_ord, d_q) : launch_coder<KIND_Q, DEC>(c, cnt, d_ord, d_q, 1, 0);
                break;
            case G_SEQ: rc = launch_coder<KIND_N, DEC>(c, cnt, d_ord, d_q, max_w[g], seq_effort); break;
            case G_E1_LOSSLESS:
                /* few images: the stages of every image spread over the whole GPU, only the range coders are one warp each */
                if (c->e1pipe == 1 || (c->e1pipe != 0 && cnt <= 64 && c->co_streams == 0)) /* measured cross-over: ~90 images (24 Kodak-size images: 84 ms against 213 ms) */ { rc = launch_e1pipe_encode(c, tasks, group[g]); pipelined = true; }
                else rc = launch_coop<0, 0>(c, cnt, d_ord, d_q, max_w[g], max_nodes[g]);
                break;
            case G_FB1:
                if (sub_lps < 32) rc = launch_subwarp(c, sub_lps, n_packs, (const int *)c->order.p + packs_at, d_q, max_nodes[g]);
                else rc = launch_coop<0, DEC ? 2 : 1>(c, cnt, d_ord, d_q, max_w[g], max_nodes[g]);
                break;
            case G_FB2: rc = launch_coop<6, DEC ? 2 : 1>(c, cnt, d_ord, d_q, max_w[g], max_nodes[g]); break;
            case G_FB3: rc = launch_coop<10, DEC ? 2 : 1>(c, cnt, d_ord, d_q, max_w[g], max_nodes[g]); break;
        }
        if (rc) return -1;
        if (g >= G_E1_LOSSLESS || (g == G_Q && coop_ok)) c->last_map = "warp-coop";
        if ((g == G_Q || g == G_E1_LOSSLESS) && pipelined) c->last_map = "gpu-pipeline";
        if (g == G_FB1 && sub_lps < 32) c->last_map = "4-streams-per-warp";
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    return 0;
}

int finish_tasks(nblic_b200_ctx *c, std::vector<Task> &tasks) {
    CK(cudaMemcpyAsync(c->h_tasks.p, c->tasks.p, sizeof(Task) * tasks.size(), cudaMemcpyDeviceToHost, c->stream)); /* run_tasks sized h_tasks */
    CK(cudaStreamSynchronize(c->stream));
    memcpy(tasks.data(), c->h_tasks.p, sizeof(Task) * tasks.size());
    CK(cudaEventElapsedTime(&c->coder_ms, c->ev0, c->ev1));
    return 0;
}

void resolve_mode(int near, int effort, int &near_out, int &effort_out) { /* R: NBLIC_main.c:182-189, NBLIC.c:768-770 */
    if (near == 0 && effort == 0) { near_out = 0; effort_out = 0; return; }
    near_out = std::min(std::max(near, 0), 9);
    effort_out = std::min(std::max(effort, 1), 3);
}

} /* namespace */

/* ============================================================================================ */
/* C ABI                                                                                        */
/* ============================================================================================ */

extern "C" {

#ifdef NBLIC_B200_SEQUENTIAL
const char *nblic_b200_version(void) { return "nblic_b200 0.2 (sm_100a, test build with the sequential kernels)"; }
#else
const char *nblic_b200_version(void) { return "nblic_b200 0.2 (sm_100a)"; }
#endif

size_t nblic_b200_stream_bound(int height, int width) {
    if (height <= 0 || width <= 0) return 8192;
    return 2 * (size_t)height * (size_t)width + 8192;
}

nblic_b200_ctx *nblic_b200_create(int device) {
    /* Concurrent streams (host lanes, the single-image pipelines) map onto the device's hardware work queues: 8 by
     * default, 32 at most.  Only effective when this is the process's first CUDA call; never overrides the caller's choice. */
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) { fail(nullptr, "no CUDA device: %s", cudaGetErrorString(e)); return nullptr; }
    if (device < 0 || device >= count) { fail(nullptr, "device %d out of range (have %d)", device, count); return nullptr; }
    nblic_b200_ctx *c = new (std::nothrow) nblic_b200_ctx;
    if (!c) return nullptr;
    c->device = device;
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking)) != cudaSuccess || (e = cudaEventCreate(&c->ev0)) != cudaSuccess ||
        (e = cudaEventCreate(&c->ev1)) != cudaSuccess || (e = cudaEventCreateWithFlags(&c->ev_copy[0], cudaEventDisableTiming)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&c->ev_copy[1], cudaEventDisableTiming)) != cudaSuccess) {
        fail(nullptr, "device %d setup failed: %s", device, cudaGetErrorString(e));
        delete c;
        return nullptr;
    }
    if (prop.major < 10) { fail(nullptr, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor); delete c; return nullptr; }
    c->sm_count = prop.multiProcessorCount;
    for (DevBuf *b : all_buffers(c)) b->st = c->stream;
    { /* keep released scratch in the device's allocation pool instead of handing it back at every synchronisation */
        cudaMemPool_t pool;
        uint64_t keep = ~0ull;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    return c;
}

void nblic_b200_destroy(nblic_b200_ctx *c) {
    if (!c) return;
    for (nblic_b200_ctx *&l : c->lane) { nblic_b200_destroy(l); l = nullptr; }
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->copy) cudaStreamSynchronize(c->copy);
    for (DevBuf *b : all_buffers(c)) b->release();
    for (nblic_b200_ctx::E1Set &e : c->e1p) {
        for (DevBuf *b : e.all) b->release();
        if (e.st) { cudaStreamSynchronize(e.st); cudaStreamDestroy(e.st); }
        if (e.done) cudaEventDestroy(e.done);
    }
    if (c->ev_e1) cudaEventDestroy(c->ev_e1);
    if (c->stream) cudaStreamSynchronize(c->stream); /* the stream-ordered frees */
    c->h_tasks.release(); c->h_order.release(); c->h_small.release();
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    for (cudaEvent_t e : c->ev_copy) if (e) cudaEventDestroy(e);
    if (c->copy) cudaStreamDestroy(c->copy);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char *nblic_b200_last_error(const nblic_b200_ctx *c) { return c ? c->error.c_str() : g_create_error.c_str(); }

int nblic_b200_set_mapping(nblic_b200_ctx *c, int mapping) {
    if (!c || mapping < NBLIC_B200_MAP_AUTO || mapping > NBLIC_B200_MAP_WARP4) return -1;
#ifndef NBLIC_B200_SEQUENTIAL
    if (mapping == NBLIC_B200_MAP_LANE) { fail(c, "MAP_LANE needs the sequential kernels, which only the test build carries"); return -1; }
#endif
    c->mapping = mapping;
    for (nblic_b200_ctx *l : c->lane) if (l) l->mapping = mapping;
    return 0;
}

int nblic_b200_set_pipeline(nblic_b200_ctx *c, int mode) {
    if (!c || mode < NBLIC_B200_PIPE_AUTO || mode > NBLIC_B200_PIPE_ALWAYS) return -1;
    c->qpipe = c->e1pipe = mode;
    return 0;
}

uint64_t nblic_b200_launch_count(const nblic_b200_ctx *c) {
    if (!c) return 0;
    uint64_t total = c->launches;
    for (const nblic_b200_ctx *l : c->lane) if (l) total += l->launches;
    return total;
}
float nblic_b200_last_coder_ms(const nblic_b200_ctx *c) { return c ? c->coder_ms : 0.f; }
const char *nblic_b200_last_mapping(const nblic_b200_ctx *c) { return c ? c->last_map : "none"; }
int nblic_b200_last_slots(const nblic_b200_ctx *c) { return c ? c->last_slots : 0; }
void *nblic_b200_stream_handle(const nblic_b200_ctx *c) { return c ? (void *)c->stream : nullptr; }

void *nblic_b200_host_alloc(size_t bytes) {
    void *p = nullptr;
    return cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) == cudaSuccess ? p : nullptr;
}
void nblic_b200_host_free(void *p) { if (p) cudaFreeHost(p); }

int nblic_b200_peek(const uint8_t *stream, size_t len, int *height, int *width, int *near, int *effort) {
    if (!stream) return -1;
    const Peek p = peek_bytes(stream, len);
    if (height) *height = p.h;
    if (width) *width = p.w;
    if (near) *near = p.near;
    if (effort) *effort = p.effort;
    return p.ok ? 0 : -1;
}

int nblic_b200_encode_batch_device(nblic_b200_ctx *c, int n, const uint8_t *d_pixels, const uint64_t *pix_off, const int *heights,
                                   const int *widths, int near, int effort, uint8_t *d_streams, uint64_t stream_cap,
                                   uint64_t *stream_off, uint8_t *d_recon, int *status) {
    if (!c) return -1;
    if (n < 0 || (n > 0 && (!d_pixels || !pix_off || !heights || !widths || !d_streams || !stream_off))) { fail(c, "bad arguments"); return -1; }
    CK(cudaSetDevice(c->device));
    int near_c, effort_c;
    resolve_mode(near, effort, near_c, effort_c);
    if (n == 0) { if (stream_off) stream_off[0] = 0; return 0; }

    std::vector<Task> tasks((size_t)n);
    size_t slot_total = 0, sym_total = 0, recon_total = 0;
    for (int i = 0; i < n; i++) {
        Task &t = tasks[(size_t)i];
        memset(&t, 0, sizeof t);
        t.h = heights[i]; t.w = widths[i]; t.near = near_c; t.effort = effort_c;
        t.k_step = std::min(std::max(3 + 2 * near_c, 3), 16);
        const bool ok = t.h > 0 && t.w > 0 && t.h <= 65535 && t.w <= 65535 && (long long)t.h * t.w <= 100000000LL;
        t.status = ok ? NBLIC_B200_OK : NBLIC_B200_BAD_DIMS;
        if (!ok) continue;
        const size_t px = (size_t)t.h * t.w;
        t.slot_cap = (u32)std::min<size_t>(nblic_b200_stream_bound(t.h, t.w), 0xfffffff0u);
        slot_total += align_up(t.slot_cap, 256);
        if (effort_c == 0) sym_total += align_up(2 * px, 256);
        if (near_c > 0 && !d_recon) recon_total += align_up(px, 256);
    }
    CK(c->slots.reserve(std::max<size_t>(slot_total, 16)));
    CK(c->sym.reserve(std::max<size_t>(sym_total, 16)));
    CK(c->recon.reserve(std::max<size_t>(recon_total, 16)));
    size_t slot_at = 0, sym_at = 0, recon_at = 0;
    for (int i = 0; i < n; i++) {
        Task &t = tasks[(size_t)i];
        if (t.status != NBLIC_B200_OK) continue;
        const size_t px = (size_t)t.h * t.w;
        t.src = d_pixels + pix_off[i];
        t.slot = (uint8_t *)c->slots.p + slot_at; slot_at += align_up(t.slot_cap, 256);
        if (effort_c == 0) { t.sym = (uint8_t *)c->sym.p + sym_at; sym_at += align_up(2 * px, 256); }
        if (near_c > 0) {
            if (d_recon) t.rec = d_recon + pix_off[i];
            else { t.rec = (uint8_t *)c->recon.p + recon_at; recon_at += align_up(px, 256); }
        }
    }
    if (run_tasks<false>(c, tasks)) return -1;

    /* compaction: exclusive scan of the lengths, then a coalesced gather */
    CK(c->offsets.reserve(sizeof(unsigned long long) * ((size_t)n + 1)));
    CK(c->flags.reserve(sizeof(int)));
    CK(cudaMemsetAsync(c->flags.p, 0, sizeof(int), c->stream));
    scan_lengths_kernel<<<1, 1024, 0, c->stream>>>((const Task *)c->tasks.p, n, (unsigned long long *)c->offsets.p);
    c->launches++;
    u32 max_cap = 0;
    for (const Task &t : tasks) max_cap = std::max(max_cap, t.slot_cap);
    const u32 chunk = 16384;
    dim3 grid((unsigned)n, (max_cap + chunk - 1) / chunk); /* image index on x (2^31-1 blocks), piece index on y (<= 200 MB / 16 KB) */
    gather_streams_kernel<<<grid, 256, 0, c->stream>>>((const Task *)c->tasks.p, (const unsigned long long *)c->offsets.p, d_streams,
                                                        stream_cap, chunk, (int *)c->flags.p);
    c->launches++;
    CK(cudaGetLastError());
    CK(c->h_small.reserve(sizeof(uint64_t) * ((size_t)n + 2)));
    uint64_t *h_off = (uint64_t *)c->h_small.p;
    int *h_flag = (int *)(h_off + n + 1);
    CK(cudaMemcpyAsync(h_off, c->offsets.p, sizeof(uint64_t) * ((size_t)n + 1), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(h_flag, c->flags.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (finish_tasks(c, tasks)) return -1;
    memcpy(stream_off, h_off, sizeof(uint64_t) * ((size_t)n + 1));
    const int overflow = *h_flag;
    int failed = 0;
    for (int i = 0; i < n; i++) {
        int st = tasks[(size_t)i].status;
        if (st == NBLIC_B200_OK && overflow && stream_off[i + 1] > stream_cap) st = NBLIC_B200_OVERFLOW;
        if (status) status[i] = st;
        failed += st != NBLIC_B200_OK;
    }
    return failed;
}

/* starts[i], lens[i]: extent of stream i inside d_streams */
int decode_device_impl(nblic_b200_ctx *c, int n, const uint8_t *d_streams, const uint64_t *starts, const uint64_t *lens,
                              uint8_t *d_pixels, const uint64_t *pix_off, const uint64_t *pix_cap, int *status) {
    /* header fields of every stream, gathered on the device */
    CK(c->offsets.reserve(sizeof(unsigned long long) * 2 * (size_t)n));
    CK(c->peeks.reserve(sizeof(Peek) * (size_t)n));
    unsigned long long *d_starts = (unsigned long long *)c->offsets.p, *d_lens = d_starts + n;
    CK(c->h_small.reserve((2 * sizeof(uint64_t) + sizeof(Peek)) * (size_t)n));
    uint64_t *h_starts = (uint64_t *)c->h_small.p;
    Peek *peeks = (Peek *)(h_starts + 2 * (size_t)n);
    memcpy(h_starts, starts, sizeof(uint64_t) * (size_t)n);
    memcpy(h_starts + n, lens, sizeof(uint64_t) * (size_t)n);
    CK(cudaMemcpyAsync(d_starts, h_starts, 2 * sizeof(uint64_t) * (size_t)n, cudaMemcpyHostToDevice, c->stream)); /* d_lens follows d_starts */
    peek_headers_kernel<<<(n + 127) / 128, 128, 0, c->stream>>>(d_streams, d_starts, d_lens, n, (Peek *)c->peeks.p);
    c->launches++;
    CK(cudaMemcpyAsync(peeks, c->peeks.p, sizeof(Peek) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));

    std::vector<Task> tasks((size_t)n);
    for (int i = 0; i < n; i++) {
        Task &t = tasks[(size_t)i];
        memset(&t, 0, sizeof t);
        const Peek &p = peeks[(size_t)i];
        t.h = p.h; t.w = p.w; t.near = p.near; t.k_step = p.k_step; t.effort = p.effort;
        t.status = p.ok ? NBLIC_B200_OK : NBLIC_B200_BAD_HEADER;
        t.slot = const_cast<uint8_t *>(d_streams) + starts[i];
        t.slot_cap = (u32)std::min<uint64_t>(lens[i], 0xfffffff0u);
        t.rec = d_pixels + pix_off[i];
        if (p.ok && p.effort == 0 && ((uintptr_t)t.slot & 1)) t.status = NBLIC_B200_BAD_HEADER; /* QNBLIC words must be 2-byte aligned */
        /* the raster size comes from the stream itself: never write more than the caller reserved */
        if (p.ok && pix_cap && (uint64_t)p.h * (uint64_t)p.w > pix_cap[i]) t.status = NBLIC_B200_OVERFLOW;
    }
    if (run_tasks<true>(c, tasks)) return -1;
    if (finish_tasks(c, tasks)) return -1;
    int failed = 0;
    for (int i = 0; i < n; i++) { if (status) status[i] = tasks[(size_t)i].status; failed += tasks[(size_t)i].status != NBLIC_B200_OK; }
    return failed;
}

int nblic_b200_decode_batch_device(nblic_b200_ctx *c, int n, const uint8_t *d_streams, const uint64_t *stream_off, uint8_t *d_pixels,
                                   const uint64_t *pix_off, const uint64_t *pix_cap, int *status) {
    if (!c) return -1;
    if (n < 0 || (n > 0 && (!d_streams || !stream_off || !d_pixels || !pix_off))) { fail(c, "bad arguments"); return -1; }
    CK(cudaSetDevice(c->device));
    if (n == 0) return 0;
    std::vector<uint64_t> lens((size_t)n);
    for (int i = 0; i < n; i++) lens[(size_t)i] = stream_off[i + 1] - stream_off[i];
    return decode_device_impl(c, n, d_streams, stream_off, lens.data(), d_pixels, pix_off, pix_cap, status);
}

static int encode_batch_one_lane(nblic_b200_ctx *c, int n, const uint8_t *const *images, const int *heights, const int *widths, int near,
                                 int effort, uint8_t *const *outs, const size_t *out_caps, size_t *out_lens, uint8_t *const *recon, int *status) {
    if (!c) return -1;
    if (n < 0 || (n > 0 && (!images || !heights || !widths || !outs || !out_caps || !out_lens))) { fail(c, "bad arguments"); return -1; }
    CK(cudaSetDevice(c->device));
    if (n == 0) return 0;
    int near_c, effort_c;
    resolve_mode(near, effort, near_c, effort_c);
    std::vector<uint64_t> pix_off((size_t)n), stream_off((size_t)n + 1);
    std::vector<int> st((size_t)n);
    size_t total = 0, stream_total = 256;
    bool want_recon = false;
    for (int i = 0; i < n; i++) {
        pix_off[(size_t)i] = total;
        const bool ok = heights[i] > 0 && widths[i] > 0 && heights[i] <= 65535 && widths[i] <= 65535 && (long long)heights[i] * widths[i] <= 100000000LL;
        if (ok) { total += align_up((size_t)heights[i] * widths[i], 256); stream_total += nblic_b200_stream_bound(heights[i], widths[i]); }
        if ((i + 1) % (c->sm_count * kChunkImagesPerSm) == 0) stream_total += 256; /* chunk bases are 256-byte aligned */
        if (recon && recon[i]) want_recon = true;
    }
    want_recon = want_recon && near_c > 0;
    CK(c->pixels.reserve(std::max<size_t>(total, 16) * (want_recon ? 2 : 1)));
    CK(c->streams.reserve(std::max<size_t>(stream_total, 16)));
    uint8_t *d_pix = (uint8_t *)c->pixels.p, *d_rec = want_recon ? d_pix + std::max<size_t>(total, 16) : nullptr;
    auto dims_fine = [&](int i) {
        return heights[i] > 0 && widths[i] > 0 && heights[i] <= 65535 && widths[i] <= 65535 && (long long)heights[i] * widths[i] <= 100000000LL;
    };
    /* Chunked pipeline: the upload of chunk k+1 and the download of chunk k-1 ride the copy stream while
     * chunk k is being coded.  A chunk is a whole number of waves for every kernel's residency (8/16/24 per SM)
     * and at least 8 waves long: a chunk boundary drains the persistent grid, which on a 10k-image batch cost
     * more (measured: -3 %) than hiding the 5 % of the step spent on PCIe gained. */
    const int chunk_n = c->sm_count * kChunkImagesPerSm;
    const int n_chunks = (n + chunk_n - 1) / chunk_n;
    auto upload = [&](int k) -> int {
        const int lo = k * chunk_n, hi = std::min(n, lo + chunk_n);
        for (int i = lo; i < hi;) { /* images that are adjacent on the host and on the device travel as one copy */
            if (!dims_fine(i)) { i++; continue; }
            size_t bytes = (size_t)heights[i] * widths[i];
            int j = i + 1;
            while (j < hi && dims_fine(j) && images[j] == images[i] + bytes && pix_off[(size_t)j] == pix_off[(size_t)i] + bytes) {
                bytes += (size_t)heights[j] * widths[j];
                j++;
            }
            CK(cudaMemcpyAsync(d_pix + pix_off[(size_t)i], images[i], bytes, cudaMemcpyHostToDevice, c->copy));
            i = j;
        }
        CK(cudaEventRecord(c->ev_copy[k & 1], c->copy));
        return 0;
    };
    if (upload(0)) { cudaStreamSynchronize(c->copy); return -1; }
    int failed = 0;
    size_t packed_base = 0; /* chunk k's packed streams start here inside c->streams */
    float coder_ms = 0.f;
    for (int k = 0; k < n_chunks; k++) {
        const int lo = k * chunk_n, hi = std::min(n, lo + chunk_n), cnt = hi - lo;
        if (k + 1 < n_chunks && upload(k + 1)) { cudaStreamSynchronize(c->copy); return -1; }
        CK(cudaStreamWaitEvent(c->stream, c->ev_copy[k & 1], 0));
        size_t bound_k = 0;
        for (int i = lo; i < hi; i++) if (dims_fine(i)) bound_k += nblic_b200_stream_bound(heights[i], widths[i]);
        uint8_t *d_packed = (uint8_t *)c->streams.p + packed_base;
        int rc = nblic_b200_encode_batch_device(c, cnt, d_pix, pix_off.data() + lo, heights + lo, widths + lo, near, effort, d_packed,
                                                std::max<size_t>(bound_k, 16), stream_off.data(), d_rec, st.data() + lo);
        if (rc < 0) { cudaStreamSynchronize(c->copy); return -1; }
        coder_ms += c->coder_ms;
        for (int i = lo; i < hi; i++) { /* the device call has completed: downloads go to the copy stream */
            out_lens[i] = 0;
            if (st[(size_t)i] == NBLIC_B200_OK) {
                const size_t at = (size_t)stream_off[(size_t)(i - lo)], len = (size_t)(stream_off[(size_t)(i - lo) + 1] - at);
                if (len > out_caps[i]) st[(size_t)i] = NBLIC_B200_OVERFLOW;
                else {
                    CK(cudaMemcpyAsync(outs[i], d_packed + at, len, cudaMemcpyDeviceToHost, c->copy));
                    out_lens[i] = len;
                    if (want_recon && recon[i])
                        CK(cudaMemcpyAsync(recon[i], d_rec + pix_off[(size_t)i], (size_t)heights[i] * widths[i], cudaMemcpyDeviceToHost, c->copy));
                }
            }
            if (status) status[i] = st[(size_t)i];
            failed += st[(size_t)i] != NBLIC_B200_OK;
        }
        packed_base += align_up(bound_k, 256);
    }
    c->coder_ms = coder_ms;
    CK(cudaStreamSynchronize(c->copy));
    return failed;
}

static int decode_batch_one_lane(nblic_b200_ctx *c, int n, const uint8_t *const *streams, const size_t *stream_lens, uint8_t *const *images,
                                 const size_t *img_caps, int *heights, int *widths, int *nears, int *efforts, int *status) {
    if (!c) return -1;
    if (n < 0 || (n > 0 && (!streams || !stream_lens || !images || !img_caps))) { fail(c, "bad arguments"); return -1; }
    CK(cudaSetDevice(c->device));
    if (n == 0) return 0;
    std::vector<uint64_t> pix_off((size_t)n), stream_off((size_t)n + 1);
    std::vector<int> st((size_t)n, NBLIC_B200_OK);
    std::vector<Peek> peeks((size_t)n);
    std::vector<uint64_t> lens((size_t)n), pix_cap((size_t)n);
    size_t pix_total = 0, stream_total = 0;
    for (int i = 0; i < n; i++) {
        Peek p = streams[i] ? peek_bytes(streams[i], stream_lens[i]) : Peek{0, 0, 0, 0, 0, 0};
        if (p.ok && (size_t)p.h * p.w > img_caps[i]) { p.ok = 0; st[(size_t)i] = NBLIC_B200_OVERFLOW; }
        else if (!p.ok) st[(size_t)i] = NBLIC_B200_BAD_HEADER;
        peeks[(size_t)i] = p;
        if (heights) heights[i] = p.h;
        if (widths) widths[i] = p.w;
        if (nears) nears[i] = p.near;
        if (efforts) efforts[i] = p.effort;
        pix_off[(size_t)i] = pix_total;
        stream_off[(size_t)i] = stream_total;
        /* a stream that fails the header check still occupies a 16-byte stub so the device sees the same verdict; one
         * whose raster does not fit img_caps[i] gets no bytes at all: its (valid) header must not reach the device, where
         * no room was reserved for its pixels */
        lens[(size_t)i] = p.ok ? (uint64_t)stream_lens[i]
                               : (st[(size_t)i] == NBLIC_B200_OVERFLOW ? 0 : (uint64_t)std::min<size_t>(streams[i] ? stream_lens[i] : 0, 16));
        pix_cap[(size_t)i] = p.ok ? (uint64_t)p.h * (uint64_t)p.w : 0;
        if (p.ok) pix_total += align_up((size_t)p.h * p.w, 256);
        stream_total += align_up(lens[(size_t)i], 16);
    }
    stream_off[(size_t)n] = stream_total;
    CK(c->pixels.reserve(std::max<size_t>(pix_total, 16)));
    CK(c->streams.reserve(std::max<size_t>(stream_total, 16)));
    const int chunk_n = c->sm_count * kChunkImagesPerSm; /* same pipeline as encode: upload k+1 / download k-1 behind the coding of chunk k */
    const int n_chunks = (n + chunk_n - 1) / chunk_n;
    auto upload = [&](int k) -> int {
        const int lo = k * chunk_n, hi = std::min(n, lo + chunk_n);
        for (int i = lo; i < hi; i++)
            if (lens[(size_t)i]) CK(cudaMemcpyAsync((uint8_t *)c->streams.p + stream_off[(size_t)i], streams[i], lens[(size_t)i], cudaMemcpyHostToDevice, c->copy));
        CK(cudaEventRecord(c->ev_copy[k & 1], c->copy));
        return 0;
    };
    if (upload(0)) { cudaStreamSynchronize(c->copy); return -1; }
    std::vector<int> dev_st((size_t)n, NBLIC_B200_OK);
    int failed = 0;
    float coder_ms = 0.f;
    for (int k = 0; k < n_chunks; k++) {
        const int lo = k * chunk_n, hi = std::min(n, lo + chunk_n);
        if (k + 1 < n_chunks && upload(k + 1)) { cudaStreamSynchronize(c->copy); return -1; }
        CK(cudaStreamWaitEvent(c->stream, c->ev_copy[k & 1], 0));
        int rc = decode_device_impl(c, hi - lo, (const uint8_t *)c->streams.p, stream_off.data() + lo, lens.data() + lo, (uint8_t *)c->pixels.p,
                                    pix_off.data() + lo, pix_cap.data() + lo, dev_st.data() + lo);
        if (rc < 0) { cudaStreamSynchronize(c->copy); return -1; }
        coder_ms += c->coder_ms;
        for (int i = lo; i < hi; i++) {
            if (st[(size_t)i] == NBLIC_B200_OK) st[(size_t)i] = dev_st[(size_t)i];
            if (status) status[i] = st[(size_t)i];
            failed += st[(size_t)i] != NBLIC_B200_OK;
        }
        for (int i = lo; i < hi;) { /* rasters that are adjacent on the device and on the host travel as one copy */
            if (st[(size_t)i] != NBLIC_B200_OK) { i++; continue; }
            size_t bytes = (size_t)peeks[(size_t)i].h * peeks[(size_t)i].w;
            int j = i + 1;
            while (j < hi && st[(size_t)j] == NBLIC_B200_OK && images[j] == images[i] + bytes && pix_off[(size_t)j] == pix_off[(size_t)i] + bytes) {
                bytes += (size_t)peeks[(size_t)j].h * peeks[(size_t)j].w;
                j++;
            }
            CK(cudaMemcpyAsync(images[i], (uint8_t *)c->pixels.p + pix_off[(size_t)i], bytes, cudaMemcpyDeviceToHost, c->copy));
            i = j;
        }
    }
    c->coder_ms = coder_ms;
    CK(cudaStreamSynchronize(c->copy));
    return failed;
}

/*
 * Host-buffer calls on batches of more than kLaneMinImagesPerSm images per SM.  One lane = upload a piece, code it,
 * download it, all on its own streams and scratch (a sub-context driven by its own host thread).  The batch is cut into
 * up to four pieces, one per lane, and the lanes run side by side: the pieces' kernels are co-resident, so the SMs see
 * the same number of coder streams as one launch over the whole batch would give them (the coder kernels are
 * latency-bound: their throughput is set by the resident streams), while the upload of piece k+1 rides under the
 * kernels of pieces <= k and the download of piece k under those of pieces > k.  Exposed PCIe time: the first
 * upload and the last download, i.e. the traffic of ONE piece (a quarter of the batch).
 * (Round 2 first tried two lanes taking alternate pieces of 24 images per SM: each launch then held a quarter of the
 * streams the 4-streams-per-warp decoder needs to fill the machine, and the end-to-end rate FELL to 0.53 of the
 * device-resident one.)
 */
constexpr int kLaneMinImagesPerSm = 8;
constexpr int kLanePieceMax = 16384; /* bounds the per-lane scratch; larger batches give every lane several pieces */

} /* extern "C" */

template <class Piece>
static int split_over_lanes(nblic_b200_ctx *c, int n, Piece piece) {
    /* Four lanes: eight (pieces of 1250 images on configs[4]) measured 0.58 of the device-resident rate end to end
     * against 0.95 with four, with 8 or 32 hardware work queues alike. */
    constexpr int kLanes = (int)(sizeof c->lane / sizeof c->lane[0]);
    const int piece_n = std::min(kLanePieceMax, std::max(c->sm_count * 2, (n + kLanes - 1) / kLanes));
    const int n_pieces = (n + piece_n - 1) / piece_n, lanes = std::min(kLanes, n_pieces);
    for (int who = 0; who < lanes; who++) {
        nblic_b200_ctx *&l = c->lane[who];
        if (!l) l = nblic_b200_create(c->device);
        if (!l) { fail(c, "lane context: %s", g_create_error.c_str()); return -1; }
        l->mapping = c->mapping;
        l->coder_ms_total = 0.f;
        l->co_streams = n;
    }
    int result[kLanes] = {};
    auto work = [&](int who) {
        for (int k = who; k < n_pieces; k += lanes) {
            const int lo = k * piece_n, cnt = std::min(n - lo, piece_n);
            const int rc = piece(c->lane[who], lo, cnt);
            if (rc < 0) { result[who] = -1; return; }
            result[who] += rc;
        }
    };
    std::vector<std::thread> others;
    for (int who = 1; who < lanes; who++) others.emplace_back(work, who);
    work(0);
    for (std::thread &t : others) t.join();
    c->coder_ms = 0.f;
    int total = 0;
    for (int who = 0; who < lanes; who++) {
        nblic_b200_ctx *l = c->lane[who];
        c->coder_ms = std::max(c->coder_ms, l->coder_ms_total); /* the lanes overlap: the longest one, not the sum */
        c->last_map = l->last_map; c->last_slots = l->last_slots;
        if (result[who] < 0) { fail(c, "%s", l->error.c_str()); total = -1; }
        else if (total >= 0) total += result[who];
    }
    CK(cudaSetDevice(c->device));
    return total;
}

extern "C" {

int nblic_b200_encode_batch(nblic_b200_ctx *c, int n, const uint8_t *const *images, const int *heights, const int *widths, int near,
                            int effort, uint8_t *const *outs, const size_t *out_caps, size_t *out_lens, uint8_t *const *recon, int *status) {
    if (!c) return -1;
    if (n <= c->sm_count * kLaneMinImagesPerSm) return encode_batch_one_lane(c, n, images, heights, widths, near, effort, outs, out_caps, out_lens, recon, status);
    if (!images || !heights || !widths || !outs || !out_caps || !out_lens) { fail(c, "bad arguments"); return -1; }
    return split_over_lanes(c, n, [&](nblic_b200_ctx *l, int lo, int cnt) {
        const int rc = encode_batch_one_lane(l, cnt, images + lo, heights + lo, widths + lo, near, effort, outs + lo, out_caps + lo, out_lens + lo,
                                             recon ? recon + lo : nullptr, status ? status + lo : nullptr);
        l->coder_ms_total += l->coder_ms;
        return rc;
    });
}

int nblic_b200_decode_batch(nblic_b200_ctx *c, int n, const uint8_t *const *streams, const size_t *stream_lens, uint8_t *const *images,
                            const size_t *img_caps, int *heights, int *widths, int *nears, int *efforts, int *status) {
    if (!c) return -1;
    if (n <= c->sm_count * kLaneMinImagesPerSm) return decode_batch_one_lane(c, n, streams, stream_lens, images, img_caps, heights, widths, nears, efforts, status);
    if (!streams || !stream_lens || !images || !img_caps) { fail(c, "bad arguments"); return -1; }
    return split_over_lanes(c, n, [&](nblic_b200_ctx *l, int lo, int cnt) {
        const int rc = decode_batch_one_lane(l, cnt, streams + lo, stream_lens + lo, images + lo, img_caps + lo, heights ? heights + lo : nullptr,
                                             widths ? widths + lo : nullptr, nears ? nears + lo : nullptr, efforts ? efforts + lo : nullptr,
                                             status ? status + lo : nullptr);
        l->coder_ms_total += l->coder_ms;
        return rc;
    });
}

int nblic_b200_debug_divcheck(nblic_b200_ctx *c, const int64_t *num, const int64_t *den, int n, int64_t *out) {
    if (!c || n < 0) return -1;
    if (n == 0) return 0;
    CK(cudaSetDevice(c->device));
    CK(c->sym.reserve(3 * sizeof(i64) * (size_t)n));
    i64 *d = (i64 *)c->sym.p;
    CK(cudaMemcpyAsync(d, num, sizeof(i64) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(d + n, den, sizeof(i64) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    divcheck_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(d, d + n, n, d + 2 * (size_t)n);
    c->launches++;
    CK(cudaMemcpyAsync(out, d + 2 * (size_t)n, sizeof(i64) * (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int nblic_b200_synth_gray(nblic_b200_ctx *c, uint8_t *d_out, int height, int width, uint32_t seed, const int32_t *occluders) {
    if (!c) return -1;
    if (!d_out || !occluders || height <= 0 || width <= 0) { fail(c, "bad arguments"); return -1; }
    CK(cudaSetDevice(c->device));
    Occluders occ;
    for (int k = 0; k < 12; k++) for (int f = 0; f < 5; f++) occ.v[k][f] = occluders[k * 5 + f];
    const long long n = (long long)height * width;
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)c->sm_count * 16);
    synth_gray_kernel<<<grid, 256, 0, c->stream>>>(d_out, height, width, seed, 0u, occ, nullptr);
    c->launches++;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int nblic_b200_synth_gray_batch(nblic_b200_ctx *c, uint8_t *d_out, int n, int height, int width, uint32_t seed0, uint32_t seed_stride,
                                const int32_t *occluders) {
    if (!c) return -1;
    if (n < 0 || (n > 0 && (!d_out || !occluders)) || height <= 0 || width <= 0) { fail(c, "bad arguments"); return -1; }
    if (n == 0) return 0;
    CK(cudaSetDevice(c->device));
    static_assert(sizeof(Occluders) == 60 * sizeof(int32_t), "occluder table layout");
    CK(c->sym.reserve(sizeof(Occluders) * (size_t)n));
    CK(cudaMemcpyAsync(c->sym.p, occluders, sizeof(Occluders) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    const long long px = (long long)height * width;
    const int gx = (int)std::min<long long>((px + 255) / 256, 64);
    Occluders none;
    memset(&none, 0, sizeof none);
    for (int at = 0; at < n; at += 32768) { /* gridDim.y <= 65535 */
        const int cnt = std::min(n - at, 32768);
        synth_gray_kernel<<<dim3((unsigned)gx, (unsigned)cnt), 256, 0, c->stream>>>(d_out + (size_t)at * (size_t)px, height, width,
                                                                                   seed0 + (uint32_t)at * seed_stride, seed_stride, none,
                                                                                   (const Occluders *)c->sym.p + at);
        c->launches++;
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

} /* extern "C" */
