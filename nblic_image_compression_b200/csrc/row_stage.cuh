/*
 * row_stage.cuh -- the causal-neighbourhood rows of a coder stream staged in shared memory by cp.async
 * (BASELINE.json north_star way 3; R: NBLIC.c:287-304, QNBLIC.c:48-79 are the loads this replaces).
 *
 * The per-pixel front ends need pixels j-2 .. j+2 of the two rows above (and, where the pixels are known in advance,
 * j-2 .. j of the current row).  Read per lane from global memory that is 12 byte loads with 11 border predicates.
 * Here a TILE of 64 pixels (+ 2 on either side) of every row travels from HBM / L2 to shared memory as 16-byte
 * cp.async (LDGSTS) requests -- whole aligned 16-byte chunks of the source, whatever the row's alignment -- one tile
 * ahead of its use, and the reference's border fallbacks are MATERIALISED as halo cells once per row end:
 *
 *      row i-1:  r1[-1] = r1[-2] = r1[0]        r1[w] = r1[w+1] = r1[w-1]         (c, q at the left; d, t at the right)
 *      row i-2:  r2[-1] = r2[-2] = r2[0]        r2[w] = r2[w+1] = r2[w-1]         (h, s; g, r)
 *      row i  :  r0[-1] = r0[-2] = r1[0]                                          (a = b and e = a at j = 0)
 *
 * so an interior pixel and a border pixel run the same predicate-free loads.  The one fallback a halo cannot express
 * is e at j = 1 (e = a, while r0[-1] holds b for j = 0): the caller patches that lane.  Rows 0 and 1 (where the
 * fallbacks depend on the column) keep the positional sampler.
 *
 * A chunk is requested only when it holds at least one byte of the image, so nothing outside the raster's own
 * 16-byte granules is touched.
 */
#pragma once
#include "codec_core.cuh"

namespace nblic {

constexpr int kStageTile = 64;   /* pixels per staged tile (default) */
constexpr int kStageLine = 96;   /* bytes per staged row: six 16-byte chunks cover 2 + 64 + 2 pixels at any source alignment */
__host__ __device__ constexpr int stage_line_bytes(int tile) { return tile + 32; } /* 2 + tile + 2 pixels + up to 15 bytes of misalignment, in whole chunks */

NB_DEV void cp_async16(void *smem_dst, const void *gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
NB_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

/* Request pixels [j_lo - 2, j_lo + TILE + 2) of the image row starting at `row` into the line `dst` (TILE + 32 bytes of
 * 16-byte aligned shared memory); lanes sl = 0 .. nl-1 of the caller share the chunks.  [lo, hi) = the raster's bytes.
 * Returns the offset of pixel j_lo inside the line. */
template <int TILE>
NB_DEV int stage_issue(uint8_t *dst, const uint8_t *row, int j_lo, const uint8_t *lo, const uint8_t *hi, int sl, int nl) {
    const unsigned long long p = (unsigned long long)(row + j_lo) - 2ull, a = p & ~15ull;
    for (int c = sl; c < stage_line_bytes(TILE) / 16; c += nl) {
        const unsigned long long src = a + 16ull * (unsigned)c;
        if (src + 16ull > (unsigned long long)lo && src < (unsigned long long)hi) cp_async16(dst + 16 * c, reinterpret_cast<const void *>(src));
    }
    return (int)(p - a) + 2;
}

/* Halo cells of a staged row of the two rows above: px -> pixel j_lo of the line.  One lane calls it after the tile landed. */
template <int TILE>
NB_DEV void stage_halo(uint8_t *px, int j_lo, int w) {
    if (j_lo == 0) px[-1] = px[-2] = px[0];
    const int e = w - j_lo; /* first column past the row end, relative to the tile */
    if (e <= TILE + 1) px[e] = px[e + 1] = px[e - 1];
}

/*
 * Double-buffered tiles of NROWS rows (3: rows i, i-1, i-2 -- the pixels are all known, lossless encoders;
 * 2: rows i-1, i-2 -- decoders and reconstruction-feedback encoders, whose current row is being produced).
 * Shared memory: 2 * NROWS * (TILE + 32) bytes at `buf`.  All lanes of the group (sl = 0 .. nl-1; the whole warp or the
 * lanes of one sub-warp stream) call every method together; the waits are warp-wide.
 */
template <int NROWS, int TILE = kStageTile> struct RowStage {
    static constexpr int kLine = stage_line_bytes(TILE);
    static constexpr int kBytes = 2 * NROWS * kLine;
    uint8_t *buf;
    const uint8_t *lo, *hi; /* raster extent */
    int w, cur, pend_i, pend_j;
    int off[NROWS], pend_off[NROWS];

    NB_DEV void start(uint8_t *smem, const uint8_t *img, int h, int width) {
        buf = smem; lo = img; hi = img + (size_t)h * width; w = width; cur = 0; pend_i = pend_j = -1;
#pragma unroll
        for (int r = 0; r < NROWS; r++) off[r] = pend_off[r] = 0;
    }
    NB_DEV uint8_t *line(int which, int r) const { return buf + (which * NROWS + r) * kLine; }
    /* pointer to pixel (row r of the window, column j) of the current tile whose first column is j_lo; r = 0 is the
     * newest row (NROWS == 3: row i, then i-1, i-2;  NROWS == 2: row i-1, then i-2) */
    NB_DEV const uint8_t *at(int r, int j_rel) const { return line(cur, r) + off[r] + j_rel; }

    NB_DEV void issue(int i, int j_lo, int sl, int nl) {
        const int newest = NROWS == 3 ? i : i - 1;
#pragma unroll
        for (int r = 0; r < NROWS; r++) pend_off[r] = stage_issue<TILE>(line(cur ^ 1, r), lo + (size_t)(newest - r) * w, j_lo, lo, hi, sl, nl);
        pend_i = i; pend_j = j_lo;
    }
    /* Make tile (i, j_lo) current (it is requested now unless it was prefetched), fix its halos, then (if next_i >= 0)
     * prefetch tile (next_i, next_j).  i >= 2.  `leader`: the one lane of the group that writes the halo cells. */
    NB_DEV void advance(int i, int j_lo, int next_i, int next_j, int sl, int nl, bool leader) {
        if (pend_i != i || pend_j != j_lo) issue(i, j_lo, sl, nl);
        cp_async_wait_all();
        __syncwarp();
        cur ^= 1;
#pragma unroll
        for (int r = 0; r < NROWS; r++) off[r] = pend_off[r];
        if (leader) {
            if (NROWS == 3) {
                stage_halo<TILE>(line(cur, 1) + off[1], j_lo, w);
                stage_halo<TILE>(line(cur, 2) + off[2], j_lo, w);
                if (j_lo == 0) { uint8_t *p0 = line(cur, 0) + off[0]; p0[-1] = p0[-2] = line(cur, 1)[off[1]]; }
            } else {
                stage_halo<TILE>(line(cur, 0) + off[0], j_lo, w);
                stage_halo<TILE>(line(cur, 1) + off[1], j_lo, w);
            }
        }
        __syncwarp();
        pend_i = pend_j = -1;
        if (next_i >= 0) issue(next_i, next_j, sl, nl);
    }
};

/* the window of pixel j (tile-relative jr) from staged rows i, i-1, i-2; the caller patches e at j == 1 */
template <int TILE>
NB_DEV void sample_staged3(const RowStage<3, TILE> &st, int jr, Nb &n, int &x) {
    const uint8_t *p0 = st.at(0, jr), *p1 = st.at(1, jr), *p2 = st.at(2, jr);
    x = p0[0]; n.a = p0[-1]; n.e = p0[-2];
    n.b = p1[0]; n.c = p1[-1]; n.q = p1[-2]; n.d = p1[1]; n.t = p1[2];
    n.f = p2[0]; n.h = p2[-1]; n.s = p2[-2]; n.g = p2[1]; n.r = p2[2];
}

} /* namespace nblic */
