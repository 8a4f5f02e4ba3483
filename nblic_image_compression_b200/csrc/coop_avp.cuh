/*
 * coop_avp.cuh -- warp-cooperative AVP: the int64 recursive weighted-least-squares predictor of
 * NBLIC effort 2 (n = 6 neighbours, m = 43 accumulators) and effort 3 (n = 10, m = 111).
 * R: NBLIC.c:112-283 (AVPsolveAxb, AVPgetVecN, AVPprecalcuate, AVPpredict, AVPupdate).
 *
 * Layout.  Accumulator k of the m-vector  [s | b(n) | A(n x n)]  lives in lane k % 32, slot k / 32:
 *   E     registers (row-local, zeroed per row)
 *   B, F  global memory, one m-vector per image column, read and written with coalesced 8-byte accesses
 * The two ridge-regularised systems of a pixel are solved side by side: half-warp 0 owns the system
 * of ridge r1, half-warp 1 that of r2.  The augmented matrices sit in shared memory (AvpSmem::ds);
 * every elimination step spreads its (n-1-k)(n-k) quotients over the 16 lanes of the half-warp.
 * All quotients of a step share the pivot as divisor, so one double-precision reciprocal per step
 * feeds an exact 64-bit division (estimate, multiply back, correct by one) -- div_rcp below.
 *
 * Exactness: products wrap modulo 2^64 (wmul), quotients truncate toward zero, pivoting picks the
 * first row of maximal magnitude, a zero pivot abandons the system (the caller falls back to the
 * gradient predictor), exactly as the reference.
 */
#pragma once
#include "codec_core.cuh"

namespace nblic {

template <int N> struct AvpGeom {
    static constexpr int M = 1 + N + N * N;
    static constexpr int NS = (M + 31) / 32; /* accumulator slots per lane */
};

struct __align__(16) AvpSmem {
    i64 ds[2][112]; /* the two systems' [s | b | A] */
    int vec[16];    /* neighbour values minus 128, reference order a b c d e f t h q g (R: NBLIC.c:164-183) */
};

/* ---- exact signed 64-bit division by a divisor whose reciprocal is known ------------------------ */
struct Rcp64 {
    double r;   /* 1 / |d| */
    u64 ud;     /* |d| */
    bool neg;   /* d < 0 */
};
NB_DEV Rcp64 make_rcp(i64 d) {
    Rcp64 rc;
    rc.neg = d < 0;
    rc.ud = rc.neg ? (u64)0 - (u64)d : (u64)d;
    rc.r = 1.0 / __ull2double_rn(rc.ud);
    return rc;
}
/* trunc(n / d), n and d any int64 with d != 0 (two's-complement wrap like x86 idiv, minus the trap).
 * Fast path: the fp64 estimate has relative error < 2^-51, so for quotients below 2^50 its truncation is
 * within one of the answer.  Larger quotients (small pivots, seen on near-lossless reconstructions) take a
 * second estimate on the remainder, which brings the error back under one before the same correction.
 * Anything else (|d| >= 2^50) uses the plain 64-bit divide. */
NB_DEV i64 div_rcp(i64 n, const Rcp64 &rc) {
    const bool nneg = n < 0;
    const u64 un = nneg ? (u64)0 - (u64)n : (u64)n;
    const double qf = __ull2double_rn(un) * rc.r;
    u64 q;
    if (rc.ud < (1ull << 49)) {
        q = __double2ull_rz(qf);
        i64 r = (i64)(un - q * rc.ud); /* |r| <= (2^-51 * un / ud + 1) * ud < 2^63: no wrap */
        if (qf >= 1125899906842624.0 /* 2^50 */) { /* refine: |r| / ud <= 2^13 + 1, its estimate is within one */
            const bool rneg = r < 0;
            const u64 ur = rneg ? (u64)0 - (u64)r : (u64)r;
            const u64 q2 = __double2ull_rz(__ull2double_rn(ur) * rc.r);
            q = rneg ? q - q2 : q + q2;
            r = (i64)(un - q * rc.ud); /* now in (-2 ud, 2 ud) */
            if (r < 0) { q--; r += (i64)rc.ud; }
            if (r >= (i64)rc.ud) { q++; r -= (i64)rc.ud; }
        }
        if (r < 0) { q--; r += (i64)rc.ud; }
        if (r >= (i64)rc.ud) { q++; }
    } else {
        q = un / rc.ud;
    }
    return (nneg != rc.neg) ? (i64)((u64)0 - q) : (i64)q;
}

NB_DEV i64 shfl64(i64 v, int src) {
    const int lo = __shfl_sync(0xffffffffu, (int)(u32)(u64)v, src);
    const int hi = __shfl_sync(0xffffffffu, (int)(u32)((u64)v >> 32), src);
    return (i64)(((u64)(u32)hi << 32) | (u64)(u32)lo);
}
NB_DEV i64 shfl64_xor(i64 v, int mask) {
    const int lo = __shfl_xor_sync(0xffffffffu, (int)(u32)(u64)v, mask);
    const int hi = __shfl_xor_sync(0xffffffffu, (int)(u32)((u64)v >> 32), mask);
    return (i64)(((u64)(u32)hi << 32) | (u64)(u32)lo);
}

/* ---- row start: E = 0 and F[j] = decay(F[j+1]) + B[j], right to left (R: NBLIC.c:186-204,817-820) ---- */
template <int N>
NB_DEV void avp_row_start(i64 (&E)[AvpGeom<N>::NS], const i64 *Brow, i64 *Frow, int w, int lane) {
    constexpr int M = AvpGeom<N>::M, NS = AvpGeom<N>::NS;
    i64 F[NS];
#pragma unroll
    for (int s = 0; s < NS; s++) { E[s] = 0; F[s] = 0; }
    for (int j = w - 1; j >= 0; j--) {
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int k = lane + 32 * s;
            if (k < M) {
                const i64 prev = j == w - 1 ? 0 : (k == 0 ? avp_decay(F[s], 0) : avp_decay(F[s], 1));
                F[s] = prev + Brow[(size_t)j * M + k];
                Frow[(size_t)j * M + k] = F[s];
            }
        }
    }
}

/*
 * Both ridge solves of one pixel.  On return every lane holds ok1/ok2 and p1/p2 (fixed point, N_FRAC
 * fractional bits); ef0 = E[0] + F[0].  `vec` must already be in sm.vec.   R: NBLIC.c:112-161,210-239
 */
template <int N>
NB_DEV void avp_predict_pair(AvpSmem &sm, const i64 (&E)[AvpGeom<N>::NS], const i64 *Fj, i64 ridge1, i64 ridge2, int lane, int &ok1, int &ok2,
                             i64 &p1, i64 &p2, i64 &ef0) {
    constexpr int M = AvpGeom<N>::M, NS = AvpGeom<N>::NS;
    /* dataset = E + F, plus the ridge on b and on diag(A), written once per system */
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int k = lane + 32 * s;
        if (k < M) {
            const i64 v = E[s] + Fj[k];
            if (k == 0) ef0 = v;
            i64 v1 = v, v2 = v;
            if (k >= 1 && k <= N) { v1 += wshl(ridge1, N_FRAC - 2); v2 += wshl(ridge2, N_FRAC - 2); }
            else if (k > N && (k - 1 - N) % (N + 1) == 0) { v1 += wmul(ridge1, N); v2 += wmul(ridge2, N); }
            sm.ds[0][k] = v1; sm.ds[1][k] = v2;
        }
    }
    ef0 = shfl64(ef0, 0);
    __syncwarp();

    const int sys = lane >> 4, hl = lane & 15;
    i64 *b = sm.ds[sys] + 1, *A = sm.ds[sys] + 1 + N;
    bool alive = true; /* uniform inside a half-warp */

    /* For n = 10 the k loops stay rolled: unrolled, that solver alone is ~90 KB of SASS and every pixel streams
     * it through the 32 KB instruction cache (ncu: no_instruction was the top stall of the effort-3 kernel;
     * rolled +4 %).  The n = 6 solver is small enough to profit from unrolling (+5 %). */
    constexpr int kUnroll = N > 6 ? 1 : N;
#pragma unroll kUnroll
    for (int k = 0; k + 1 < N; k++) { /* forward elimination with partial pivoting */
        /* pivot = first row of maximal magnitude in column k (the order the reference's scan produces) */
        i64 best;
        int piv;
        if constexpr (N > 6) { /* lane hl looks at row k + hl, a 16-lane butterfly keeps (larger magnitude, then smaller row) */
            best = k + hl < N ? labs64(A[(k + hl) * N + k]) : (i64)0x8000000000000000ull;
            piv = k + hl < N ? k + hl : 127;
#pragma unroll
            for (int m = 8; m >= 1; m >>= 1) {
                const i64 ov = shfl64_xor(best, m);
                const int oi = __shfl_xor_sync(0xffffffffu, piv, m);
                if (ov > best || (ov == best && oi < piv)) { best = ov; piv = oi; }
            }
        } else { /* six rows or fewer: every lane scans them (cheaper than four shuffle rounds) */
            piv = k;
            best = labs64(A[k * N + k]);
            for (int r = k + 1; r < N; r++) {
                const i64 m = labs64(A[r * N + k]);
                if (m > best) { best = m; piv = r; }
            }
        }
        __syncwarp();
        if (alive && piv != k) { /* swap rows k and piv: columns k..N-1 and b */
            if (hl < N - k) { const i64 t0 = A[k * N + k + hl], t1 = A[piv * N + k + hl]; A[k * N + k + hl] = t1; A[piv * N + k + hl] = t0; }
            else if (hl == N - k) { const i64 t0 = b[k], t1 = b[piv]; b[k] = t1; b[piv] = t0; }
        }
        __syncwarp();
        const i64 d = A[k * N + k];
        if (d == 0) alive = false;
        if (alive) {
            const Rcp64 rc = make_rcp(d);
            const int W = N - k, cnt = (N - 1 - k) * W; /* per row: columns k+1..N-1, then b */
            constexpr u32 kWMagic[11] = {0, 65537, 32769, 21846, 16385, 13108, 10923, 9363, 8193, 7282, 6554}; /* 65536 / W + 1 */
            const u32 w_magic = kWMagic[W];             /* e / W == (e * w_magic) >> 16 for e < 128 */
            for (int e = hl; e < cnt; e += 16) {
                const int ro = (int)(((u32)e * w_magic) >> 16), cc = e - ro * W, r = k + 1 + ro;
                const i64 f = A[r * N + k];
                if (f != 0) { /* one call site for both kinds of entry: the two would otherwise run one after the other */
                    const bool in_a = cc < W - 1;
                    i64 *own = in_a ? &A[r * N + k + 1 + cc] : &b[r];
                    const i64 above = in_a ? A[k * N + k + 1 + cc] : b[k];
                    *own -= div_rcp(wmul(above, f), rc);
                }
            }
        }
        __syncwarp();
    }
#pragma unroll kUnroll
    for (int k = N - 1; k > 0; k--) { /* back substitution on b */
        const i64 d = A[k * N + k];
        if (d == 0) alive = false;
        if (alive && hl < k) {
            const i64 f = A[hl * N + k];
            if (f != 0) { const Rcp64 rc = make_rcp(d); b[hl] -= div_rcp(wmul(b[k], f), rc); }
        }
        __syncwarp();
    }
    /* px = 128.0 + sum_k round(4 b_k v_k / A_kk)   (R: NBLIC.c:227-237; A_00 is checked here, as there) */
    i64 term = 0;
    if (A[0] == 0) alive = false;
    if (alive && hl < N) {
        const i64 d = A[hl * N + hl];
        term = div_rcp(wshl(wmul(b[hl], (i64)sm.vec[hl]), 2) + (d >> 1), make_rcp(d));
    }
#pragma unroll
    for (int m = 8; m >= 1; m >>= 1) term += shfl64_xor(term, m);
    const i64 px = clampl(((i64)128 << N_FRAC) + term, 0, (i64)255 << N_FRAC);
    p1 = shfl64(px, 0); p2 = shfl64(px, 16);
    ok1 = __shfl_sync(0xffffffffu, (int)alive, 0); ok2 = __shfl_sync(0xffffffffu, (int)alive, 16);
    __syncwarp();
}

/* Accumulator update after the pixel value x is known.  R: NBLIC.c:242-283 */
template <int N>
NB_DEV void avp_learn_coop(AvpSmem &sm, i64 (&E)[AvpGeom<N>::NS], i64 *Bj, int x, i64 s_now, i64 s_sum, int lane) {
    constexpr int M = AvpGeom<N>::M, NS = AvpGeom<N>::NS;
    const i64 xc = x - 128;
    s_sum = clampl(s_sum + (1 << N_FRAC), 1 << N_FRAC, 16 << N_FRAC);
    const i64 half = s_sum >> 1;
    const Rcp64 rc = make_rcp(s_sum);
#pragma unroll
    for (int s = 0; s < NS; s++) {
        const int k = lane + 32 * s;
        if (k < M) {
            i64 t = s_now;
            if (k > 0) { /* b: x' * v_r << 28; A: v_r * v_c << 18 -- one shared division */
                const int idx = k - 1 - N; /* >= 0 for the A block */
                const i64 left = k <= N ? xc : (i64)sm.vec[idx / N];
                const i64 right = (i64)sm.vec[k <= N ? k - 1 : idx % N];
                t = div_rcp(wshl(wmul(left, right), k <= N ? 28 : 18) + half, rc);
            }
            const i64 nb = (k == 0 ? avp_decay(Bj[k], 0) : avp_decay(Bj[k], 1)) + t;
            Bj[k] = nb;
            E[s] = (k == 0 ? avp_decay(E[s], 0) : avp_decay(E[s], 1)) + nb;
        }
    }
}

} /* namespace nblic */
