/*
 * include/nblic_b200.h -- C ABI of libnblic_b200.so, the B200 (sm_100a) NBLIC / QNBLIC codec.
 *
 * Two layers, both plain C (pointers and sizes only, no C++/torch types):
 *
 *  1. The DROP-IN symbols.  Exactly the five functions the reference codec exports, with the
 *     reference's argument meaning, return values and side effects, so that the reference's own
 *     CLI (src/NBLIC_main.c:184-189,223-226) links against this library unchanged:
 *         NBLICcompress / NBLICdecompress                      replaces src/NBLIC.h:54,72  (src/NBLIC.c:915-926)
 *         QNBLICcompress / QNBLICcompressMultiThread /
 *         QNBLICdecompress                                     replaces src/QNBLIC.h:14-18 (src/QNBLIC.c:493-655,872-883)
 *     They are batch-of-one wrappers over layer 2 (csrc/nblic_dropin.c).
 *
 *  2. The BATCH ABI (nblic_b200_*): many images per call, one coder stream per warp (or per lane,
 *     see nblic_b200_set_mapping), host-buffer and device-resident variants.  This is what a
 *     throughput caller binds (cgo / JNI / ctypes stub in INTEGRATION.md).
 *
 * There is no CPU fallback: every entry point returns -1 when no CUDA device is usable.
 */
#ifndef NBLIC_B200_H
#define NBLIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- limits, as in src/NBLIC.h:29-31 and src/QNBLIC.h:9-11 ---------------------------------- */
#define NBLIC_MAX_HEIGHT     65535
#define NBLIC_MAX_WIDTH      65535
#define NBLIC_MAX_IMG_SIZE   100000000
#define QNBLIC_MAX_HEIGHT    65535
#define QNBLIC_MAX_WIDTH     65535
#define QNBLIC_MAX_IMG_SIZE  100000000

/* ---- layer 1: drop-in symbols ---------------------------------------------------------------- */

/* src/NBLIC.h:54.  Returns the stream length in bytes (>0) or -1.  *p_near is clipped to 0..9 and
 * *p_effort to 1..3 in place (src/NBLIC.c:768-770).  With *p_near > 0 the image buffer is overwritten
 * with the reconstruction, as the reference does (src/NBLIC.c:876,916). */
int NBLICcompress(int verbose, unsigned char *p_buf, unsigned char *p_img, int height, int width,
                  int *p_near, int *p_effort);

/* src/NBLIC.h:72.  Returns 0 / -1; writes height, width, near and effort parsed from the header.
 * The reference ABI carries no stream length: unless nblic_b200_hint_input_len() was called (see below) the
 * wrapper takes min(nblic_b200_stream_bound(h, w), readable extent of p_buf) bytes -- it probes, and never
 * faults on, the end of the caller's mapping, so an exact-size buffer is fine. */
int NBLICdecompress(int verbose, unsigned char *p_buf, unsigned char *p_img, int *p_height, int *p_width,
                    int *p_near, int *p_effort);

/* src/QNBLIC.h:14-18.  Compress returns the number of uint16 WORDS (the caller doubles it,
 * src/NBLIC_main.c:184-186) or -1; decompress returns 0 / -1 and rejects anything that does not
 * start with "Q0.2" before touching CUDA (the CLI uses it as a format sniff, src/NBLIC_main.c:223). */
int QNBLICdecompress(uint16_t *p_buf, unsigned char *p_img, int *p_height, int *p_width);
int QNBLICcompress(uint16_t *p_buf, unsigned char *p_img, int height, int width);
int QNBLICcompressMultiThread(uint16_t *p_buf, unsigned char *p_img, int height, int width);

/* Optional side channel for the two legacy decompress calls: number of valid bytes at p_buf for the
 * NEXT decompress call on this thread (reset after use).  Without it the wrappers copy
 * min(nblic_b200_stream_bound(h, w), 200 MB, readable extent) bytes through a fault-safe bounce buffer. */
void nblic_b200_hint_input_len(size_t n_bytes);

/* Worst-case stream size in bytes for an h x w image (any effort / near); what callers should
 * allocate per output stream. */
size_t nblic_b200_stream_bound(int height, int width);

/* ---- layer 2: batch ABI ---------------------------------------------------------------------- */

typedef struct nblic_b200_ctx nblic_b200_ctx;

/* One context per (host thread, GPU).  device = CUDA ordinal.  NULL when the device cannot be used. */
nblic_b200_ctx *nblic_b200_create(int device);
void nblic_b200_destroy(nblic_b200_ctx *ctx);
/* Message of the last failure on this context ("" if none).  ctx may be NULL (creation failures). */
const char *nblic_b200_last_error(const nblic_b200_ctx *ctx);

/* Stream-to-thread mapping of the coder kernels. */
enum {
    NBLIC_B200_MAP_AUTO = 0, /* pick by batch size                                               */
    NBLIC_B200_MAP_WARP = 1, /* one coder stream per warp, adaptive state in shared memory         */
    NBLIC_B200_MAP_LANE = 2, /* one coder stream per lane, the plain sequential formulation: only in the TEST build
                                (libnblic_b200_seq.so, the parity tests' second opinion); the product library returns -1 */
    NBLIC_B200_MAP_WARP4 = 3 /* as WARP, but the effort-1 decoder packs four streams of equal size into a warp
                                (csrc/subwarp_nblic.cuh); AUTO picks it from ~40 streams per SM on       */
};
int nblic_b200_set_mapping(nblic_b200_ctx *ctx, int mapping);

/* Single-image pipelines: a lossless effort-0 / effort-1 ENCODE of a small batch spreads every
 * stage but the entropy coder over the whole GPU (csrc/pipe_qnblic.cuh, csrc/pipe_nblic.cuh) instead of giving each
 * image one warp -- the bit-exact form of the reference's -t option (src/QNBLIC.c:660-866).  Same bytes either way. */
enum {
    NBLIC_B200_PIPE_AUTO = -1,  /* by batch size: small batches are pipelined (default) */
    NBLIC_B200_PIPE_NEVER = 0,
    NBLIC_B200_PIPE_ALWAYS = 1
};
int nblic_b200_set_pipeline(nblic_b200_ctx *ctx, int mode);

/* Per-image result codes written to `status` (may be NULL). */
enum {
    NBLIC_B200_OK = 0,
    NBLIC_B200_BAD_DIMS = 1,      /* src/NBLIC.c:717-729 / src/QNBLIC.c:33-45                       */
    NBLIC_B200_BAD_HEADER = 2,    /* src/NBLIC.c:698-712,733-745 / src/QNBLIC.c:475-486             */
    NBLIC_B200_OVERFLOW = 3,      /* output capacity too small                                      */
    NBLIC_B200_CORRUPT = 4        /* decode only: the stream drives the coder into a state no encoder output reaches
                                     (the reference indexes out of bounds / never returns there)    */
};

/*
 * Encode n images held in HOST memory.  effort 0 with near 0 selects QNBLIC ("Q0.2"), anything else
 * NBLIC ("NBLIC0.3") with near clipped to 0..9 and effort to 1..3 (dispatch of src/NBLIC_main.c:182-189).
 *   images[i]    height[i] x width[i] uint8 raster, top-down, no padding
 *   outs[i]      receives stream i; out_caps[i] bytes available; out_lens[i] = bytes written
 *   recon        NULL, or n pointers (entries may be NULL) receiving the reconstruction (near > 0)
 * Returns 0 when every image succeeded, the number of failed images (>0), or -1 on a CUDA error.
 */
int nblic_b200_encode_batch(nblic_b200_ctx *ctx, int n, const uint8_t *const *images, const int *heights,
                            const int *widths, int near, int effort, uint8_t *const *outs,
                            const size_t *out_caps, size_t *out_lens, uint8_t *const *recon, int *status);

/*
 * Decode n streams held in HOST memory (either container, sniffed per stream).
 *   images[i]    receives the raster; img_caps[i] bytes available
 *   heights / widths / nears / efforts   per-image outputs (effort 0 = QNBLIC); any may be NULL
 * Same return convention as encode.
 */
int nblic_b200_decode_batch(nblic_b200_ctx *ctx, int n, const uint8_t *const *streams, const size_t *stream_lens,
                            uint8_t *const *images, const size_t *img_caps, int *heights, int *widths,
                            int *nears, int *efforts, int *status);

/* Page-locked (pinned) host memory for the host-buffer calls above: copies from / to such buffers are asynchronous
 * DMA transfers that overlap the coder kernels (pageable buffers work too, through the driver's staging).  What the
 * reference's FileIO.c buffers (src/NBLIC_main.c:141: static 200 MB arrays) become in a pipelined caller
 * (csrc/nblic_batch_cli.c).  NULL on failure; a CUDA device must be usable. */
void *nblic_b200_host_alloc(size_t bytes);
void nblic_b200_host_free(void *p);

/* Parse a stream header on the host (no CUDA).  Returns 0 and fills the outputs, or -1.
 * effort 0 = "Q0.2".  Follows src/NBLIC.c:698-745 and src/QNBLIC.c:475-486. */
int nblic_b200_peek(const uint8_t *stream, size_t len, int *height, int *width, int *near, int *effort);

/*
 * Device-resident variants: pixel and stream buffers are CUDA device pointers on ctx's device; the
 * small per-image tables stay on the host.  Work is issued on ctx's own stream and is complete
 * when the call returns.
 *   d_pixels + pix_off[i]        image i (uint8 raster)                      pix_off: n entries
 *   d_streams, stream_cap        packed output; stream i occupies [stream_off[i], stream_off[i+1])
 *   stream_off                   n+1 entries written by encode, read by decode
 *   d_recon                      NULL or a buffer laid out like d_pixels (near > 0 reconstruction)
 */
int nblic_b200_encode_batch_device(nblic_b200_ctx *ctx, int n, const uint8_t *d_pixels, const uint64_t *pix_off,
                                   const int *heights, const int *widths, int near, int effort,
                                   uint8_t *d_streams, uint64_t stream_cap, uint64_t *stream_off,
                                   uint8_t *d_recon, int *status);
/* pix_cap[i] = bytes reserved for raster i at d_pixels + pix_off[i] (host array, n entries).  The raster size comes
 * from the stream header in device memory, so a stream whose height x width exceeds pix_cap[i] is reported
 * NBLIC_B200_OVERFLOW and not decoded.  NULL disables the check (trusted streams only). */
int nblic_b200_decode_batch_device(nblic_b200_ctx *ctx, int n, const uint8_t *d_streams, const uint64_t *stream_off,
                                   uint8_t *d_pixels, const uint64_t *pix_off, const uint64_t *pix_cap, int *status);

/* Deterministic synthetic gray image written to device memory; occluders is
 * the 12 x 5 int32 table {cx, cy, rad, off, kind} (nblic_image_compression_b200/synth.py). */
int nblic_b200_synth_gray(nblic_b200_ctx *ctx, uint8_t *d_out, int height, int width, uint32_t seed,
                          const int32_t *occluders);

/* n equal-size images, seeds seed0 + k * seed_stride, packed back to back at d_out; occluders = n tables of 12 x 5 int32. */
int nblic_b200_synth_gray_batch(nblic_b200_ctx *ctx, uint8_t *d_out, int n, int height, int width, uint32_t seed0,
                                uint32_t seed_stride, const int32_t *occluders);

/* Test hook: out[i] = trunc(num[i] / den[i]) computed by the kernels' reciprocal-based exact 64-bit
 * division (csrc/coop_avp.cuh: div_rcp); den[i] == 0 yields 0.  Host arrays. */
int nblic_b200_debug_divcheck(nblic_b200_ctx *ctx, const int64_t *num, const int64_t *den, int n, int64_t *out);

/* Instrumentation for bench.py: kernels launched by this context so far, and the CUDA-event
 * duration (ms, on ctx's stream) of the coder kernel of the most recent batch call. */
uint64_t nblic_b200_launch_count(const nblic_b200_ctx *ctx);
float nblic_b200_last_coder_ms(const nblic_b200_ctx *ctx);
/* coder streams the most recent cooperative launch could hold resident at once (fill = streams / slots) */
int nblic_b200_last_slots(const nblic_b200_ctx *ctx);
/* The CUDA stream (cudaStream_t) every call of this context issues its work on, for callers that
 * want to bracket calls with their own CUDA events. */
void *nblic_b200_stream_handle(const nblic_b200_ctx *ctx);
/* Name of the mapping the most recent batch call actually used ("warp" / "lane"). */
const char *nblic_b200_last_mapping(const nblic_b200_ctx *ctx);
/* Library version string. */
const char *nblic_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* NBLIC_B200_H */
