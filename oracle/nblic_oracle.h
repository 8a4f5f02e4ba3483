/*
 * oracle/nblic_oracle.h -- CPU restatement of the NBLIC / QNBLIC codecs (TEST INFRASTRUCTURE).
 *
 * This is the parity oracle for the B200 kernels: a plain-C re-expression of the algorithm in
 * /root/reference/src/NBLIC.c and /root/reference/src/QNBLIC.c, written from the behaviour of those
 * files (each function cites the reference lines it follows).  It is NOT part of the product path:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity is PINNED: tests/test_oracle.py checks this oracle byte-for-byte against the committed
 * golden vectors under tests/golden/ (produced by the unmodified reference, see
 * tests/golden/make_golden.py) and, when oracle/_ref/libnblic_ref.so exists, live against the
 * reference itself on random and edge-case inputs.
 */
#ifndef NBLIC_ORACLE_H
#define NBLIC_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* QNBLIC ("Q0.2", effort 0, lossless).  Returns the number of uint16 words written, or -1. */
int oracle_q_encode(const uint8_t *img, int height, int width, uint16_t *out_words);
/* Returns 0 / -1.  in_words_avail = number of readable words (reads past it return 0). */
int oracle_q_decode(const uint16_t *in_words, long in_words_avail, uint8_t *img, int *height, int *width);

/* NBLIC ("NBLIC0.3", effort 1..3, near 0..9).  img is overwritten with the reconstruction when
 * near>0 (NBLIC.c:876,916).  *near / *effort are clipped in place.  Returns bytes written or -1. */
int oracle_n_encode(uint8_t *img, int height, int width, int *near, int *effort, uint8_t *out);
/* Returns 0 / -1. */
int oracle_n_decode(const uint8_t *in, long in_avail, uint8_t *img, int *height, int *width, int *near, int *effort);

#ifdef __cplusplus
}
#endif
#endif
