/*
 * oracle/divsufsort.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Stand-in for the libdivsufsort API that the reference includes at
 * bce.cpp:36 and calls at bce.cpp:901 (divbwt) and bce.cpp:1091
 * (inverse_bw_transform).  libdivsufsort (akamiru/libdivsufsort, a fork of
 * y-256/libdivsufsort; no version pinned by the reference)
 * is not vendored in /root/reference and not installed, so the unmodified
 * reference cannot link without something that provides these two symbols.
 *
 * This header and divsufsort_shim.c are written from the published
 * *semantics* of those two calls (a suffix array of a string is unique, so
 * any correct suffix sorter yields the same divbwt output):
 *
 *   divbwt(T, U, A, n)  -> U[0] = T[n-1], followed by T[SA[i]-1] for every
 *                          i with SA[i] != 0, in suffix order; returns the
 *                          primary index  (rank of suffix 0) + 1.
 *                          T and U may alias.  n <= 1: copies, returns n.
 *   inverse_bw_transform(T, U, A, n, idx) -> inverse of the above.
 *
 * The sorter behind it is a plain SA-IS (induced sorting), O(n).
 * Nothing under oracle/ may be linked into the product library.
 */
#ifndef BCE_ORACLE_DIVSUFSORT_H
#define BCE_ORACLE_DIVSUFSORT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint8_t sauchar_t;
typedef int32_t saint_t;
typedef int32_t saidx_t;

/* suffix array of T[0..n) (plain suffixes, shorter-is-smaller). 0 on success. */
saint_t divsufsort(const sauchar_t *T, saidx_t *SA, saidx_t n);

/* BWT with primary index, upstream conventions (see header comment). */
saidx_t divbwt(const sauchar_t *T, sauchar_t *U, saidx_t *A, saidx_t n);

/* inverse of divbwt. 0 on success, -1 on bad arguments. */
saint_t inverse_bw_transform(const sauchar_t *T, sauchar_t *U, saidx_t *A,
                             saidx_t n, saidx_t idx);

#ifdef __cplusplus
}
#endif

#endif
