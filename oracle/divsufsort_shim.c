/*
 * oracle/divsufsort_shim.c -- TEST INFRASTRUCTURE, not product code.
 *
 * From-scratch implementation of the three libdivsufsort entry points declared
 * in oracle/divsufsort.h, used (a) to link the UNMODIFIED reference
 * (/root/reference/bce.cpp:36, :901, :1091) into oracle/_ref/, and (b) by the
 * C restatement in bce_oracle.c.  The suffix sorter is SA-IS (Nong, Zhang,
 * Chan 2009): classify S/L types, induce-sort LMS substrings, name them,
 * recurse on the reduced string if names collide, induce the final order.
 * The end-of-string sentinel is virtual (smaller than every symbol).
 */
#include "divsufsort.h"

#include <stdlib.h>
#include <string.h>

typedef int32_t sa_t;

typedef struct {
  const void *text;
  int width;            /* 1 = bytes, 4 = int32 symbols (recursion levels) */
  sa_t n;
  sa_t sigma;           /* alphabet size */
  uint8_t *stype;       /* bit i set  <=>  suffix i is S-type */
} sais_level;

static inline sa_t sym(const sais_level *lv, sa_t i) {
  return lv->width == 1 ? (sa_t)((const uint8_t *)lv->text)[i]
                        : ((const sa_t *)lv->text)[i];
}
static inline int is_s(const sais_level *lv, sa_t i) {
  return (lv->stype[i >> 3] >> (i & 7)) & 1;
}
static inline int is_lms(const sais_level *lv, sa_t i) {
  return i > 0 && is_s(lv, i) && !is_s(lv, i - 1);
}

static void bucket_bounds(const sa_t *cnt, sa_t *bkt, sa_t sigma, int ends) {
  sa_t run = 0;
  for (sa_t c = 0; c < sigma; ++c) {
    run += cnt[c];
    bkt[c] = ends ? run : run - cnt[c];
  }
}

/* One full induction sweep: L-types left to right, then S-types right to left.
 * On entry SA holds LMS suffixes at their bucket ends and -1 elsewhere. */
static void induce(const sais_level *lv, sa_t *SA, const sa_t *cnt, sa_t *bkt) {
  const sa_t n = lv->n;
  bucket_bounds(cnt, bkt, lv->sigma, 0);
  /* the virtual sentinel suffix is the smallest: it induces suffix n-1 (always L) */
  SA[bkt[sym(lv, n - 1)]++] = n - 1;
  for (sa_t i = 0; i < n; ++i) {
    sa_t j = SA[i];
    if (j > 0 && !is_s(lv, j - 1)) SA[bkt[sym(lv, j - 1)]++] = j - 1;
  }
  bucket_bounds(cnt, bkt, lv->sigma, 1);
  for (sa_t i = n - 1; i >= 0; --i) {
    sa_t j = SA[i];
    if (j > 0 && is_s(lv, j - 1)) SA[--bkt[sym(lv, j - 1)]] = j - 1;
  }
}

/* are the LMS substrings starting at p and q identical (symbols and types)? */
static int lms_equal(const sais_level *lv, sa_t p, sa_t q) {
  const sa_t n = lv->n;
  for (sa_t d = 0;; ++d) {
    sa_t a = p + d, b = q + d;
    if (a >= n || b >= n) return 0;          /* one of them runs into the sentinel */
    if (sym(lv, a) != sym(lv, b) || is_s(lv, a) != is_s(lv, b)) return 0;
    if (d > 0) {
      int ea = is_lms(lv, a), eb = is_lms(lv, b);
      if (ea || eb) return ea && eb;
    }
  }
}

static int sais_run(const void *text, sa_t *SA, sa_t n, sa_t sigma, int width) {
  if (n <= 0) return 0;
  if (n == 1) { SA[0] = 0; return 0; }

  sais_level lv;
  lv.text = text; lv.width = width; lv.n = n; lv.sigma = sigma;
  lv.stype = (uint8_t *)calloc(((size_t)n >> 3) + 1, 1);
  sa_t *cnt = (sa_t *)calloc((size_t)sigma, sizeof(sa_t));
  sa_t *bkt = (sa_t *)malloc((size_t)sigma * sizeof(sa_t));
  if (!lv.stype || !cnt || !bkt) { free(lv.stype); free(cnt); free(bkt); return -2; }

  /* classification, right to left; position n-1 is L because of the sentinel */
  for (sa_t i = n - 2; i >= 0; --i) {
    sa_t a = sym(&lv, i), b = sym(&lv, i + 1);
    if (a < b || (a == b && is_s(&lv, i + 1))) lv.stype[i >> 3] |= (uint8_t)(1u << (i & 7));
  }
  for (sa_t i = 0; i < n; ++i) cnt[sym(&lv, i)]++;

  /* pass 1: sort the LMS substrings */
  for (sa_t i = 0; i < n; ++i) SA[i] = -1;
  bucket_bounds(cnt, bkt, sigma, 1);
  sa_t m = 0;
  for (sa_t i = 1; i < n; ++i)
    if (is_lms(&lv, i)) { SA[--bkt[sym(&lv, i)]] = i; ++m; }
  induce(&lv, SA, cnt, bkt);

  int rc = 0;
  if (m > 0) {
    sa_t *order = (sa_t *)malloc((size_t)m * sizeof(sa_t));   /* LMS starts, substring-sorted */
    sa_t *name_of = (sa_t *)malloc(((size_t)n / 2 + 1) * sizeof(sa_t)); /* by start/2 */
    sa_t *reduced = (sa_t *)malloc((size_t)m * sizeof(sa_t));
    sa_t *starts = (sa_t *)malloc((size_t)m * sizeof(sa_t));
    sa_t *SA1 = (sa_t *)malloc((size_t)m * sizeof(sa_t));
    if (!order || !name_of || !reduced || !starts || !SA1) {
      free(order); free(name_of); free(reduced); free(starts); free(SA1);
      free(lv.stype); free(cnt); free(bkt);
      return -2;
    }
    sa_t k = 0;
    for (sa_t i = 0; i < n; ++i)
      if (SA[i] > 0 && is_lms(&lv, SA[i])) order[k++] = SA[i];

    sa_t names = 0;
    for (sa_t i = 0; i < m; ++i) {
      if (i == 0 || !lms_equal(&lv, order[i - 1], order[i])) ++names;
      name_of[order[i] >> 1] = names - 1;
    }
    k = 0;
    for (sa_t i = 1; i < n; ++i)
      if (is_lms(&lv, i)) { starts[k] = i; reduced[k] = name_of[i >> 1]; ++k; }

    if (names < m) {
      rc = sais_run(reduced, SA1, m, names, 4);
    } else {
      for (sa_t i = 0; i < m; ++i) SA1[reduced[i]] = i;
    }

    /* pass 2: seed with fully sorted LMS suffixes, induce everything */
    if (rc == 0) {
      for (sa_t i = 0; i < n; ++i) SA[i] = -1;
      bucket_bounds(cnt, bkt, sigma, 1);
      for (sa_t i = m - 1; i >= 0; --i) {
        sa_t p = starts[SA1[i]];
        SA[--bkt[sym(&lv, p)]] = p;
      }
      induce(&lv, SA, cnt, bkt);
    }
    free(order); free(name_of); free(reduced); free(starts); free(SA1);
  }
  free(lv.stype); free(cnt); free(bkt);
  return rc;
}

saint_t divsufsort(const sauchar_t *T, saidx_t *SA, saidx_t n) {
  if (!T || !SA || n < 0) return -1;
  return sais_run(T, SA, n, 256, 1);
}

saidx_t divbwt(const sauchar_t *T, sauchar_t *U, saidx_t *A, saidx_t n) {
  (void)A;                       /* the reference passes NULL (bce.cpp:901) */
  if (!T || !U || n < 0) return -1;
  if (n <= 1) { if (n == 1) U[0] = T[0]; return n; }

  saidx_t *SA = (saidx_t *)malloc((size_t)n * sizeof(saidx_t));
  sauchar_t *out = (sauchar_t *)malloc((size_t)n);
  if (!SA || !out) { free(SA); free(out); return -2; }
  if (sais_run(T, SA, n, 256, 1) != 0) { free(SA); free(out); return -2; }

  saidx_t primary = -1, w = 1;
  out[0] = T[n - 1];
  for (saidx_t i = 0; i < n; ++i) {
    if (SA[i] == 0) primary = i + 1;
    else out[w++] = T[SA[i] - 1];
  }
  memcpy(U, out, (size_t)n);     /* T and U may be the same buffer */
  free(SA); free(out);
  return primary;
}

saint_t inverse_bw_transform(const sauchar_t *T, sauchar_t *U, saidx_t *A,
                             saidx_t n, saidx_t idx) {
  (void)A;
  if (!T || !U || n < 0 || idx < 0 || n < idx || (n > 0 && idx == 0)) return -1;
  if (n <= 1) { if (n == 1) U[0] = T[0]; return 0; }

  /* Rows 0..n of the sentinel matrix; row idx ends in the sentinel and is the
   * one divbwt dropped.  step[u] = row reached by an LF step from the row that
   * holds BWT entry u. */
  saidx_t *step = (saidx_t *)malloc((size_t)n * sizeof(saidx_t));
  sauchar_t *out = (sauchar_t *)malloc((size_t)n);
  if (!step || !out) { free(step); free(out); return -2; }
  saidx_t first[256], seen[256];
  memset(first, 0, sizeof first);
  memset(seen, 0, sizeof seen);
  for (saidx_t u = 0; u < n; ++u) first[T[u]]++;
  saidx_t run = 1;                                   /* row 0 starts with the sentinel */
  for (int c = 0; c < 256; ++c) { saidx_t t = first[c]; first[c] = run; run += t; }
  for (saidx_t u = 0; u < n; ++u) step[u] = first[T[u]] + seen[T[u]]++;

  saidx_t row = 0;
  for (saidx_t k = n - 1; k >= 0; --k) {
    saidx_t u = row < idx ? row : row - 1;
    if (row == idx || u < 0 || u >= n) { free(step); free(out); return -1; }
    out[k] = T[u];
    row = step[u];
  }
  memcpy(U, out, (size_t)n);
  free(step); free(out);
  return 0;
}
