/*
 * oracle/bce_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement (plain C) of the reference's compression front end and of
 * the host-side archive writer, used ONLY as the checker by tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * The product library (libbce_gpu.so) never links or calls anything here.
 *
 * Pinning: the reference ships no tests or golden vectors.  The
 * restatement is pinned against (a) the three known-answer archives that the
 * unmodified reference produced (tests/golden/kat.json) and
 * (b) the unmodified reference itself, compiled from /root/reference/bce.cpp
 * into oracle/_ref/ by oracle/Makefile (tests/test_oracle_vs_ref.py).
 *
 * Every function cites the reference lines it follows (bce.cpp:line).
 */
#ifndef BCE_ORACLE_H
#define BCE_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { uint32_t sym, k, c1, c2, cs; } bceo_tuple;   /* args of coder.set, bce.cpp:1302 */

typedef struct {
  bceo_tuple *tuples[8];     /* per stream (= wavelet level), in (round, position) order */
  size_t count[8];
  size_t cap[8];
  uint32_t C[8];             /* C[i] = zeros of level (i+7)%8, bce.cpp:1128 */
  uint64_t visits[8];        /* node visits per level (n-1 each for primitive input) */
  uint64_t rounds;           /* iterations of the do..while, bce.cpp:1246-1371 */
  uint64_t peak_frontier;    /* max over rounds of the total queue length */
} bceo_cse_result;

/* words per level of the rank dictionary: n/32 + 1 (bce.cpp:135) */
static inline size_t bceo_rank_words(uint32_t n) { return (size_t)n / 32 + 1; }

/* File::rotate, bce.cpp:858-894: index i of the least rotation (smallest i on ties). */
uint32_t bceo_least_rotation(const uint8_t *T, size_t n);

/* File::rotate + File::bwt, bce.cpp:858-910: cyclic BWT L[0..n) of T and offset.
 * SA_out (optional, n entries): start index in T of the rotation in row r.
 * For non-primitive T the order of identical rotations in SA_out is ascending index. */
int bceo_bwt(const uint8_t *T, uint32_t n, uint8_t *L, uint32_t *offset, uint32_t *SA_out);

/* RankFile ctor, bce.cpp:944-972 + Rank::build :138-145.
 * ranks[j] must hold bceo_rank_words(n) words; word = [32 data bits | 32-bit rank]. */
void bceo_wavelet(const uint8_t *L, uint32_t n, uint64_t *const ranks[8]);

/* Rank::get<1>, bce.cpp:147-151 */
uint32_t bceo_rank1(const uint64_t *rank, uint32_t index);

/* BCE::encode roots :1124-1130 + BCE::code(mode=1) :1236-1373. */
int bceo_cse(const uint64_t *const ranks[8], uint32_t n, bceo_cse_result *out);
void bceo_cse_free(bceo_cse_result *r);

/* AdaptiveCoder<31> (encode side) :484-724, VCoder :362-378, header/concat :1134-1157.
 * cfg: 9 rows x 32 context-bit bytes (NULL = default table :713-724).
 * *words is malloc'd (caller frees with bceo_free). */
int bceo_encode_archive(const bceo_cse_result *cse, uint32_t n, uint32_t offset,
                        const uint8_t *cfg, uint16_t **words, size_t *nwords);

/* whole `bce -c` pipeline on a memory buffer (main :1403-1427 minus file I/O) */
int bceo_compress(const uint8_t *T, uint32_t n, const uint8_t *cfg,
                  uint16_t **words, size_t *nwords);

/* unbwt::bitwise :999-1038 (needs no suffix sorter) */
void bceo_unbwt_bitwise(const uint64_t *const ranks[8], uint32_t offset, uint32_t n, uint8_t *out);
/* unbwt::bytewise :1043-1103 (wavelet -> bytes, inverse_bw_transform idx=1, rotate) */
int bceo_unbwt_bytewise(const uint64_t *const ranks[8], uint32_t offset, uint32_t n, uint8_t *out);
/* first half of unbwt::bytewise :1050-1085 only: the BWT bytes back from the 8 levels */
void bceo_wavelet_to_bytes(const uint64_t *const ranks[8], uint32_t n, uint8_t *L);

/* the default 9x32 config table, bce.cpp:713-724 */
const uint8_t *bceo_default_config(void);

/* Checksums used to pin streams too large to store (tests/golden/big_vectors.json):
 *   calls  (count x 5 words s,k,c1,c2,cs): h_j = ((((s*A + k)*A + c1)*A + c2)*A + cs), A = 0x9E3779B97F4A7C15
 *   words  (count packed words):           h_j = w_j
 * out[0] = SUM h_j, out[1] = SUM h_j * (2 (first + j) + 1), mod 2^64 -- the same sums oracle/ref_tap.cpp keeps
 * per stream of the unmodified reference and bce_gpu_resident_checksum returns for the device. */
void bceo_call_checksum(const uint32_t *tuples, size_t count, uint64_t first, uint64_t out[2]);
void bceo_word_checksum(const uint32_t *words, size_t count, uint64_t first, uint64_t out[2]);

void bceo_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
