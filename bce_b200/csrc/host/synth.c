/*
 * synth.c -- deterministic synthetic inputs for the configs BASELINE.json names
 * (SURVEY.md 8d).  Integer-only, PRNG = splitmix64, so the same (kind, n, seed)
 * gives the same bytes on every host.  All generators produce primitive strings
 * with overwhelming probability (the reference does not round-trip exact powers,
 * SURVEY.md 4-6).
 *
 *   kind 0  markov2-text   order-2 Markov chain over 64 symbols (bytes 32..95), sparse rows
 *   kind 1  enwik-shaped   order-2 Markov over 96 printable symbols + copy model
 *                          (p = 1/2048 per byte, length Geom(mean 48) capped at 4096)
 *   kind 2  mixed-binary   1 MiB segments cycling: uniform bytes / skewed two-symbol (p=.95)
 *                          / zero runs U[100,3000] split by 4..64 random bytes / LE u32 counters
 *   kind 3  uniform        uniform random bytes
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t *state) {
  uint64_t z = (*state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

/* successor table of an order-2 chain: for every context 8 candidate symbols;
 * candidate j is drawn with probability 2^-(j+1) (the last two share 2^-7) */
static uint8_t *markov_table(uint32_t sigma, uint64_t seed) {
  size_t rows = (size_t)sigma * sigma;
  uint8_t *tab = (uint8_t *)malloc(rows * 8);
  if (!tab) return NULL;
  uint64_t st = seed ^ 0xA5A5A5A5DEADBEEFull;
  for (size_t r = 0; r < rows; ++r) {
    uint64_t z = splitmix64(&st);
    for (int j = 0; j < 8; ++j) tab[r * 8 + j] = (uint8_t)((z >> (8 * j)) % sigma);
  }
  return tab;
}

static inline uint32_t pick(uint64_t r) {
  uint32_t j = (uint32_t)__builtin_ctzll(r | (1ull << 63));
  return j > 7 ? 7 : j;
}

static int gen_markov(uint8_t *out, size_t n, uint64_t seed, uint32_t sigma, uint8_t base, int copies) {
  uint8_t *tab = markov_table(sigma, seed);
  if (!tab) return -2;
  uint64_t st = seed * 0x2545F4914F6CDD1Dull + 1;
  uint32_t a = 0, b = 0;
  size_t i = 0;
  while (i < n) {
    uint64_t r = splitmix64(&st);
    if (copies && i > 64 && (r >> 40 & 2047) == 0) {
      /* copy: geometric length, one 16-bit trial per step, success 1365/65536 ~ 1/48 */
      uint32_t len = 1;
      uint64_t bits = splitmix64(&st);
      int have = 4;
      while (len < 4096) {
        if (!have) { bits = splitmix64(&st); have = 4; }
        uint32_t t = (uint32_t)(bits & 0xFFFF); bits >>= 16; --have;
        if (t < 1365) break;
        ++len;
      }
      size_t from = (size_t)(splitmix64(&st) % i);
      for (uint32_t k = 0; k < len && i < n; ++k, ++i) out[i] = out[from + k];
      a = (uint32_t)(out[i - 2] - base) % sigma;
      b = (uint32_t)(out[i - 1] - base) % sigma;
      continue;
    }
    uint32_t c = tab[((size_t)a * sigma + b) * 8 + pick(r)];
    out[i++] = (uint8_t)(base + c);
    a = b; b = c;
  }
  free(tab);
  return 0;
}

static void gen_mixed(uint8_t *out, size_t n, uint64_t seed) {
  uint64_t st = seed * 0x9E3779B97F4A7C15ull + 7;
  const size_t seg = (size_t)1 << 20;
  for (size_t base = 0, s = 0; base < n; base += seg, ++s) {
    size_t len = n - base < seg ? n - base : seg;
    uint8_t *p = out + base;
    switch (s & 3) {
      case 0:
        for (size_t i = 0; i < len; i += 8) {
          uint64_t r = splitmix64(&st);
          for (size_t k = 0; k < 8 && i + k < len; ++k) p[i + k] = (uint8_t)(r >> (8 * k));
        }
        break;
      case 1:
        for (size_t i = 0; i < len; i += 4) {
          uint64_t r = splitmix64(&st);
          for (size_t k = 0; k < 4 && i + k < len; ++k)
            p[i + k] = ((r >> (16 * k)) & 0xFFFF) < 62259 ? 0x41 : 0x7A;   /* 62259/65536 = .95 */
        }
        break;
      case 2: {
        size_t i = 0;
        while (i < len) {
          uint64_t r = splitmix64(&st);
          size_t run = 100 + (size_t)(r % 2901);
          size_t gap = 4 + (size_t)((r >> 32) % 61);
          for (size_t k = 0; k < run && i < len; ++k) p[i++] = 0;
          uint64_t g = 0;
          for (size_t k = 0; k < gap && i < len; ++k) {
            if ((k & 7) == 0) g = splitmix64(&st);
            p[i++] = (uint8_t)(g >> (8 * (k & 7)));
          }
        }
        break;
      }
      default: {
        uint32_t v = (uint32_t)splitmix64(&st);
        for (size_t i = 0; i < len; i += 4, ++v)
          for (size_t k = 0; k < 4 && i + k < len; ++k) p[i + k] = (uint8_t)(v >> (8 * k));
        break;
      }
    }
  }
}

/* returns 0 on success */
int bce_synth(int kind, uint8_t *out, size_t n, uint64_t seed) {
  if (!out) return -1;
  switch (kind) {
    case 0: return gen_markov(out, n, seed, 64, 32, 0);
    case 1: return gen_markov(out, n, seed, 96, 32, 1);
    case 2: gen_mixed(out, n, seed); return 0;
    case 3: {
      uint64_t st = seed ^ 0x1234567ull;
      for (size_t i = 0; i < n; i += 8) {
        uint64_t r = splitmix64(&st);
        for (size_t k = 0; k < 8 && i + k < n; ++k) out[i + k] = (uint8_t)(r >> (8 * k));
      }
      return 0;
    }
    default: return -1;
  }
}
