// cse_probe.cuh -- EXPERIMENT BUILDS ONLY (BCE_GPU_EXPERIMENTS): how fast is one wide round when no warp ever waits
// for another?  Every warp takes chunks of 32*ITEMS nodes of the flat frontier, computes them exactly as
// cse_wide_kernel does and writes children and counts into per-chunk slots (fixed places: no prefix sums, no chained
// scan, no staging, no barriers).  The output is not a valid frontier for the shipped kernels; the run stops after
// the timed round (BCE_GPU_CSE_DBG_ROUND=r BCE_GPU_CSE_DBG_FLAGS=8|16|32: 1, 2 or 4 nodes per lane).
#pragma once

namespace bce {

template <int ITEMS, int MINB>
__global__ void __launch_bounds__(256, MINB) cse_slot_probe_kernel(CseArgs a) {
  constexpr uint32_t CH = 32 * ITEMS;
  __shared__ uint32_t s_first[17];                    // first chunk of (level, half), 16 entries + total
  __shared__ uint32_t s_cnt[8][2];
  CseDeviceState* S = a.st;
  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  const uint32_t round = S->round;
  const int cur = round & 1, nxt = cur ^ 1;
  if (tid < 16) s_cnt[tid >> 1][tid & 1] = S->cnt[cur][tid >> 1][tid & 1];
  __syncthreads();
  if (tid == 0) {
    uint32_t t = 0;
    for (int i = 0; i < 16; ++i) { s_first[i] = t; t += (s_cnt[i >> 1][i & 1] + CH - 1) / CH; }
    s_first[16] = t;
  }
  __syncthreads();
  const uint32_t total = s_first[16];
  uint32_t* const counts = reinterpret_cast<uint32_t*>(a.desc);
  for (uint32_t t = blockIdx.x * 8u + warp; t < total; t += gridDim.x * 8u) {
    int lh = 0;
#pragma unroll
    for (int i = 1; i < 16; ++i)
      if (t >= s_first[i]) lh = i;
    const int l = lh >> 1, h = lh & 1, ln = (l + 1) & 7;
    const uint32_t q = t - s_first[lh];
    const uint32_t count = s_cnt[l][h];
    const uint32_t first = q * CH + lane;
    const uint32_t* __restrict__ fs = a.fs[cur][l];
    const uint32_t* __restrict__ fa = a.fa[cur][l];
    const uint32_t* __restrict__ fb = a.fb[cur][l];
    const uint64_t* __restrict__ R = a.ranks[l];
    uint32_t ns[ITEMS], na[ITEMS], nb[ITEMS];
    uint64_t wa[ITEMS], wb[ITEMS], wc[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const uint32_t idx = first + 32u * j;
      ns[j] = na[j] = nb[j] = 0;
      if (idx < count) {
        const uint32_t at = h ? a.cap - 1u - idx : idx;
        ns[j] = __ldcg(fs + at); na[j] = __ldcg(fa + at); nb[j] = __ldcg(fb + at);
      }
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      wa[j] = wb[j] = wc[j] = 0;
      if (first + 32u * j < count) {
        wa[j] = __ldg(R + (ns[j] >> 5));
        wb[j] = __ldg(R + ((ns[j] + na[j] + nb[j]) >> 5));
        wc[j] = __ldg(R + ((ns[j] + na[j]) >> 5));
      }
    }
    // slots: chunk q of (l, h) owns CH places for zero-children, CH for one-children, 2 CH emission words
    const uint32_t slot = (s_first[lh] - s_first[l * 2] + q) * CH;          // chunk index inside the level
    uint32_t* __restrict__ gs = a.fs[nxt][ln];
    uint32_t* __restrict__ ga = a.fa[nxt][ln];
    uint32_t* __restrict__ gb = a.fb[nxt][ln];
    const uint32_t zbase = slot, obase = a.cap / 2 + slot;
    uint32_t* __restrict__ ew = a.emit[l] + size_t(slot) * 3u;
    const uint32_t one_base = a.C[ln];
    uint32_t cz = 0, co = 0, ce = 0;
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      bool fz = false, fo = false;
      uint32_t zs_ = 0, za_ = 0, zb_ = 0, os_ = 0, oa_ = 0, ob_ = 0, e0 = 0, e1 = 0, e2 = 0, nw = 0;
      const uint32_t x0 = na[j], x1 = nb[j], x = x0 + x1;
      if (first + 32u * j < count) {
        const uint32_t s = ns[j];
        const uint32_t s1 = rank1_word(wa[j], s);
        const uint32_t c1 = rank1_word(wb[j], s + x) - s1;
        const uint32_t s0 = s - s1;
        const uint32_t z0 = (s + x0 - rank1_word(wc[j], s + x0)) - s0;
        zs_ = s0;
        os_ = one_base + s1;
        if (c1 == 0) { fz = true; za_ = x0; zb_ = x1; }
        else if (c1 == x) { fo = true; oa_ = x0; ob_ = x1; }
        else {
          const uint32_t c0 = x - c1;
          const uint32_t lo = x0 > c1 ? x0 - c1 : 0u;
          const uint32_t hi = x0 - (c1 > x1 ? c1 - x1 : 0u);
          const uint32_t z1 = c0 - z0, o1 = x1 - z1, o0c = c1 - o1;
          if (hi != lo) nw = count_words(a, l, z0 - lo, hi - lo + 1, c0, x1, x, e0, e1, e2);
          if (z0 && z1) { fz = true; za_ = z0; zb_ = z1; }
          if (o0c && o1) { fo = true; oa_ = o0c; ob_ = o1; }
        }
      }
      const unsigned bz = __ballot_sync(0xffffffffu, fz);
      const unsigned bo = __ballot_sync(0xffffffffu, fo);
      const unsigned be = __ballot_sync(0xffffffffu, nw != 0u);
      if (fz) { const uint32_t p = zbase + cz + __popc(bz & lt_mask); if (p < a.cap) { gs[p] = zs_; ga[p] = za_; gb[p] = zb_; } }
      if (fo) { const uint32_t p = obase + co + __popc(bo & lt_mask); if (p < a.cap) { gs[p] = os_; ga[p] = oa_; gb[p] = ob_; } }
      cz += __popc(bz);
      co += __popc(bo);
      if (be) {
        const unsigned b3 = __ballot_sync(0xffffffffu, nw == 3u);
        if (nw && nw <= 3u) put_words(ew + ce + __popc(be & lt_mask) + 2u * __popc(b3 & lt_mask), nw, e0, e1, e2, x1, x);
        ce += __popc(be) + 2u * __popc(b3);
      }
    }
    if (lane == 0 && t < 3u * a.desc_tiles) counts[t] = cz | (co << 8) | (ce << 16);
  }
}

}  // namespace bce
