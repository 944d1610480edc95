// local_sort.cuh -- the sort of a doubling round >= 1 without radix passes.  Included by suffix_sort.cu.
//
// After round 0 the working set is ordered by group, and a round only has to order every group by
// the rank of its rotations' second halves: key = (dense group id << 32) | rank, group ids ascending
// in slot order.  Groups are short (a few hundred rotations on text), so almost every group lies
// inside one 4096-slot tile, and a tile can be sorted where it is:
//
//   ls_classify_kernel   per tile: how many slots at its front / back belong to a group that crosses
//                        the tile boundary ("non-local" slots; binary search on the sorted group ids)
//   ls_scan_kernel       exclusive scan of those counts over the tiles
//   ls_sort_kernel       loads the tile, bitonic-sorts one 64-bit word per slot (relative group id, second, slot)
//                        -- 16 per thread: strides below 16 in registers, strides inside a warp by shuffles,
//                        the six widest exchanges through shared memory --, writes it back; copies the
//                        non-local slots, in slot order, to the fall-back list
//   (radix sort of the fall-back list, a few per cent of the working set)
//   ls_place_kernel      k-th sorted fall-back element -> k-th fall-back slot (group ids ascend in both)
//
// One read and one write of the working set and ~14 warp instructions per element instead of 8 radix
// passes at ~3.8 each.  Ties may come out in any order (they are told apart by later rounds).
#pragma once

namespace bce {

constexpr int LS_THREADS = 256;
constexpr int LS_ITEMS = 16;
constexpr int LS_TILE = LS_THREADS * LS_ITEMS;      // 4096

// counts of non-local slots at the front (pre) and back (suf) of every tile
__global__ void __launch_bounds__(256) ls_classify_kernel(const uint32_t* __restrict__ gdv, uint32_t m, uint32_t tiles,
                                                          uint32_t* __restrict__ pre, uint32_t* __restrict__ suf,
                                                          uint32_t* __restrict__ cnt) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= tiles) return;
  const uint32_t start = t * uint32_t(LS_TILE), end = min(m, start + uint32_t(LS_TILE));
  auto gd = [&](uint32_t q) { return gdv[q]; };
  uint32_t p = 0, s = 0;
  if (start > 0 && gd(start - 1) == gd(start)) {      // first group continues from the previous tile
    const uint32_t g = gd(start);
    uint32_t lo = start, hi = end;                    // first slot in [start, end) with gd > g
    while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (gd(mid) <= g) lo = mid + 1; else hi = mid; }
    p = lo - start;
  }
  if (end < m && gd(end) == gd(end - 1)) {            // last group continues into the next tile
    const uint32_t g = gd(end - 1);
    uint32_t lo = start, hi = end;                    // first slot in [start, end) with gd >= g
    while (lo < hi) { const uint32_t mid = lo + (hi - lo) / 2; if (gd(mid) < g) lo = mid + 1; else hi = mid; }
    s = end - lo;
  }
  if (p + s > end - start) { p = end - start; s = 0; }   // one group covers the whole tile
  pre[t] = p;
  suf[t] = s;
  cnt[t] = p + s;
}

// exclusive scan of cnt over the tiles (one CTA; a few hundred thousand values at most); total -> *total
__global__ void __launch_bounds__(1024) ls_scan_kernel(const uint32_t* __restrict__ cnt, uint32_t tiles,
                                                       uint32_t* __restrict__ off, uint32_t* __restrict__ total) {
  __shared__ uint32_t s_scan[32];
  __shared__ uint32_t s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < tiles; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < tiles ? cnt[i] : 0u;
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan<uint32_t, 1024>(v, s_scan, tot);
    const uint32_t carry = s_carry;
    if (i < tiles) off[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) s_carry = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

struct LocalSortArgs {
  const uint32_t* vin;      // rotation start of every slot
  const uint32_t* gd;       // dense group id of every slot (ascending)
  const uint32_t* rnk;      // rank array: second = rnk[(idx + h) mod n], or idx itself in the tie-break round
  uint32_t n, h;
  int tiebreak;
  uint64_t* kout;
  uint32_t* vout;
  uint32_t m;
  const uint32_t* pre;
  const uint32_t* suf;
  const uint32_t* off;
  uint64_t* fb_key;         // fall-back list (slot order)
  uint32_t* fb_idx;
  uint32_t* fb_slot;
};

// compare-exchange without branches: afterwards a <= b when up, a >= b otherwise
__device__ __forceinline__ void ls_cmpx(uint64_t& a, uint64_t& b, bool up) {
  const bool sw = (a > b) == up;
  const uint64_t na = sw ? b : a, nb = sw ? a : b;
  a = na;
  b = nb;
}

// What is sorted is ONE 64-bit word per slot: (group id relative to the tile's first << 44) | (second << 12)
// | slot in the tile.  Dense group ids grow by at most one per slot, so the relative id fits 12 bits; the
// slot makes the words distinct and finds the rotation index again afterwards (it stays in shared memory).
__global__ void __launch_bounds__(LS_THREADS) ls_sort_kernel(LocalSortArgs a) {
  __shared__ uint64_t s_key[LS_TILE];
  __shared__ uint32_t s_val[LS_TILE];
  const unsigned tid = threadIdx.x;
  const uint32_t tile = blockIdx.x;
  const uint32_t start = tile * uint32_t(LS_TILE);
  const uint32_t valid = min(uint32_t(LS_TILE), a.m - start);
  const uint32_t gd0 = a.gd[start];

  // coalesced load into shared memory; the non-local slots go to the fall-back list on the way
  {
    const uint32_t p = a.pre[tile], sfx = a.suf[tile], o = a.off[tile];
    for (uint32_t j = tid; j < uint32_t(LS_TILE); j += LS_THREADS) {
      const bool in = j < valid;
      const uint32_t v = in ? a.vin[start + j] : 0u;
      const uint32_t g = in ? a.gd[start + j] : 0u;
      // the key of the doubling step is made here (K4's gather): 16 independent gathers per thread in flight,
      // and they overlap with the sorting networks of the other tiles on the SM
      uint32_t second = v;
      if (in && !a.tiebreak) {
        uint64_t jj = uint64_t(v) + a.h;
        if (jj >= a.n) jj -= a.n;
        second = a.rnk[jj];
      }
      const uint64_t k = (uint64_t(g) << 32) | second;
      // padding sorts to the end of the tile
      s_key[j] = in ? (uint64_t(g - gd0) << 44) | (uint64_t(second) << 12) | j : ~0ull;
      s_val[j] = v;
      if (in) {
        uint32_t f = 0xFFFFFFFFu;
        if (j < p) f = o + j;
        else if (j >= valid - sfx) f = o + p + (j - (valid - sfx));
        if (f != 0xFFFFFFFFu) { a.fb_key[f] = k; a.fb_idx[f] = v; a.fb_slot[f] = start + j; }
      }
    }
  }
  __syncthreads();
  // thread t holds elements 16 t .. 16 t + 15
  uint64_t k[LS_ITEMS];
#pragma unroll
  for (int r = 0; r < LS_ITEMS; ++r) k[r] = s_key[tid * LS_ITEMS + r];
  __syncthreads();                                   // s_key is reused for the exchanges between warps

  // bitonic sort, ascending over the element index i = 16 tid + r.
  // (1) kk = 2 .. 16: every partner is inside the thread
#pragma unroll
  for (int lk = 1; lk <= 4; ++lk) {
    const int kk = 1 << lk;
#pragma unroll
    for (int lj = lk - 1; lj >= 0; --lj) {
      const int j = 1 << lj;
#pragma unroll
      for (int r = 0; r < LS_ITEMS; ++r)
        if ((r & j) == 0) ls_cmpx(k[r], k[r | j], kk < LS_ITEMS ? ((r & kk) == 0) : ((tid & 1u) == 0));
    }
  }
  // (2) kk = 32 .. 4096: strides of 16 and more pair whole threads (partner thread tid ^ tj, same r): inside a
  //     warp by shuffles, between warps through shared memory; strides below 16 are inside the thread again.
  //     These loops stay loops: unrolled, the kernel was 30 K instructions and starved on instruction fetch.
#pragma unroll 1
  for (uint32_t kk = 32; kk <= uint32_t(LS_TILE); kk <<= 1) {
    const bool up = ((tid * LS_ITEMS) & kk) == 0;
#pragma unroll 1
    for (uint32_t tj = kk / (2 * LS_ITEMS); tj > 0; tj >>= 1) {
      const bool keep_min = ((tid & tj) == 0) == up;          // lower index of the pair keeps the min when ascending
      if (tj < 32) {
#pragma unroll
        for (int r = 0; r < LS_ITEMS; ++r) {
          const uint64_t o = __shfl_xor_sync(0xffffffffu, k[r], tj);
          k[r] = (keep_min == (o < k[r])) ? o : k[r];         // distinct words: o < k or o > k
        }
      } else {
        // element r of thread t at r * 256 + t: conflict free on both sides
#pragma unroll
        for (int r = 0; r < LS_ITEMS; ++r) s_key[r * LS_THREADS + tid] = k[r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < LS_ITEMS; ++r) {
          const uint64_t o = s_key[r * LS_THREADS + (tid ^ tj)];
          k[r] = (keep_min == (o < k[r])) ? o : k[r];
        }
        __syncthreads();
      }
    }
#pragma unroll
    for (int lj = 3; lj >= 0; --lj) {
      const int j = 1 << lj;
#pragma unroll
      for (int r = 0; r < LS_ITEMS; ++r)
        if ((r & j) == 0) ls_cmpx(k[r], k[r | j], up);
    }
  }
  // back through shared memory for a coalesced store; the slot in the word fetches the rotation index
#pragma unroll
  for (int r = 0; r < LS_ITEMS; ++r) s_key[tid * LS_ITEMS + r] = k[r];
  __syncthreads();
  for (uint32_t j = tid; j < valid; j += LS_THREADS) {
    const uint64_t w = s_key[j];
    a.kout[start + j] = (uint64_t(gd0 + uint32_t(w >> 44)) << 32) | uint32_t(w >> 12);
    a.vout[start + j] = s_val[uint32_t(w) & 4095u];
  }
}

// histogram of the rank-scatter partition digit (idx >> shift) over the working set (K4 counts it on the
// radix path; here nothing else reads idx in a persistent shape)
__global__ void __launch_bounds__(256) ls_idx_hist_kernel(const uint32_t* __restrict__ idx, uint32_t m, int shift,
                                                          uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[8][256];
  const unsigned tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 8 * 256; i += 256) (&h[0][0])[i] = 0;
  __syncthreads();
  for (uint64_t q = uint64_t(blockIdx.x) * 256 + tid; q < m; q += uint64_t(gridDim.x) * 256)
    atomicAdd(&h[warp][(idx[q] >> shift) & 255u], 1u);
  __syncthreads();
  uint32_t v = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) v += h[w][tid];
  if (v) atomicAdd(&hist[tid], v);
}

__global__ void __launch_bounds__(256) ls_place_kernel(const uint64_t* __restrict__ sk, const uint32_t* __restrict__ sv,
                                                       const uint32_t* __restrict__ slot, uint32_t count,
                                                       uint64_t* __restrict__ kout, uint32_t* __restrict__ vout) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= count) return;
  const uint32_t q = slot[p];
  kout[q] = sk[p];
  vout[q] = sv[p];
}

}  // namespace bce
