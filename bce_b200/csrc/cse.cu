// cse.cu -- stage B2: the CSE level loop (Compression by Substring Enumeration).
//
// Replaces BCE::code(mode = 1) (bce.cpp:1236-1373), its pArray queues (:226-356) and the
// root set-up (:1124-1130, :1238-1240).  A node (level i, position s, x0, x1) is the interval
// [s, s + x0 + x1) of level i's bit vector; per node the reference does three rank queries,
// hands one bounded count to coder_[i].set (:1302) and pushes up to two children to level
// (i+1)%8.  Nodes are independent; what the archive depends on is only the ORDER of the
// emitted counts per stream: ascending round, ascending position.
//
// Device design: one persistent cooperative kernel runs many rounds.  Per round all 8
// levels are processed together; the frontier of a level is a flat array sorted by position
// (zero-half from the front, one-half from the back of one buffer).  A tile of 1024 nodes:
//   load nodes (128-bit), 12 independent rank-word gathers per thread, derive
//   (emit?, zero-child?, one-child?) -> block scan of the three counts -> chained scan over
//   the tiles of the level (warp-wide decoupled look-back, three carried values)
//   -> write counts, zero-children and one-children at their exact ordered positions.
// So the stable partition into the next level's halves and the emission order fall out of
// prefix sums; no atomics decide any position.  A grid-wide barrier separates rounds: a
// monotonically increasing arrival counter (all CTAs are co-resident: cooperative launch) with a
// bounded spin, so that a logic error surfaces as BCE_GPU_E_INTERNAL and never as a hung GPU.
//
// Algorithmic HBM bytes: 48 B per node visit + 20 B per emitted count.
#include <cooperative_groups.h>

#include <algorithm>

#include "ctx.h"

namespace cg = cooperative_groups;

namespace bce {

constexpr int CS_THREADS = 256;
// nodes per thread of the wide kernel: a template parameter (1, 2 or 4); fewer nodes per thread
// = fewer registers = more resident CTAs to overlap the load -> gather -> scan -> look-back chain
constexpr int CS_MAX_TILE = CS_THREADS * 4;

enum : uint32_t { kCseRunning = 0, kCseDone = 1, kCseDrain = 2, kCseOverflow = 3, kCseRunaway = 4,
                  kCseGoWide = 5, kCseGoNarrow = 6 };

// Narrow frontiers (the long tail: rounds ~ 8 x longest repeat, most of them with a handful of
// nodes) run in ONE thread-block cluster of 8 CTAs, one CTA per level, frontiers in shared
// memory, children handed to the next level's CTA through distributed shared memory, rounds
// separated by the hardware cluster barrier instead of a grid-wide one.
constexpr int NR_THREADS = 1024;
constexpr int NR_CAP = 1024;                       // nodes per level held in shared memory
constexpr uint32_t kNarrowLeave = NR_CAP / 2;      // a node has <= 2 children: the next round always fits
constexpr uint32_t kNarrowEnter = NR_CAP / 4;      // wide -> narrow once every level is at most this

struct CseDeviceState {
  uint32_t cnt[2][8][2];                 // [round parity][level][half] frontier sizes
  unsigned long long emitted[2][8];      // [round parity][level] counts in the emission buffer
  unsigned long long visits;
  unsigned long long peak_frontier;
  uint32_t round;
  uint32_t status;
  uint32_t err;                          // chained-scan watchdog
  uint32_t barrier_fail;                 // round at which the grid barrier timed out (+1), 0 = never
  unsigned long long arrivals;           // grid barrier: total CTA arrivals since cse_begin
  unsigned long long barriers;           // grid barriers completed since cse_begin (host-maintained between launches)
};

struct CseArgs {
  const uint64_t* ranks[8];
  uint32_t C[8];                         // first position of the one-half of level i (bce.cpp:1259)
  uint32_t* fs[2][8];                    // frontier, structure of arrays: position
  uint32_t* fa[2][8];                    //   x0
  uint32_t* fb[2][8];                    //   x1
  uint32_t cap;                          // nodes per frontier buffer, multiple of 4
  bce_tuple* emit[8];
  unsigned long long ecap[8];
  uint64_t* desc;                        // 3 x desc_tiles
  uint32_t desc_tiles;
  uint32_t max_rounds;
  uint32_t round_limit;                  // no input needs more than 8 n rounds: beyond it something is broken
  uint32_t use_narrow;                   // 1 = hand narrow frontiers to the cluster kernel
  uint32_t dbg;                          // timing experiments only (results become wrong): 1 no look-back, 2 no gathers, 4 no flush
  unsigned long long min_nodes, max_nodes;   // pipelined wide kernel: leave (kCseGoWide) when the frontier is outside
  CseDeviceState* st;
};

struct CseHost {
  CseArgs args;
  int grid = 0;
  bool narrow = true;                    // which kernel runs next (the root frontier is narrow)
  int items = 2;                         // nodes per thread of the wide kernel
  uint32_t last_round = 0;
  const void* wide_fn = nullptr;
  size_t wide_smem = 0;                  // dynamic shared memory of the wide kernel
  // automatic mode: 1024-node tiles while the frontier is huge, 512-node tiles below
  bool auto_items = false;
  const void* var_fn[2] = {nullptr, nullptr};      // [0] = 2 nodes per thread, [1] = 4
  size_t var_smem[2] = {0, 0};
  int var_grid[2] = {0, 0};
  unsigned long long known_nodes = 0;    // frontier size after the last launch
  uint32_t n = 0;
  size_t pinned_off[8] = {};
};

__device__ __forceinline__ uint32_t rank1_word(uint64_t w, uint32_t pos) {      // Rank::get<1>, bce.cpp:147-151
  return uint32_t(w) + __popc(uint32_t(w >> 32) & ((1u << (pos & 31u)) - 1u));
}
__device__ __forceinline__ uint32_t vol_load(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long vol_load64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Grid-wide barrier.  Every CTA adds one arrival; barrier number b (0-based, counted since
// cse_begin) is passed when the counter reaches (b + 1) * gridDim.x.  The counter only grows,
// so there is no reset race.  Returns false (and records the failure) if the others do not
// arrive within the spin budget.
__device__ __forceinline__ bool grid_barrier(CseDeviceState* S, unsigned long long index, uint32_t round) {
  __syncthreads();
  bool ok = true;
  if (threadIdx.x == 0) {
    const unsigned long long target = (index + 1) * gridDim.x;
    __threadfence();
    atomicAdd(&S->arrivals, 1ull);
    uint32_t spins = 0;
    while (vol_load64(&S->arrivals) < target) {
      if (++spins > (1u << 22)) {
        atomicExch(&S->barrier_fail, round + 1);
        atomicExch(&S->err, 2u);
        ok = false;
        break;
      }
      __nanosleep(20);
    }
    __threadfence();
  }
  __syncthreads();
  return ok;
}

__global__ void cse_init_kernel(CseArgs a, uint32_t n) {
  // roots: node (0, C[i], n - C[i]) in the zero-half of level i when both are non-zero
  // (bce.cpp:1238-1240)
  const int i = threadIdx.x;
  CseDeviceState* S = a.st;
  if (i < 8) {
    uint32_t z = a.C[i], o = n - a.C[i];
    bool root = z && o;
    if (root) { a.fs[0][i][0] = 0; a.fa[0][i][0] = z; a.fb[0][i][0] = o; }
    S->cnt[0][i][0] = root ? 1u : 0u;
    S->cnt[0][i][1] = 0;
    S->cnt[1][i][0] = S->cnt[1][i][1] = 0;
    S->emitted[0][i] = S->emitted[1][i] = 0;
  }
  if (i == 0) { S->visits = 0; S->peak_frontier = 0; S->round = 0; S->status = kCseRunning; S->err = 0;
                S->barrier_fail = 0; S->arrivals = 0; S->barriers = 0; }
}

__global__ void cse_reset_emitted_kernel(CseDeviceState* S) {
  if (threadIdx.x < 8) S->emitted[0][threadIdx.x] = S->emitted[1][threadIdx.x] = 0;
  if (threadIdx.x == 0 && S->status == kCseDrain) S->status = kCseRunning;
}

// vector load of ITEMS consecutive frontier entries; `rev` = stored back to front
template <int ITEMS>
__device__ __forceinline__ void load_items(const uint32_t* base, uint32_t first, bool rev, uint32_t cap, uint32_t (&out)[ITEMS]) {
  if constexpr (ITEMS == 4) {
    uint4 v = __ldcg(reinterpret_cast<const uint4*>(base + (rev ? cap - 4 - first : first)));
    if (rev) { out[0] = v.w; out[1] = v.z; out[2] = v.y; out[3] = v.x; }
    else { out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w; }
  } else if constexpr (ITEMS == 2) {
    uint2 v = __ldcg(reinterpret_cast<const uint2*>(base + (rev ? cap - 2 - first : first)));
    if (rev) { out[0] = v.y; out[1] = v.x; } else { out[0] = v.x; out[1] = v.y; }
  } else {
    out[0] = __ldcg(base + (rev ? cap - 1 - first : first));
  }
}

template <int CS_ITEMS>
__global__ void __launch_bounds__(CS_THREADS, CS_ITEMS == 4 ? 2 : (CS_ITEMS == 2 ? 4 : 6)) cse_rounds_kernel(CseArgs a) {
  constexpr int CS_TILE = CS_THREADS * CS_ITEMS;
  __shared__ uint64_t s_scan[CS_THREADS / 32];
  __shared__ uint32_t s_prefix[3];
  __shared__ uint32_t s_cnt[8][2];
  __shared__ uint32_t s_tstart[8][2];       // first global tile of (level, half)
  __shared__ unsigned long long s_emitted[8];
  __shared__ uint32_t s_flags[2];

  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  CseDeviceState* S = a.st;
  uint32_t round = vol_load(&S->round);
  uint32_t rounds_done = 0;
  // the previous launch (either kernel) left its exit status behind; every CTA looks at the
  // status only after the first grid barrier, so CTA 0 may clear it here
  if (blockIdx.x == 0 && tid == 0) S->status = kCseRunning;
  const unsigned long long barrier0 = vol_load64(&S->barriers);   // barriers passed by earlier launches

  for (;;) {
    const int cur = round & 1, nxt = cur ^ 1;
    if (tid < 16) s_cnt[tid >> 1][tid & 1] = vol_load(&S->cnt[cur][tid >> 1][tid & 1]);
    if (tid >= 32 && tid < 40) s_emitted[tid - 32] = vol_load64(&S->emitted[cur][tid - 32]);
    __syncthreads();
    if (tid == 0) {
      uint32_t t = 0;
      unsigned long long nodes = 0;
      uint32_t drain = 0, widest = 0;
      for (int l = 0; l < 8; ++l) {
        unsigned long long lvl = 0;
        for (int h = 0; h < 2; ++h) {
          s_tstart[l][h] = t;
          t += (s_cnt[l][h] + CS_TILE - 1) / CS_TILE;
          lvl += s_cnt[l][h];
        }
        nodes += lvl;
        widest = max(widest, uint32_t(min(lvl, 0xFFFFFFFFull)));
        if (s_emitted[l] + lvl > a.ecap[l]) drain = 1;   // a round emits at most one count per node
      }
      s_flags[0] = t;
      s_flags[1] = nodes == 0 ? kCseDone
                 : round >= a.round_limit ? kCseRunaway
                 : (a.use_narrow && widest <= kNarrowEnter) ? kCseGoNarrow
                 : drain ? kCseDrain : kCseRunning;
      if (blockIdx.x == 0 && s_flags[1] == kCseRunning && rounds_done < a.max_rounds) {
        S->visits += nodes;
        if (nodes > S->peak_frontier) S->peak_frontier = nodes;
      }
    }
    __syncthreads();
    const uint32_t total_tiles = s_flags[0];
    const uint32_t decision = s_flags[1];
    if (decision != kCseRunning || rounds_done >= a.max_rounds) {
      if (blockIdx.x == 0 && tid == 0) { S->status = decision; S->round = round; S->barriers = barrier0 + rounds_done; }
      break;
    }
    // levels without nodes hand an empty frontier (and their emission cursor) to the next round
    if (blockIdx.x == 0 && tid < 8) {
      const int l = tid;
      if (s_cnt[l][0] + s_cnt[l][1] == 0) {
        S->cnt[nxt][(l + 1) & 7][0] = 0;
        S->cnt[nxt][(l + 1) & 7][1] = 0;
        S->emitted[nxt][l] = s_emitted[l];
      }
    }
    const uint32_t tag = round + 1;

    for (uint32_t tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      // which (level, half) does this tile belong to?
      int l = 0, hh = 0;
#pragma unroll
      for (int k = 1; k < 16; ++k)
        if (tile >= s_tstart[k >> 1][k & 1]) { l = k >> 1; hh = k & 1; }
      // (s_tstart is non-decreasing; an empty (level, half) shares its start with the next one,
      //  the loop keeps the last match, which is the non-empty owner)
      const uint32_t count = s_cnt[l][hh];
      const uint32_t o0 = (tile - s_tstart[l][hh]) * CS_TILE + tid * CS_ITEMS;
      const int nv = o0 >= count ? 0 : int(min(uint32_t(CS_ITEMS), count - o0));
      const int ln = (l + 1) & 7;

      uint32_t ns[CS_ITEMS], na[CS_ITEMS], nb[CS_ITEMS];
#pragma unroll
      for (int j = 0; j < CS_ITEMS; ++j) ns[j] = na[j] = nb[j] = 0;
      if (nv) {                        // one-half is stored back to front: ordinal o at cap-1-o
        load_items<CS_ITEMS>(a.fs[cur][l], o0, hh != 0, a.cap, ns);
        load_items<CS_ITEMS>(a.fa[cur][l], o0, hh != 0, a.cap, na);
        load_items<CS_ITEMS>(a.fb[cur][l], o0, hh != 0, a.cap, nb);
      }
      // three rank words per node, all independent (bce.cpp:1265, 1271, 1301)
      const uint64_t* __restrict__ R = a.ranks[l];
      uint64_t wa[CS_ITEMS], wb[CS_ITEMS], wc[CS_ITEMS];
#pragma unroll
      for (int j = 0; j < CS_ITEMS; ++j) {
        if (j < nv) {
          wa[j] = __ldg(R + (ns[j] >> 5));
          wb[j] = __ldg(R + ((ns[j] + na[j] + nb[j]) >> 5));
          wc[j] = __ldg(R + ((ns[j] + na[j]) >> 5));
        } else { wa[j] = wb[j] = wc[j] = 0; }
      }
      uint32_t s0v[CS_ITEMS], s1v[CS_ITEMS], n1x[CS_ITEMS], n0x0[CS_ITEMS];
      uint32_t fz = 0, fo = 0, fe = 0;           // bit j: node j has zero-child / one-child / emits
#pragma unroll
      for (int j = 0; j < CS_ITEMS; ++j) {
        if (j < nv) {
          const uint32_t s = ns[j], x0 = na[j], x1 = nb[j], x = x0 + x1;
          const uint32_t s1 = rank1_word(wa[j], s);
          const uint32_t c1 = rank1_word(wb[j], s + x) - s1;         // _1x
          const uint32_t s0 = s - s1;
          const uint32_t z0 = (s + x0 - rank1_word(wc[j], s + x0)) - s0;   // _0x0 by rank (:1301)
          s0v[j] = s0; s1v[j] = s1; n1x[j] = c1; n0x0[j] = z0;
          if (c1 == 0) fz |= 1u << j;                                 // :1274
          else if (c1 == x) fo |= 1u << j;                            // :1282
          else {
            const uint32_t c0 = x - c1;
            const uint32_t lo = x0 > c1 ? x0 - c1 : 0u;               // :1290-1294
            const uint32_t hi = x0 - (c1 > x1 ? c1 - x1 : 0u);
            if (hi != lo) fe |= 1u << j;                              // :1299
            const uint32_t z1 = c0 - z0;                              // _0x1
            const uint32_t o1 = x1 - z1;                              // _1x1
            const uint32_t o0c = c1 - o1;                             // _1x0
            if (z0 && z1) fz |= 1u << j;                              // :1338
            if (o0c && o1) fo |= 1u << j;                             // :1345
          }
        }
      }
      // ordered positions by prefix sums: (zero-children | one-children | counts) in 3 x 21 bits
      const uint64_t mine = uint64_t(__popc(fz)) | (uint64_t(__popc(fo)) << 21) | (uint64_t(__popc(fe)) << 42);
      uint64_t tile_tot;
      const uint64_t excl = block_exclusive_scan<uint64_t, CS_THREADS>(mine, s_scan, tile_tot);
      const uint32_t first = s_tstart[l][0];
      if (warp < 3) {
        const uint32_t agg = uint32_t(tile_tot >> (21 * warp)) & 0x1FFFFFu;
        const uint32_t pre = lookback_warp(a.desc + size_t(warp) * a.desc_tiles, tile, first, tag, agg, &S->err);
        if (lane == 0) s_prefix[warp] = pre;
      }
      __syncthreads();
      uint32_t pz = s_prefix[0] + (uint32_t(excl) & 0x1FFFFFu);
      uint32_t po = s_prefix[1] + (uint32_t(excl >> 21) & 0x1FFFFFu);
      unsigned long long pe = s_emitted[l] + s_prefix[2] + (uint32_t(excl >> 42) & 0x1FFFFFu);

      uint32_t* __restrict__ zs = a.fs[nxt][ln];
      uint32_t* __restrict__ za = a.fa[nxt][ln];
      uint32_t* __restrict__ zb = a.fb[nxt][ln];
      bce_tuple* __restrict__ em = a.emit[l];
      const uint32_t one_base = a.C[ln];
#pragma unroll
      for (int j = 0; j < CS_ITEMS; ++j) {
        if (j < nv) {
          const uint32_t x0 = na[j], x1 = nb[j], x = x0 + x1, c1 = n1x[j];
          uint32_t za0, za1, oa0, oa1;                                 // child payloads
          if (c1 == 0) { za0 = x0; za1 = x1; oa0 = oa1 = 0; }
          else if (c1 == x) { oa0 = x0; oa1 = x1; za0 = za1 = 0; }
          else {
            const uint32_t c0 = x - c1;
            za0 = n0x0[j]; za1 = c0 - za0;
            oa1 = x1 - za1; oa0 = c1 - oa1;
            if (fe >> j & 1u) {
              const uint32_t lo = x0 > c1 ? x0 - c1 : 0u;
              const uint32_t hi = x0 - (c1 > x1 ? c1 - x1 : 0u);
              if (pe < a.ecap[l]) {
                bce_tuple t;
                t.sym = za0 - lo; t.k = hi - lo + 1; t.c1 = c0; t.c2 = x1; t.cs = x;   // :1302
                em[pe] = t;
              }
              ++pe;
            }
          }
          if (fz >> j & 1u) {
            if (pz < a.cap) { zs[pz] = s0v[j]; za[pz] = za0; zb[pz] = za1; }
            ++pz;
          }
          if (fo >> j & 1u) {
            if (po < a.cap) {
              const uint32_t at = a.cap - 1 - po;
              zs[at] = one_base + s1v[j]; za[at] = oa0; zb[at] = oa1;
            }
            ++po;
          }
        }
      }
      // the last tile of a level knows the level's totals: publish the next frontier sizes
      const uint32_t level_tiles = (s_cnt[l][0] + CS_TILE - 1) / CS_TILE + (s_cnt[l][1] + CS_TILE - 1) / CS_TILE;
      if (tile == first + level_tiles - 1 && tid == 0) {
        const uint32_t tz = s_prefix[0] + (uint32_t(tile_tot) & 0x1FFFFFu);
        const uint32_t to = s_prefix[1] + (uint32_t(tile_tot >> 21) & 0x1FFFFFu);
        const uint32_t te = s_prefix[2] + (uint32_t(tile_tot >> 42) & 0x1FFFFFu);
        S->cnt[nxt][ln][0] = tz;
        S->cnt[nxt][ln][1] = to;
        S->emitted[nxt][l] = s_emitted[l] + te;
        if (uint64_t(tz) + to > a.cap) atomicExch(&S->status, uint32_t(kCseOverflow));
      }
      __syncthreads();        // s_prefix / s_scan are reused by the next tile
    }

    grid_barrier(S, barrier0 + rounds_done, round);
    ++round;
    ++rounds_done;
    if (vol_load(&S->status) != kCseRunning || vol_load(&S->err) != 0) {
      if (blockIdx.x == 0 && tid == 0) { S->round = round; S->barriers = barrier0 + rounds_done; }
      break;
    }
  }
}

}  // namespace bce

#include "cse_wide.cuh"   // cse_wide_kernel<ITEMS>: the software-pipelined wide kernel (default)

namespace bce {

// ---------------------------------------------------------------------------------
// narrow mode: one cluster, one CTA per level
// ---------------------------------------------------------------------------------
struct NarrowShared {
  uint32_t s[2][NR_CAP], a[2][NR_CAP], b[2][NR_CAP];   // this level's frontier, double buffered
  uint32_t cz[2], co[2];                                // zero-/one-half sizes of buffer p
  uint32_t all_cnt[2][8];                               // every level's frontier size (all-gathered)
  unsigned long long all_emitted[2][8];                 // every level's emission cursor (all-gathered)
  uint64_t scan[NR_THREADS / 32];
  uint32_t decision;
};

__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(NR_THREADS, 1) cse_narrow_kernel(CseArgs a) {
  cg::cluster_group cluster = cg::this_cluster();
  __shared__ NarrowShared sh;
  const unsigned tid = threadIdx.x;
  const int l = int(cluster.block_rank());          // level handled by this CTA
  const int ln = (l + 1) & 7;
  CseDeviceState* S = a.st;
  const uint64_t* __restrict__ R = a.ranks[l];

  uint32_t round = S->round;
  const int gpar = round & 1;
  // frontier of this level from the wide layout (zero-half ascending, one-half from the back)
  {
    const uint32_t cz = S->cnt[gpar][l][0], co = S->cnt[gpar][l][1];
    if (tid == 0) { sh.cz[0] = cz; sh.co[0] = co; }
    for (uint32_t t = tid; t < cz + co && t < NR_CAP; t += NR_THREADS) {
      const uint32_t idx = t < cz ? t : a.cap - 1 - (t - cz);
      sh.s[0][t] = a.fs[gpar][l][idx];
      sh.a[0][t] = a.fa[gpar][l][idx];
      sh.b[0][t] = a.fb[gpar][l][idx];
    }
    if (tid < 8) {
      sh.all_cnt[0][tid] = S->cnt[gpar][tid][0] + S->cnt[gpar][tid][1];
      sh.all_emitted[0][tid] = S->emitted[gpar][tid];
    }
  }
  unsigned long long cursor = S->emitted[gpar][l];
  unsigned long long visits = 0, peak = 0;
  int p = 0;
  cluster.sync();                                    // nobody writes into a CTA that is still loading

  uint32_t status;
  for (;;) {
    if (tid == 0) {
      uint32_t total = 0, widest = 0, drain = 0;
      for (int k = 0; k < 8; ++k) {
        const uint32_t cnt = sh.all_cnt[p][k];
        total += cnt;
        widest = max(widest, cnt);
        if (sh.all_emitted[p][k] + cnt > a.ecap[k]) drain = 1;
      }
      sh.decision = total == 0 ? kCseDone
                  : round >= a.round_limit ? kCseRunaway
                  : widest > kNarrowLeave ? kCseGoWide
                  : drain ? kCseDrain : kCseRunning;
      if (sh.decision == kCseRunning) { visits += total; peak = max(peak, (unsigned long long)total); }
    }
    __syncthreads();
    status = sh.decision;
    if (status != kCseRunning) break;

    const uint32_t cz = sh.cz[p], co = sh.co[p];
    const bool live = tid < cz + co;
    uint32_t s = 0, x0 = 0, x1 = 0;
    if (live) { s = sh.s[p][tid]; x0 = sh.a[p][tid]; x1 = sh.b[p][tid]; }
    uint32_t fz = 0, fo = 0, fe = 0, s0 = 0, s1 = 0, c1 = 0, z0 = 0;
    const uint32_t x = x0 + x1;
    if (live) {
      const uint64_t wa = __ldg(R + (s >> 5));
      const uint64_t wb = __ldg(R + ((s + x) >> 5));
      const uint64_t wc = __ldg(R + ((s + x0) >> 5));
      s1 = rank1_word(wa, s);
      c1 = rank1_word(wb, s + x) - s1;
      s0 = s - s1;
      z0 = (s + x0 - rank1_word(wc, s + x0)) - s0;
      if (c1 == 0) fz = 1;
      else if (c1 == x) fo = 1;
      else {
        const uint32_t c0 = x - c1;
        const uint32_t lo = x0 > c1 ? x0 - c1 : 0u;
        const uint32_t hi = x0 - (c1 > x1 ? c1 - x1 : 0u);
        fe = hi != lo;
        const uint32_t z1 = c0 - z0, o1 = x1 - z1, o0c = c1 - o1;
        fz = z0 && z1;
        fo = o0c && o1;
      }
    }
    const uint64_t mine = uint64_t(fz) | (uint64_t(fo) << 21) | (uint64_t(fe) << 42);
    uint64_t tot;
    const uint64_t excl = block_exclusive_scan<uint64_t, NR_THREADS>(mine, sh.scan, tot);
    const uint32_t tz = uint32_t(tot) & 0x1FFFFFu, to = uint32_t(tot >> 21) & 0x1FFFFFu, te = uint32_t(tot >> 42) & 0x1FFFFFu;
    const int q = p ^ 1;
    // the next level's CTA receives its frontier directly in its shared memory
    uint32_t* rs = cluster.map_shared_rank(&sh.s[q][0], ln);
    uint32_t* ra = cluster.map_shared_rank(&sh.a[q][0], ln);
    uint32_t* rb = cluster.map_shared_rank(&sh.b[q][0], ln);
    if (live) {
      uint32_t za0, za1, oa0, oa1;
      if (c1 == 0) { za0 = x0; za1 = x1; oa0 = oa1 = 0; }
      else if (c1 == x) { oa0 = x0; oa1 = x1; za0 = za1 = 0; }
      else {
        const uint32_t c0 = x - c1;
        za0 = z0; za1 = c0 - z0;
        oa1 = x1 - za1; oa0 = c1 - oa1;
        if (fe) {
          const uint32_t lo = x0 > c1 ? x0 - c1 : 0u;
          const uint32_t hi = x0 - (c1 > x1 ? c1 - x1 : 0u);
          const unsigned long long pe = cursor + (uint32_t(excl >> 42) & 0x1FFFFFu);
          if (pe < a.ecap[l]) {
            bce_tuple t;
            t.sym = za0 - lo; t.k = hi - lo + 1; t.c1 = c0; t.c2 = x1; t.cs = x;      // bce.cpp:1302
            a.emit[l][pe] = t;
          }
        }
      }
      if (fz) { const uint32_t at = uint32_t(excl) & 0x1FFFFFu; rs[at] = s0; ra[at] = za0; rb[at] = za1; }
      if (fo) { const uint32_t at = tz + (uint32_t(excl >> 21) & 0x1FFFFFu); rs[at] = a.C[ln] + s1; ra[at] = oa0; rb[at] = oa1; }
    }
    cursor += te;
    if (tid == 0) {
      *cluster.map_shared_rank(&sh.cz[q], ln) = tz;
      *cluster.map_shared_rank(&sh.co[q], ln) = to;
    }
    if (tid < 8) {            // all-gather: level ln's next size and this level's cursor, to every CTA
      cluster.map_shared_rank(&sh.all_cnt[q][0], tid)[ln] = tz + to;
      cluster.map_shared_rank(&sh.all_emitted[q][0], tid)[l] = cursor;
    }
    cluster.sync();
    p = q;
    ++round;
  }

  // hand the state back in the wide layout
  {
    const int opar = round & 1;
    const uint32_t cz = sh.cz[p], co = sh.co[p];
    for (uint32_t t = tid; t < cz + co; t += NR_THREADS) {
      const uint32_t idx = t < cz ? t : a.cap - 1 - (t - cz);
      a.fs[opar][l][idx] = sh.s[p][t];
      a.fa[opar][l][idx] = sh.a[p][t];
      a.fb[opar][l][idx] = sh.b[p][t];
    }
    if (tid == 0) {
      S->cnt[opar][l][0] = cz;
      S->cnt[opar][l][1] = co;
      S->emitted[opar][l] = cursor;
      if (l == 0) {
        S->round = round;
        S->status = status;
        S->visits += visits;
        if (peak > S->peak_frontier) S->peak_frontier = peak;
      }
    }
  }
  cluster.sync();     // no CTA may leave while its shared memory can still be a target
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
static size_t env_size(const char* name, size_t dflt) {
  const char* v = getenv(name);
  if (!v || !*v) return dflt;
  return size_t(strtoull(v, nullptr, 10));
}

void cse_destroy(Ctx* c) {
  delete c->cse;
  c->cse = nullptr;
  c->cse_active = false;
}

int cse_begin(Ctx* c, uint32_t n) {
  if (!c->ranks_resident) { set_error(c, "cse_begin: wavelet matrix not built"); return BCE_GPU_E_STATE; }
  if (!c->cse) c->cse = new CseHost();
  CseHost* H = c->cse;
  H->n = n;
  cudaStream_t st = c->stream;

  // ---- sizing: frontier first (correctness), emission with what is left ------------
  const size_t budget = scratch_budget(c);
  const size_t cap_full = ((size_t(n) / 2 + 4) + 3) & ~size_t(3);
  size_t cap = cap_full;
  auto frontier_bytes = [](size_t cp) { return 48 * Carver::need(cp, 4); };
  const int items = int(env_size("BCE_GPU_CSE_ITEMS", 0));
  H->auto_items = items == 0 && env_size("BCE_GPU_CSE_DIRECT", 0) == 0;
  H->items = (items == 1 || items == 4) ? items : 2;
  H->known_nodes = 0;
  const size_t tile = size_t(CS_THREADS);          // descriptors sized for the smallest tile of any variant
  auto desc_tiles_for = [tile](size_t cp) { return 8 * (cp / tile + 2); };
  while (cap > 4096 && frontier_bytes(cap) > budget / 2) cap = (cap / 2 + 3) & ~size_t(3);
  const size_t desc_tiles = desc_tiles_for(cap);
  const size_t desc_bytes = Carver::need(3 * desc_tiles, 8);
  const size_t pinned_limit = env_size("BCE_GPU_PINNED_LIMIT", size_t(8) << 30);
  size_t left = budget > frontier_bytes(cap) + desc_bytes ? budget - frontier_bytes(cap) - desc_bytes : 0;
  size_t ecap = size_t(n);                                  // a level emits at most n-1 counts in total
  const size_t per_level_min = cap + CS_MAX_TILE;           // one round must always fit
  if (ecap * 8 * sizeof(bce_tuple) > left) ecap = left / (8 * sizeof(bce_tuple));
  if (ecap * 8 * sizeof(bce_tuple) > pinned_limit) ecap = pinned_limit / (8 * sizeof(bce_tuple));
  if (ecap < per_level_min) ecap = per_level_min;
  const size_t need = frontier_bytes(cap) + desc_bytes + 8 * Carver::need(ecap, sizeof(bce_tuple)) + 4096;
  BCE_TRY(c->scratch.ensure(c, need));
  Carver cv(c->scratch.p, c->scratch.cap);

  CseArgs& a = H->args;
  for (int p = 0; p < 2; ++p)
    for (int l = 0; l < 8; ++l) {
      a.fs[p][l] = cv.take<uint32_t>(cap);
      a.fa[p][l] = cv.take<uint32_t>(cap);
      a.fb[p][l] = cv.take<uint32_t>(cap);
    }
  a.desc = cv.take<uint64_t>(3 * desc_tiles);
  for (int l = 0; l < 8; ++l) { a.emit[l] = cv.take<bce_tuple>(ecap); a.ecap[l] = ecap; }
  if (!cv.ok()) { set_error(c, "cse_begin: scratch carve failed (need %zu)", need); return BCE_GPU_E_NOMEM; }
  const size_t words = size_t(n) / 32 + 1;
  for (int l = 0; l < 8; ++l) { a.ranks[l] = c->ranks.as<uint64_t>() + size_t(l) * words; a.C[l] = c->C[l]; }
  a.cap = uint32_t(cap);
  a.desc_tiles = uint32_t(desc_tiles);
  a.max_rounds = 0x7FFFFFFFu;
  a.round_limit = uint32_t(std::min<uint64_t>(uint64_t(n) * 8 + 64, 0xFFFFFFF0ull));
  a.use_narrow = env_size("BCE_GPU_NO_NARROW", 0) ? 0u : 1u;
  a.dbg = 0;
  a.min_nodes = 0;
  a.max_nodes = ~0ull;
  H->narrow = a.use_narrow != 0;
  H->last_round = 0;
  a.st = reinterpret_cast<CseDeviceState*>(c->small.as<char>() + kSmallCse);
  static_assert(sizeof(CseDeviceState) <= 1024, "state must fit its slot in Ctx::small");

  BCE_CUDA(c, cudaMemsetAsync(a.desc, 0, 3 * desc_tiles * sizeof(uint64_t), st));
  cse_init_kernel<<<1, 32, 0, st>>>(a, n);
  c->stats.gpu_launches++;
  BCE_CUDA(c, cudaGetLastError());

  {
    int per_sm = 0;
    const bool direct = env_size("BCE_GPU_CSE_DIRECT", 0) != 0 || H->items == 1;   // unpipelined variant
    H->wide_smem = 0;
    if (!direct && H->items == 4) {
      H->wide_fn = (const void*)cse_wide_kernel<4>;
      H->wide_smem = 2 * sizeof(WideStage<4>);
      BCE_CUDA(c, cudaFuncSetAttribute(cse_wide_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(H->wide_smem)));
      BCE_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cse_wide_kernel<4>, CS_THREADS, H->wide_smem));
    } else if (!direct) {
      H->wide_fn = (const void*)cse_wide_kernel<2>;
      H->wide_smem = 2 * sizeof(WideStage<2>);
      BCE_CUDA(c, cudaFuncSetAttribute(cse_wide_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(H->wide_smem)));
      BCE_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cse_wide_kernel<2>, CS_THREADS, H->wide_smem));
    } else if (H->items == 4) {
      H->wide_fn = (const void*)cse_rounds_kernel<4>;
      BCE_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cse_rounds_kernel<4>, CS_THREADS, 0));
    } else if (H->items == 1) {
      H->wide_fn = (const void*)cse_rounds_kernel<1>;
      BCE_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cse_rounds_kernel<1>, CS_THREADS, 0));
    } else {
      H->wide_fn = (const void*)cse_rounds_kernel<2>;
      BCE_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, cse_rounds_kernel<2>, CS_THREADS, 0));
    }
    if (per_sm < 1) { set_error(c, "cse kernel does not fit on an SM"); return BCE_GPU_E_CUDA; }
    H->grid = per_sm * c->sm_count;
    if (H->auto_items) {
      int p2 = 0, p4 = 0;
      H->var_fn[0] = (const void*)cse_wide_kernel<2>; H->var_smem[0] = 2 * sizeof(WideStage<2>);
      H->var_fn[1] = (const void*)cse_wide_kernel<4>; H->var_smem[1] = 2 * sizeof(WideStage<4>);
      BCE_CUDA(c, cudaFuncSetAttribute(cse_wide_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(H->var_smem[0])));
      BCE_CUDA(c, cudaFuncSetAttribute(cse_wide_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(H->var_smem[1])));
      BCE_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p2, cse_wide_kernel<2>, CS_THREADS, H->var_smem[0]));
      BCE_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&p4, cse_wide_kernel<4>, CS_THREADS, H->var_smem[1]));
      if (p2 < 1 || p4 < 1) { set_error(c, "cse kernel does not fit on an SM"); return BCE_GPU_E_CUDA; }
      H->var_grid[0] = p2 * c->sm_count;
      H->var_grid[1] = p4 * c->sm_count;
    }
  }
  c->cse_active = true;
  c->cse_done = false;
  BCE_TRACE("cse_begin n=%u cap=%zu ecap=%zu grid=%d", n, cap, ecap, H->grid);
  return BCE_GPU_OK;
}

// Runs rounds until the loop ends or an emission buffer may overflow; resident = leave the
// counts in device memory (measurement), otherwise copy them to pinned host memory.
int cse_advance(Ctx* c, bool resident, bce_cse_batch* out) {
  if (!c->cse_active || !c->cse) { set_error(c, "cse_next without cse_begin"); return BCE_GPU_E_STATE; }
  CseHost* H = c->cse;
  cudaStream_t st = c->stream;
  if (out) memset(out, 0, sizeof *out);
  if (c->cse_done) { if (out) out->done = 1; return BCE_GPU_OK; }

  CseDeviceState* h_state = reinterpret_cast<CseDeviceState*>(c->pinned_small.as<char>() + 40 * 1024);
  BCE_CUDA(c, cudaEventRecord(c->ev[2], st));
  float ms = 0;
  for (int hops = 0;; ++hops) {
    if (hops > 100000) { set_error(c, "cse: wide/narrow ping-pong"); return BCE_GPU_E_INTERNAL; }
    const bool was_narrow = H->narrow;
    BCE_CUDA(c, cudaEventRecord(c->ev[0], st));
    {
      const uint32_t dbg_round = uint32_t(env_size("BCE_GPU_CSE_DBG_ROUND", 0));
      H->args.max_rounds = 0x7FFFFFFFu;
      H->args.dbg = 0;
      if (dbg_round && !H->narrow) {
        if (H->last_round < dbg_round) H->args.max_rounds = dbg_round - H->last_round;   // stop right before it
        else if (H->last_round == dbg_round) { H->args.max_rounds = 1; H->args.dbg = uint32_t(env_size("BCE_GPU_CSE_DBG_FLAGS", 0)); }
      }
    }
    if (H->narrow) {
      cse_narrow_kernel<<<8, NR_THREADS, 0, st>>>(H->args);
      BCE_CUDA(c, cudaGetLastError());
    } else {
      const void* fn = H->wide_fn;
      size_t smem = H->wide_smem;
      int grid = H->grid;
      H->args.min_nodes = 0;
      H->args.max_nodes = ~0ull;
      if (H->auto_items) {
        constexpr unsigned long long kBig = 2000000, kLeaveBig = 1000000, kLeaveSmall = 3000000;
        const int v = H->known_nodes >= kBig ? 1 : 0;
        fn = H->var_fn[v]; smem = H->var_smem[v]; grid = H->var_grid[v];
        if (v) H->args.min_nodes = kLeaveBig; else H->args.max_nodes = kLeaveSmall;
      }
      void* kargs[] = {&H->args};
      BCE_CUDA(c, cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(CS_THREADS), kargs, smem, st));
    }
    c->stats.gpu_launches++;
    c->stats.cse_launches++;
    BCE_CUDA(c, cudaMemcpyAsync(h_state, H->args.st, sizeof(CseDeviceState), cudaMemcpyDeviceToHost, st));
    BCE_CUDA(c, cudaEventRecord(c->ev[1], st));
    BCE_CUDA(c, cudaStreamSynchronize(st));
    {
      float lms = 0;
      BCE_CUDA(c, cudaEventElapsedTime(&lms, c->ev[0], c->ev[1]));
      if (was_narrow) { c->stats.ms_cse_narrow += lms; c->stats.cse_rounds_narrow += h_state->round - H->last_round; }
      BCE_TRACE("cse %s kernel: rounds %u..%u status=%u err=%u visits=%llu %.3f ms dbg=%u", was_narrow ? "narrow" : "wide",
                H->last_round, h_state->round, h_state->status, h_state->err, h_state->visits, lms, H->args.dbg);
      if (H->args.dbg) { set_error(c, "cse: timing experiment round done (%.3f ms)", lms); return BCE_GPU_E_INTERNAL; }
      H->last_round = h_state->round;
    }
    {
      const int par = h_state->round & 1;
      H->known_nodes = 0;
      for (int l = 0; l < 8; ++l) H->known_nodes += h_state->cnt[par][l][0] + h_state->cnt[par][l][1];
    }
    if (h_state->err) break;
    if (h_state->status == kCseRunning && H->args.max_rounds != 0x7FFFFFFFu) continue;   // stopped on request
    if (h_state->status == kCseGoWide) { H->narrow = false; continue; }
    if (h_state->status == kCseGoNarrow) { H->narrow = true; continue; }
    break;
  }
  BCE_CUDA(c, cudaEventRecord(c->ev[3], st));
  BCE_CUDA(c, cudaEventSynchronize(c->ev[3]));
  BCE_CUDA(c, cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]));
  c->stats.ms_cse += ms;

  if (h_state->err) {
    set_error(c, "cse: %s watchdog fired (round %u, barrier_fail %u, arrivals %llu, barriers %llu, grid %d)",
              h_state->err == 2 ? "grid-barrier" : "chained-scan", h_state->round, h_state->barrier_fail,
              h_state->arrivals, h_state->barriers, H->grid);
    return BCE_GPU_E_INTERNAL;
  }
  if (h_state->status == kCseRunaway) {
    set_error(c, "cse: level loop still running after %u rounds (n = %u)", h_state->round, H->n);
    return BCE_GPU_E_INTERNAL;
  }
  if (h_state->status == kCseOverflow) {
    set_error(c, "cse: node frontier exceeded %u nodes per level at round %u", H->args.cap, h_state->round);
    return BCE_GPU_E_FRONTIER;
  }
  if (h_state->status != kCseDone && h_state->status != kCseDrain) {
    set_error(c, "cse: unexpected kernel status %u", h_state->status);
    return BCE_GPU_E_INTERNAL;
  }
  const int par = h_state->round & 1;
  size_t total = 0, cnt[8];
  for (int l = 0; l < 8; ++l) { cnt[l] = size_t(h_state->emitted[par][l]); total += cnt[l]; }
  if (h_state->status == kCseDrain && total == 0) {
    set_error(c, "cse: kernel asked to drain empty emission buffers at round %u", h_state->round);
    return BCE_GPU_E_INTERNAL;
  }
  c->stats.cse_tuples += total;
  c->stats.cse_visits = h_state->visits;
  c->stats.cse_rounds = h_state->round ? h_state->round : 1;   // the reference's do..while runs at least once
  c->stats.cse_peak_frontier = h_state->peak_frontier;

  if (!resident && out) {
    BCE_TRY(c->pinned_emit.ensure(c, (total + 8) * sizeof(bce_tuple)));
    bce_tuple* hp = c->pinned_emit.as<bce_tuple>();
    BCE_CUDA(c, cudaEventRecord(c->ev[2], st));
    size_t at = 0;
    for (int l = 0; l < 8; ++l) {
      out->tuples[l] = hp + at;
      out->count[l] = cnt[l];
      if (cnt[l])
        BCE_CUDA(c, cudaMemcpyAsync(hp + at, H->args.emit[l], cnt[l] * sizeof(bce_tuple), cudaMemcpyDeviceToHost, st));
      at += cnt[l];
    }
    BCE_CUDA(c, cudaEventRecord(c->ev[3], st));
    BCE_CUDA(c, cudaEventSynchronize(c->ev[3]));
    BCE_CUDA(c, cudaEventElapsedTime(&ms, c->ev[2], c->ev[3]));
    c->stats.ms_d2h += ms;
  }
  if (h_state->status == kCseDone) {
    c->cse_done = true;
    if (out) out->done = 1;
  } else {
    cse_reset_emitted_kernel<<<1, 32, 0, st>>>(H->args.st);
    c->stats.gpu_launches++;
    BCE_CUDA(c, cudaGetLastError());
  }
  return BCE_GPU_OK;
}

}  // namespace bce
