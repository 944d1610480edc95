// radix_chunked.cuh -- one LSD radix pass as reduce -> scan -> scatter over CTA-owned chunks.
// Included by radix_sort.cu.
//
// Every CTA owns one contiguous chunk of the input (a multiple of the tile size).
//   radix_upsweep_kernel    per-chunk digit histogram                         (reads 8 B/elem)
//   radix_chunk_scan_kernel exclusive scan over (digit, chunk): where every chunk's run of
//                           every digit starts in the output                  (tiny)
//   radix_downsweep_kernel  the CTA walks its chunk tile by tile with running digit offsets in
//                           shared memory: rank in shared memory, stage, write per-digit bursts.
//                           No CTA ever waits for another one (no chained scan), and the next
//                           tile's keys/values are prefetched into registers while the current
//                           tile is written out.
// Against the single-pass onesweep kernel this reads the keys once more (32 instead of 24 bytes
// per element and pass) but has no look-back stalls.  Measured on B200 it is nevertheless the
// SLOWER of the two (profiles/r1_radix_experiments.md); kept selectable (BCE_GPU_RADIX=chunked)
// as the baseline for the next attempt at the sort core.
#pragma once

namespace bce {

constexpr int RD_THREADS = 256;
constexpr int RD_ITEMS = 12;
constexpr int RD_TILE = RD_THREADS * RD_ITEMS;
constexpr int RD_WARPS = RD_THREADS / 32;

__global__ void __launch_bounds__(RD_THREADS) radix_upsweep_kernel(const uint64_t* __restrict__ keys, uint32_t m,
                                                                   uint32_t chunk, int shift,
                                                                   uint32_t* __restrict__ hist) {
  __shared__ uint32_t whist[RD_WARPS][256];
  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < RD_WARPS * 256; i += RD_THREADS) (&whist[0][0])[i] = 0;
  __syncthreads();
  const uint64_t start = uint64_t(blockIdx.x) * chunk;
  const uint64_t end = min(uint64_t(m), start + chunk);
  // a warp reads 32 consecutive keys at a time, 4 such rows in flight
  for (uint64_t row = start + uint64_t(warp) * 32; row < end; row += uint64_t(RD_WARPS) * 32 * 4) {
    uint64_t k[4];
    bool in[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t i = row + uint64_t(u) * RD_WARPS * 32 + lane;
      in[u] = i < end;
      k[u] = in[u] ? keys[i] : 0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t d = in[u] ? uint32_t(k[u] >> shift) & 255u : 256u;      // 256 = not an element
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      if (in[u] && lane == unsigned(__ffs(peers) - 1)) whist[warp][d] += __popc(peers);
      __syncwarp();
    }
  }
  __syncthreads();
  uint32_t s = 0;
#pragma unroll
  for (int w = 0; w < RD_WARPS; ++w) s += whist[w][tid];
  hist[size_t(blockIdx.x) * 256 + tid] = s;
}

// offs[b][d] = (elements with a smaller digit) + (elements with digit d in chunks before b)
__global__ void __launch_bounds__(256) radix_chunk_scan_kernel(const uint32_t* __restrict__ hist, uint32_t chunks,
                                                               uint32_t* __restrict__ offs) {
  __shared__ uint32_t s_scan[8];
  const unsigned d = threadIdx.x;
  uint32_t total = 0;
  for (uint32_t b = 0; b < chunks; ++b) total += hist[size_t(b) * 256 + d];
  uint32_t all;
  uint32_t run = block_exclusive_scan<uint32_t, 256>(total, s_scan, all);
  for (uint32_t b = 0; b < chunks; ++b) {
    offs[size_t(b) * 256 + d] = run;
    run += hist[size_t(b) * 256 + d];
  }
}

struct RadixDown {
  const uint64_t* kin;
  const uint32_t* vin;      // may be NULL: keys only
  uint64_t* kout;
  uint32_t* vout;
  uint32_t m;
  uint32_t chunk;
  int shift;
  const uint32_t* offs;     // [chunks][256]
};

__global__ void __launch_bounds__(RD_THREADS, 3) radix_downsweep_kernel(RadixDown p) {
  __shared__ uint64_t s_keys[RD_TILE];
  __shared__ uint32_t s_vals[RD_TILE];
  __shared__ uint32_t s_whist[RD_WARPS][256];
  __shared__ uint32_t s_goff[256];        // where the next element of every digit goes (global)
  __shared__ uint32_t s_dstart[256];      // tile-local start of every digit
  __shared__ uint32_t s_cnt[256];
  __shared__ uint32_t s_scan[RD_WARPS];

  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint64_t start = uint64_t(blockIdx.x) * p.chunk;
  if (start >= p.m) return;
  const uint64_t end = min(uint64_t(p.m), start + p.chunk);
  const uint32_t tiles = uint32_t((end - start + RD_TILE - 1) / RD_TILE);
  const bool has_vals = p.vin != nullptr;
  s_goff[tid] = p.offs[size_t(blockIdx.x) * 256 + tid];

  uint64_t key[RD_ITEMS];
  uint32_t val[RD_ITEMS];
  uint32_t pos[RD_ITEMS];
  auto load_tile = [&](uint32_t t) {
    const uint64_t wbase = start + uint64_t(t) * RD_TILE + warp * (32 * RD_ITEMS);
#pragma unroll
    for (int k = 0; k < RD_ITEMS; ++k) {
      const uint64_t i = wbase + k * 32 + lane;
      const bool in = i < end;
      key[k] = in ? p.kin[i] : ~0ull;            // padding sorts to the very end of the tile
      val[k] = (in && has_vals) ? p.vin[i] : 0u;
    }
  };
  load_tile(0);

  for (uint32_t t = 0; t < tiles; ++t) {
    const uint32_t valid = uint32_t(min(uint64_t(RD_TILE), end - (start + uint64_t(t) * RD_TILE)));
    for (int i = tid; i < RD_WARPS * 256; i += RD_THREADS) (&s_whist[0][0])[i] = 0;
    __syncthreads();                         // also: previous tile's write-out and offset update are done
    // rank inside the warp's chunk of the tile, in index order (stable)
#pragma unroll
    for (int k = 0; k < RD_ITEMS; ++k) {
      const uint32_t d = uint32_t(key[k] >> p.shift) & 255u;
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      const unsigned leader = __ffs(peers) - 1;
      uint32_t before = 0;
      if (lane == leader) {
        before = s_whist[warp][d];
        s_whist[warp][d] = before + __popc(peers);
      }
      before = __shfl_sync(0xffffffffu, before, leader);
      pos[k] = before + __popc(peers & lanemask_lt());
      __syncwarp();
    }
    __syncthreads();
    {   // thread d owns digit d
      const uint32_t d = tid;
      uint32_t sum = 0;
#pragma unroll
      for (int w = 0; w < RD_WARPS; ++w) {
        const uint32_t c = s_whist[w][d];
        s_whist[w][d] = sum;
        sum += c;
      }
      uint32_t tot;
      const uint32_t dstart = block_exclusive_scan<uint32_t, RD_THREADS>(sum, s_scan, tot);
      s_dstart[d] = dstart;
      s_cnt[d] = (d == 255u) ? sum - (uint32_t(RD_TILE) - valid) : sum;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RD_ITEMS; ++k) {
      const uint32_t d = uint32_t(key[k] >> p.shift) & 255u;
      const uint32_t at = pos[k] + s_dstart[d] + s_whist[warp][d];
      s_keys[at] = key[k];
      s_vals[at] = val[k];
    }
    // the registers are free: fetch the next tile while this one is written out
    if (t + 1 < tiles) load_tile(t + 1);
    __syncthreads();
    for (uint32_t j = tid; j < valid; j += RD_THREADS) {
      const uint64_t k64 = s_keys[j];
      const uint32_t d = uint32_t(k64 >> p.shift) & 255u;
      const uint32_t g = s_goff[d] + (j - s_dstart[d]);
      p.kout[g] = k64;
      if (has_vals) p.vout[g] = s_vals[j];
    }
    __syncthreads();
    s_goff[tid] += s_cnt[tid];
  }
}

}  // namespace bce
