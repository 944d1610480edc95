// wavelet.cu -- stage B1: the 8-level LSB-first wavelet matrix with rank directory.
//
// Replaces the RankFile constructor body (bce.cpp:944-972) and Rank::build (:138-145).
// Level j holds bit j of the BWT bytes stably sorted by their low j bits, i.e. level j+1's
// byte order is level j's order stably partitioned by bit j (zeros first).  The reference
// reaches the same layout with 8n single-bit read-modify-writes at per-context cursors;
// here every level is one streaming pass:
//
//   read 16 bytes/thread -> bit j of each -> block scan of zero counts -> chained scan over
//   tiles (one carried value) -> (a) rank words [bits << 32 | ones before] straight from the
//   scan, (b) bytes staged in shared memory as [zeros | ones] and written as two bursts.
//
// HBM traffic per level: read n, write n, write n/4 (rank words)  ->  ~19 n for 8 levels
// (SURVEY.md 8d).  The zero totals Z_j needed up front come from one byte histogram.
#include "ctx.h"

namespace bce {

constexpr int WV_THREADS = 256;
constexpr int WV_BYTES = 64;                       // per thread: two whole rank words
constexpr int WV_TILE = WV_THREADS * WV_BYTES;     // 16384 positions

__global__ void __launch_bounds__(256) byte_hist_kernel(const uint8_t* __restrict__ L, uint32_t n,
                                                        uint32_t* __restrict__ hist) {
  __shared__ uint32_t h[8][256];                    // one copy per warp: fewer same-address clashes
  for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) (&h[0][0])[i] = 0;
  __syncthreads();
  const unsigned warp = threadIdx.x >> 5;
  const uint32_t words = n / 4;
  const uint32_t stride = gridDim.x * blockDim.x;
  const uint32_t* L32 = reinterpret_cast<const uint32_t*>(L);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += stride) {
    uint32_t w = L32[i];
    atomicAdd(&h[warp][w & 255u], 1u);
    atomicAdd(&h[warp][(w >> 8) & 255u], 1u);
    atomicAdd(&h[warp][(w >> 16) & 255u], 1u);
    atomicAdd(&h[warp][w >> 24], 1u);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3u)) atomicAdd(&h[0][L[words * 4 + threadIdx.x]], 1u);
  __syncthreads();
  for (int v = threadIdx.x; v < 256; v += blockDim.x) {
    uint32_t s = 0;
    for (int w = 0; w < 8; ++w) s += h[w][v];
    if (s) atomicAdd(&hist[v], s);
  }
}

void byte_hist_launch(Ctx* c, const uint8_t* L, uint32_t n, uint32_t* d_hist) {
  uint64_t hb64 = (uint64_t(n) / 4 + 255) / 256 + 1, hbmax = uint64_t(c->sm_count) * 8;
  int hb = int(hb64 < hbmax ? hb64 : hbmax);
  byte_hist_kernel<<<hb, 256, 0, c->stream>>>(L, n, d_hist);
}

// zeros[j] = number of bytes whose bit j is 0
__global__ void zeros_from_hist_kernel(const uint32_t* __restrict__ hist, uint32_t* __restrict__ zeros) {
  __shared__ uint32_t z[8];
  if (threadIdx.x < 8) z[threadIdx.x] = 0;
  __syncthreads();
  uint32_t v = threadIdx.x, c = hist[v];
  for (int j = 0; j < 8; ++j)
    if (!((v >> j) & 1u) && c) atomicAdd(&z[j], c);
  __syncthreads();
  if (threadIdx.x < 8) zeros[threadIdx.x] = z[threadIdx.x];
}

struct WaveletPass {
  const uint8_t* in;        // bytes in level-j order
  uint8_t* out;             // bytes in level-(j+1) order (NULL for the last level)
  uint64_t* rank;           // level j rank words
  uint32_t n;
  int bit;
  const uint32_t* zeros;    // [8] totals
  uint64_t* desc;
  uint32_t* ticket;
  uint32_t tag;
  uint32_t* err;
};

// Copies len bytes from shared memory (any offset) to global memory (any alignment): single bytes up
// to the first 16-byte boundary of the destination, then 16-byte stores assembled from aligned
// shared-memory words with funnel shifts, single bytes for the rest.
__device__ __forceinline__ void copy_run(uint8_t* __restrict__ dst, const uint8_t* s_src, uint32_t len, unsigned tid) {
  const uint32_t mis = uint32_t(reinterpret_cast<uintptr_t>(dst)) & 15u;
  const uint32_t head = min(len, (16u - mis) & 15u);
  if (tid < head) dst[tid] = s_src[tid];
  const uint32_t chunks = (len - head) / 16;
  const uint8_t* sb = s_src + head;
  uint4* d16 = reinterpret_cast<uint4*>(dst + head);
  const uint32_t sh = (uint32_t(reinterpret_cast<uintptr_t>(sb)) & 3u) * 8;
  const uint32_t* sw = reinterpret_cast<const uint32_t*>(sb - (sh >> 3));
  for (uint32_t c = tid; c < chunks; c += WV_THREADS) {
    const uint32_t* x = sw + c * 4;
    const uint32_t x0 = x[0], x1 = x[1], x2 = x[2], x3 = x[3], x4 = sh ? x[4] : 0u;
    d16[c] = make_uint4(__funnelshift_r(x0, x1, sh), __funnelshift_r(x1, x2, sh), __funnelshift_r(x2, x3, sh),
                        __funnelshift_r(x3, x4, sh));
  }
  const uint32_t done = head + chunks * 16;
  if (tid < len - done) dst[done + tid] = s_src[done + tid];
}

// Lane l of a warp holds the 16-byte pieces wbase/16 + 32 r + l (r = 0 .. 3) of the warp's 2 KB: every
// load of a row is one coalesced 512-byte access (a first version gave each thread 64 consecutive
// bytes: 20 sectors per load request in ncu), zero counts are scanned per row with one packed
// warp scan, and two neighbouring lanes of a row make one rank word.
constexpr int WV_ROWS = WV_BYTES / 16;             // 16-byte pieces per thread

__global__ void __launch_bounds__(WV_THREADS) wavelet_pass_kernel(WaveletPass p) {
  __shared__ __align__(16) uint8_t s_bytes[WV_TILE + 16];
  __shared__ uint32_t s_wz[WV_THREADS / 32];
  __shared__ uint32_t s_tile, s_carry;

  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_tile = atomicAdd(p.ticket, 1u);
  __syncthreads();
  const uint32_t tile = s_tile;
  const uint32_t base = tile * uint32_t(WV_TILE);
  const uint32_t wbase = base + warp * (32 * WV_BYTES);

  // input buffers are padded to a multiple of 64 bytes past n, so a piece that starts below n is readable
  uint32_t w[WV_ROWS][4];
  uint32_t ones[WV_ROWS], nvalid[WV_ROWS];
  uint64_t packed = 0;                               // zero counts of the four pieces, 16 bits each
#pragma unroll
  for (int r = 0; r < WV_ROWS; ++r) {
    const uint32_t pos = wbase + (r * 32 + lane) * 16;
    nvalid[r] = pos >= p.n ? 0u : min(16u, p.n - pos);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (nvalid[r]) v = *reinterpret_cast<const uint4*>(p.in + pos);
    w[r][0] = v.x; w[r][1] = v.y; w[r][2] = v.z; w[r][3] = v.w;
    uint32_t o = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t t = (w[r][q] >> p.bit) & 0x01010101u;
      o |= ((t | (t >> 7) | (t >> 14) | (t >> 21)) & 15u) << (4 * q);
    }
    if (nvalid[r] < 16) o &= (1u << nvalid[r]) - 1u;
    ones[r] = o;
    packed |= uint64_t(nvalid[r] - __popc(o)) << (16 * r);
  }
  uint64_t inc = packed;                             // inclusive scan over the lanes, all four rows at once
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint64_t o = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= unsigned(d)) inc += o;
  }
  const uint64_t rowtot = __shfl_sync(0xffffffffu, inc, 31);
  const uint64_t excl = inc - packed;
  uint32_t zrow[WV_ROWS];                            // zeros of the warp before this lane's piece of row r
  uint32_t acc = 0;
#pragma unroll
  for (int r = 0; r < WV_ROWS; ++r) {
    zrow[r] = acc + (uint32_t(excl >> (16 * r)) & 0xFFFFu);
    acc += uint32_t(rowtot >> (16 * r)) & 0xFFFFu;
  }
  if (lane == 0) s_wz[warp] = acc;
  __syncthreads();
  uint32_t z_warp = 0, tile_zeros = 0;
#pragma unroll
  for (int k = 0; k < WV_THREADS / 32; ++k) {
    const uint32_t z = s_wz[k];
    if (unsigned(k) < warp) z_warp += z;
    tile_zeros += z;
  }
  if (tid < 32) {                        // warp-wide chained scan (a single thread made every tile wait, see profiles/)
    const uint32_t pre = lookback_warp_wide<4>(p.desc, tile, 0u, p.tag, tile_zeros, p.err);
    if (tid == 0) s_carry = pre;
  }
  __syncthreads();
  const uint32_t carry = s_carry;

  // (a) rank words: two neighbouring lanes of a row make one 32-bit data word
#pragma unroll
  for (int r = 0; r < WV_ROWS; ++r) {
    const uint32_t pos = wbase + (r * 32 + lane) * 16;
    const uint32_t hi = __shfl_down_sync(0xffffffffu, ones[r], 1);
    if (!(lane & 1u) && pos <= p.n) {
      const uint32_t z_before = carry + z_warp + zrow[r];              // zeros in [0, pos)
      p.rank[pos / 32] = (uint64_t(ones[r] | (hi << 16)) << 32) | (pos - z_before);
    }
  }
  if (!p.out) return;

  // (b) stable partition through shared memory: zeros of the tile first, then ones
#pragma unroll
  for (int r = 0; r < WV_ROWS; ++r) {
    const uint32_t local = warp * (32 * WV_BYTES) + (r * 32 + lane) * 16;   // tile-local position of the piece
    uint32_t zl = z_warp + zrow[r];
    uint32_t ol = tile_zeros + (local - zl);
    // positions past n sit at the very end of the last tile and are never copied out
    if (nvalid[r] == 16) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint8_t b = uint8_t(w[r][k >> 2] >> (8 * (k & 3)));
        const uint32_t one = (ones[r] >> k) & 1u;
        s_bytes[one ? ol : zl] = b;
        ol += one;
        zl += one ^ 1u;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        if (uint32_t(k) < nvalid[r]) {
          const uint8_t b = uint8_t(w[r][k >> 2] >> (8 * (k & 3)));
          if ((ones[r] >> k) & 1u) s_bytes[ol++] = b; else s_bytes[zl++] = b;
        }
      }
    }
  }
  __syncthreads();
  const uint32_t tile_valid = base >= p.n ? 0u : min(uint32_t(WV_TILE), p.n - base);
  const uint32_t tile_ones = tile_valid - tile_zeros;
  const uint32_t gz = carry;                                // zeros before the tile
  const uint32_t go = p.zeros[p.bit] + (base - carry);      // ones region starts at Z_j
  copy_run(p.out + gz, s_bytes, tile_zeros, tid);
  copy_run(p.out + go, s_bytes + tile_zeros, tile_ones, tid);
}

__global__ void roots_kernel(const uint32_t* __restrict__ zeros, uint32_t* __restrict__ C) {
  // C[i] = zeros of level (i+7)%8, bce.cpp:1128
  if (threadIdx.x < 8) C[threadIdx.x] = zeros[(threadIdx.x + 7) & 7];
}

// Builds Ctx::ranks (8 x words) from Ctx::bwt.  Needs 2 x (n + 64) bytes of scratch.
int wavelet_build(Ctx* c, uint32_t n) {
  cudaStream_t st = c->stream;
  BCE_TRACE("wavelet_build n=%u", n);
  const size_t words = size_t(n) / 32 + 1;
  BCE_TRY(c->ranks.ensure(c, 8 * words * sizeof(uint64_t)));
  const size_t padded = (size_t(n) + 64 + 255) & ~size_t(255);
  BCE_TRY(c->scratch.ensure(c, 2 * padded));
  uint8_t* bufA = c->scratch.as<uint8_t>();
  uint8_t* bufB = bufA + padded;
  const uint32_t tiles = n / WV_TILE + 1;             // position n always falls inside a tile
  BCE_TRY(c->desc.ensure(c, size_t(tiles) * sizeof(uint64_t)));

  char* small = c->small.as<char>();
  uint32_t* d_hist = reinterpret_cast<uint32_t*>(small + kSmallWavelet);
  uint32_t* d_zeros = d_hist + 256;
  uint32_t* d_C = d_zeros + 8;
  uint32_t* d_ticket = d_C + 8;
  uint32_t* d_err = reinterpret_cast<uint32_t*>(small + kSmallErr);
  BCE_CUDA(c, cudaMemsetAsync(d_hist, 0, (256 + 8 + 8 + 8) * 4, st));
  BCE_CUDA(c, cudaMemsetAsync(d_err, 0, 4, st));

  BCE_CUDA(c, cudaEventRecord(c->ev[0], st));
  const uint8_t* L = c->bwt.as<uint8_t>();
  byte_hist_launch(c, L, n, d_hist);
  zeros_from_hist_kernel<<<1, 256, 0, st>>>(d_hist, d_zeros);
  roots_kernel<<<1, 32, 0, st>>>(d_zeros, d_C);
  c->stats.gpu_launches += 3;
  BCE_CUDA(c, cudaGetLastError());

  const uint8_t* in = L;
  for (int j = 0; j < 8; ++j) {
    WaveletPass p;
    p.in = in;
    p.out = j < 7 ? ((j & 1) ? bufB : bufA) : nullptr;
    p.rank = c->ranks.as<uint64_t>() + size_t(j) * words;
    p.n = n; p.bit = j;
    p.zeros = d_zeros;
    p.desc = c->desc.as<uint64_t>();
    p.ticket = d_ticket + j;
    p.tag = uint32_t(next_tag(c));
    p.err = d_err;
    wavelet_pass_kernel<<<tiles, WV_THREADS, 0, st>>>(p);
    c->stats.gpu_launches++;
    BCE_CUDA(c, cudaGetLastError());
    in = p.out;
  }
  uint32_t* h = c->pinned_small.as<uint32_t>() + 8192;
  BCE_CUDA(c, cudaMemcpyAsync(h, d_C, 32, cudaMemcpyDeviceToHost, st));
  BCE_CUDA(c, cudaMemcpyAsync(h + 8, d_err, 4, cudaMemcpyDeviceToHost, st));
  BCE_CUDA(c, cudaEventRecord(c->ev[1], st));
  BCE_CUDA(c, cudaEventSynchronize(c->ev[1]));
  float ms = 0;
  BCE_CUDA(c, cudaEventElapsedTime(&ms, c->ev[0], c->ev[1]));
  c->stats.ms_wavelet += ms;
  if (h[8]) { set_error(c, "wavelet: chained-scan watchdog fired"); return BCE_GPU_E_INTERNAL; }
  for (int i = 0; i < 8; ++i) c->C[i] = h[i];
  c->ranks_resident = true;
  return BCE_GPU_OK;
}

}  // namespace bce
