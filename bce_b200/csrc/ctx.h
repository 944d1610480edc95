// ctx.h -- the context behind bce_gpu_ctx and the internal stage entry points.
#pragma once

#include <stdlib.h>

#include "common.cuh"

struct bce_gpu_ctx {};   // opaque to callers; bce::Ctx derives from it

namespace bce {

constexpr int kMaxSortRounds = 48;

// Knobs read from the environment exist only in experiment builds (-DBCE_GPU_EXPERIMENTS, see
// bce_b200/build.py); the shipped library takes every such decision from the defaults measured in
// profiles/ and never calls getenv on a launch path.
inline size_t exp_env(const char* name, size_t dflt) {
#ifdef BCE_GPU_EXPERIMENTS
  const char* v = getenv(name);
  if (v && *v) return size_t(strtoull(v, nullptr, 10));
#else
  (void)name;
#endif
  return dflt;
}

struct CseDeviceState;   // cse.cu

struct Ctx : bce_gpu_ctx {
  int device = 0;
  int sm_count = 0;
  size_t total_mem = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t h2d_stream = nullptr;        // bce_gpu_prefetch_input: the next input's upload, beside the level loop
  cudaEvent_t ev_prefetch[2] = {};
  const uint8_t* prefetch_src = nullptr;    // what the pending prefetch copies (compared by address and size)
  uint32_t prefetch_n = 0;
  bool prefetch_pending = false;
  cudaEvent_t ev[8] = {};
  cudaEvent_t pass_ev[256] = {};     // start/stop pairs around individual radix passes (resolved lazily)
  int pass_ev_n = 0;
  char err[512] = {0};
  size_t scratch_limit = 0;

  // persistent device data
  DevBuf text;        // T, padded (n + 64)
  DevBuf bwt;         // L, padded
  DevBuf ranks;       // 8 x rank_words u64
  DevBuf scratch;     // carved per stage
  DevBuf small;       // counters, histograms, descriptors' tickets, state structs
  DevBuf desc;        // tile descriptors (tagged, never cleared between passes)
  PinnedBuf pinned_small;   // host mirror for small read-backs
  PinnedBuf pinned_emit;    // emitted counts handed to the caller (two buffers alternate)
  PinnedBuf pinned_emit2;
  PinnedBuf pinned_io;      // staging for pageable caller buffers
  DevBuf scan_tmp;          // bce -s: sort buffers, bucketed symbols and bucket tables of one batch
  DevBuf pack_tmp;          // a batch's words as 20 bits each, on their way to the host (bce_gpu_cse_next_words20)

  uint32_t desc_tag = 0;    // monotonically increasing pass tag (30 bits used)

  // options (bce_gpu_set_option)
  size_t emit_batch_bytes = size_t(1) << 30;   // BCE_GPU_OPT_EMIT_BATCH_BYTES
  uint32_t local_sort_min = 1u << 20;          // BCE_GPU_OPT_LOCAL_SORT_MIN
  bool resident_checksum = false;              // BCE_GPU_OPT_RESIDENT_CHECKSUM
  uint64_t slot_enter_nodes = 2000000;         // BCE_GPU_OPT_SLOT_ENTER_NODES
  uint64_t mid_enter_nodes = 400000;           // BCE_GPU_OPT_MID_ENTER_NODES
  bool no_narrow_kernels = false;              // BCE_GPU_OPT_NO_NARROW_KERNELS

  // state of the current input
  uint32_t n = 0;
  bool text_resident = false;
  bool bwt_resident = false;
  bool ranks_resident = false;
  uint32_t offset = 0;
  uint32_t C[8] = {};

  // emission format for the next cse_begin (bce_gpu_set_emit_mode)
  uint32_t emit_mode = 0;
  uint8_t emit_cfg[8][32] = {};
  bool cse_resident = false;   // counts stay in device memory (front_resident)
  int cse_emit_mode_active = 0; // mode the current run was started with

  // CSE run state (host side)
  bool cse_active = false;
  bool cse_done = false;
  struct CseHost* cse = nullptr;

  bce_gpu_stats stats = {};
};

// layout of Ctx::small (device) in bytes; Ctx::pinned_small mirrors it on the host
constexpr size_t kSmallHist = 0;                              // 8 x 256 u32  digit histograms
constexpr size_t kSmallBase = kSmallHist + 8 * 256 * 4;       // 8 x 256 u32  digit starts
constexpr size_t kSmallTicket = kSmallBase + 8 * 256 * 4;     // 16 u32       radix tile dispensers
constexpr size_t kSmallErr = kSmallTicket + 64;               // 1 u32        chained-scan watchdog flag
constexpr size_t kSmallRerankTicket = kSmallErr + 64;         // 1 u32
constexpr size_t kSmallRerankTotals = kSmallRerankTicket + 64;  // 2 u32
constexpr size_t kSmallWavelet = kSmallRerankTotals + 64;     // 256 u32 hist + 8 u32 zeros + 8 tickets
constexpr size_t kSmallCse = kSmallWavelet + 2048;            // CseDeviceState
constexpr size_t kSmallUnbwt = kSmallCse + 1024;              // inverse-BWT counters
constexpr size_t kSmallChecksum = 44 * 1024;                  // 8 x 2 u64    per-stream checksums of the resident emission
constexpr size_t kSmallRadixCursor = 52 * 1024;               // 256 u32      write cursors of a sort's first (unordered) pass
constexpr size_t kSmallPartHist = 48 * 1024;                  // 256 u32      histogram of the rank-scatter partition digit
constexpr size_t kSmallBytes = 64 * 1024;

// ---- stage entry points (each launches kernels on ctx->stream) --------------------
// radix_sort.cu: LSD radix sort of (u64 key, u32 value) pairs on digit shifts[0..npass).
// Buffers are ping-ponged; *out_k / *out_v receive the pointers that hold the result.
// Where the digit histograms of a sort come from (default: one extra read of the keys).
struct RadixHistSource {
  const uint8_t* window_text = nullptr;   // keys are the 8-byte cyclic windows of this text: its byte histogram serves every pass
  bool keys_from_text = false;            // ... and keyA / valA are NOT filled: the first pass makes keys and indices from the
                                          // text (if no pass runs at all -- one repeated byte -- the caller has to pack them)
  const uint32_t* dev_hist = nullptr;     // [npass][256] already counted on the device (e.g. by the kernel that wrote the keys);
                                          // may be Ctx::small + kSmallHist itself
  const uint32_t* host_hist = nullptr;    // [npass][256] known on the host
  bool stable_first = false;              // every pass stable, the first included: equal keys keep their input order
                                          // (the suffix sorter does not need that: ties are told apart by later rounds)
};
int radix_sort_pairs(Ctx* c, uint64_t* keyA, uint64_t* keyB, uint32_t* valA, uint32_t* valB,
                     uint32_t m, const int* shifts, int npass, uint64_t** out_k, uint32_t** out_v,
                     int* passes_run, const RadixHistSource* src = nullptr);
constexpr int kRadixMaxPasses = 8;
struct RadixShifts { int s[kRadixMaxPasses]; };
size_t radix_desc_words(uint32_t m);

// suffix_sort.cu
int suffix_sort_bwt(Ctx* c, uint32_t n, uint32_t* sa_host_or_null);
// wavelet.cu
int wavelet_build(Ctx* c, uint32_t n);
// cse.cu
int cse_begin(Ctx* c, uint32_t n);
struct CseWordBatch { const uint32_t* words[8]; size_t count[8]; int done; };
// pack20: out->words[i] points at 20-bit words, two in 5 bytes (BCE_EMIT_CODER: every word is below 2^20)
int cse_advance(Ctx* c, bool resident, CseWordBatch* out, bool pack20 = false);
int cse_advance_buckets(Ctx* c, bce_scan_buckets* out);
void cse_destroy(Ctx* c);
// unbwt.cu
int unbwt_run(Ctx* c, uint32_t offset, uint32_t n, uint8_t* out_host);

// api.cu helpers
int next_tag(Ctx* c);
bool is_pinned(const void* p);
int h2d(Ctx* c, void* dst, const void* src, size_t bytes);
int d2h(Ctx* c, void* dst, const void* src, size_t bytes);
int radix_init_device(Ctx* c);   // radix_sort.cu: per-device function attributes
size_t scratch_budget(Ctx* c);

}  // namespace bce
