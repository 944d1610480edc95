/*
 * oracle/ref_tap.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * Compiles the UNMODIFIED reference translation unit (/root/reference/bce.cpp,
 * found through -I, never copied into this repo) into a shared library and taps
 * it through its own extension seam: the `policy_coder` template parameter of
 * `BCE<>` (bce.cpp:1111).  A recording coder receives exactly the
 * (s, k, c1, c2, cs) calls the real AdaptiveCoder would (bce.cpp:1302, :1129),
 * so the streams below ARE the reference's output, not a restatement of it.
 * A File subclass exposes the protected BWT buffer (bce.cpp:921).
 *
 * Output of the build goes to oracle/_ref/ only (see oracle/Makefile).
 */
#define main bce_reference_main
#include "bce.cpp"
#undef main

#include <cstring>
#include <string>

namespace {

struct TapStore {
  std::vector<uint32_t> adaptive[9];   /* 5 words per call of set(s,k,c1,c2,cs) */
  std::vector<uint32_t> uniform[9];    /* 2 words per call of set(s,k)          */
  bool record = true, checksum = true;
  uint64_t calls[9] = {0};
  /* order-sensitive checksum of the adaptive calls of a stream, kept when asked for (flag 8) so that
   * inputs too large to record (1 GB: 44 GB of tuples) are still pinned call by call:
   *   h_j = ((((s*A + k)*A + c1)*A + c2)*A + cs)  mod 2^64,  A = 0x9E3779B97F4A7C15
   *   sum = SUM_j h_j,   wsum = SUM_j h_j * (2 j + 1)       (j = 0-based call index)      */
  uint64_t sum[9] = {0}, wsum[9] = {0};
  void reset() {
    for (auto& v : adaptive) std::vector<uint32_t>().swap(v);
    for (auto& v : uniform) std::vector<uint32_t>().swap(v);
    std::memset(calls, 0, sizeof calls);
    std::memset(sum, 0, sizeof sum);
    std::memset(wsum, 0, sizeof wsum);
  }
};
TapStore g_tap;

/* satisfies the coder policy listed in SURVEY.md 8b */
class TapCoder : public VCoder<TapCoder> {
 public:
  using value_type = std::vector<uint16_t>;
  static constexpr const int max = 31;

  TapCoder(int i) : id_(i < 0 || i > 7 ? 8 : i) {}
  explicit TapCoder(int i, value_type&&) : id_(i < 0 || i > 7 ? 8 : i) {}

  void set(uint32_t s, uint32_t k, uint32_t c1, uint32_t c2, uint32_t cs) {
    if (g_tap.checksum) {
      const uint64_t A = 0x9E3779B97F4A7C15ull;
      uint64_t h = ((((uint64_t(s) * A + k) * A + c1) * A + c2) * A + cs);
      g_tap.sum[id_] += h;
      g_tap.wsum[id_] += h * (2 * g_tap.calls[id_] + 1);
    }
    g_tap.calls[id_]++;
    if (!g_tap.record) return;
    auto& v = g_tap.adaptive[id_];
    v.push_back(s); v.push_back(k); v.push_back(c1); v.push_back(c2); v.push_back(cs);
  }
  void set(uint32_t s, uint32_t k) {   /* a handful of calls per run (roots, header): always kept */
    auto& v = g_tap.uniform[id_];
    v.push_back(s); v.push_back(k);
  }
  uint32_t get(uint32_t, uint32_t, uint32_t, uint32_t) { return 0; }
  uint32_t get(uint32_t) { return 0; }
  void flush() {}
  const value_type& data() const { return data_; }
  void clear() {}
  static void load_config(std::string) {}

 private:
  int id_;
  value_type data_;
};

/* File's buffer is protected (bce.cpp:920-924): a subclass may read it */
struct BwtFile : File {
  explicit BwtFile(const std::string& p) : File(p) {
    if (status_ == 0) { rotate(); bwt(); }
  }
  const std::vector<unsigned char>& bytes() const { return map_; }
};

uint32_t g_n = 0, g_offset = 0;
std::vector<uint8_t> g_bwt;
std::vector<uint64_t> g_ranks[8];
double g_t_front = 0, g_t_encode = 0;

}  // namespace

extern "C" {

/* Run the reference front end on a file.  flags: 1 = also keep the BWT bytes,
 * 2 = also keep the 8 rank arrays (rebuilt through Rank's public get/bit),
 * 4 = do not record tuples (timing runs: only count them), 8 = keep the per-stream call
 * checksums (always on when recording). Returns 0 / -1. */
int bce_ref_front(const char* path, int flags) {
  g_tap.reset();
  g_tap.record = !(flags & 4);
  g_tap.checksum = g_tap.record || (flags & 8);
  g_bwt.clear();
  for (auto& r : g_ranks) r.clear();

  if (flags & 1) {
    BwtFile f{std::string(path)};
    if (f.status()) return -1;
    g_bwt.assign(f.bytes().begin(), f.bytes().end());
  }
  auto t0 = std::chrono::high_resolution_clock::now();
  RankFile file{std::string(path)};                       /* bce.cpp:1411 */
  if (file.status()) return -1;
  auto t1 = std::chrono::high_resolution_clock::now();
  g_n = static_cast<uint32_t>(file.size());
  g_offset = static_cast<uint32_t>(file.offset());
  if (flags & 2) {
    for (int j = 0; j < 8; ++j) {
      size_t words = g_n / 32 + 1;
      g_ranks[j].resize(words);
      for (size_t w = 0; w < words; ++w) {
        uint64_t bits = 0;
        for (uint32_t b = 0; b < 32; ++b) {
          uint64_t p = w * 32 + b;
          if (p < g_n) bits |= static_cast<uint64_t>(file.ranks[j].bit(static_cast<uint32_t>(p))) << b;
        }
        uint32_t before = file.ranks[j].get<1>(static_cast<uint32_t>(w * 32 <= g_n ? w * 32 : g_n));
        g_ranks[j][w] = (bits << 32) | before;
      }
    }
  }
  auto t2 = std::chrono::high_resolution_clock::now();
  BCE<TapCoder, unbwt::noop> bce;
  bce.encode(file);                                       /* bce.cpp:1417 */
  auto t3 = std::chrono::high_resolution_clock::now();
  g_t_front = std::chrono::duration<double>(t1 - t0).count();
  g_t_encode = std::chrono::duration<double>(t3 - t2).count();
  return 0;
}

uint32_t bce_ref_n(void) { return g_n; }
uint32_t bce_ref_offset(void) { return g_offset; }
const uint8_t* bce_ref_bwt(void) { return g_bwt.data(); }
const uint64_t* bce_ref_ranks(int level) { return g_ranks[level].data(); }
double bce_ref_seconds_rankfile(void) { return g_t_front; }   /* rotate + BWT + wavelet */
double bce_ref_seconds_encode(void) { return g_t_encode; }    /* CSE loop with the tap coder */
uint64_t bce_ref_calls(int stream) { return g_tap.calls[stream]; }
uint64_t bce_ref_checksum(int stream, int weighted) { return weighted ? g_tap.wsum[stream] : g_tap.sum[stream]; }

/* stream 0..7 = wavelet levels, 8 = header coder */
const uint32_t* bce_ref_adaptive(int stream, size_t* calls) {
  *calls = g_tap.adaptive[stream].size() / 5;
  return g_tap.adaptive[stream].data();
}
const uint32_t* bce_ref_uniform(int stream, size_t* calls) {
  *calls = g_tap.uniform[stream].size() / 2;
  return g_tap.uniform[stream].data();
}

}  // extern "C"
