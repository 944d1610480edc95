/*
 * oracle/ref_gpu_binding.cpp -- TEST INFRASTRUCTURE: INTEGRATION.md, compiled.
 *
 * The UNMODIFIED reference translation unit (/root/reference/bce.cpp, found through -I, never
 * copied into this repo) with the few lines a maintainer adds to bind libbce_gpu.so
 * (include/bce_gpu.h) at its two seams:
 *
 *   compress   GpuFrontEnd replaces RankFile (bce.cpp:932-984): File::rotate + File::bwt (divbwt,
 *              :901) + the wavelet build run on the device; encode_gpu is BCE::encode
 *              (:1117-1167) with `code(coder_, C, file.ranks, n, 1)` (:1132) replaced by batches
 *              from bce_gpu_cse_next handed to the reference's OWN AdaptiveCoder<31>::set (:1302).
 *              Flush, header and concatenation (:1134-1157) are the reference's statements.
 *   decompress unbwt::gpu is one more policy_unbwt next to unbwt::bytewise (:1041-1103): the
 *              reference's BCE::decode (:1169-1233) runs unchanged and hands its rank
 *              dictionaries to bce_gpu_unbwt instead of inverse_bw_transform (:1091).
 *
 * tests/test_gpu_integration.py runs the resulting tool (`oracle/_ref/bce_ref_gpu -c / -d`) and
 * compares its archives byte for byte with `bce_ref -c` and with this repository's own `bce`.
 * Rank keeps its word array private (bce.cpp:126-221); the accessor INTEGRATION.md adds to it
 * is emulated here by compiling the reference with `private` visible (no source edit).
 */
#define main bce_reference_main
#define private public
#include "bce.cpp"
#undef private
#undef main

extern "C" {
#include "bce_gpu.h"
}

#include <cstdlib>

namespace {

/* replaces RankFile: nothing but the file bytes stays on the host */
struct GpuFrontEnd : File {
  explicit GpuFrontEnd(const std::string& path) : File(path) {}
  bce_gpu_ctx* ctx = nullptr;
  uint32_t C[8];
  int open() {
    if (int rc = bce_gpu_open(0, &ctx)) return rc;
    /* rotate() + bwt() + wavelet build + root set-up, all on the device */
    return bce_gpu_compress_front(ctx, map_.data(), uint32_t(size_), &offset_, C);
  }
  ~GpuFrontEnd() { bce_gpu_close(ctx); }
};

/* BCE::encode (bce.cpp:1117-1167) over the device's batches */
template <class coder_type>
typename coder_type::value_type encode_gpu(GpuFrontEnd& file) {
  auto n = file.size();
  std::array<coder_type, 8> coder_ = {0, 1, 2, 3, 4, 5, 6, 7};       /* :1124 */
  for (int i = 0; i < 8; ++i) coder_[i].set(file.C[i], n + 1);       /* :1129 */

  bce_cse_batch batch;                                                /* replaces code(...), :1132 */
  do {
    if (bce_gpu_cse_next(file.ctx, &batch) != BCE_GPU_OK) {
      printf("GPU front end failed: %s\n", bce_gpu_last_error(file.ctx));
      std::exit(3);
    }
#ifdef _OPENMP
    #pragma omp parallel for                                          /* streams are independent (:1250) */
#endif
    for (int i = 0; i < 8; ++i)
      for (size_t j = 0; j < batch.count[i]; ++j) {
        const bce_tuple& t = batch.tuples[i][j];
        coder_[i].set(t.sym, t.k, t.c1, t.c2, t.cs);                  /* :1302 */
      }
  } while (!batch.done);

  auto size = 0u;                                                     /* :1134-1138 */
  for (int i = 0; i < 8; ++i) {
    coder_[i].flush();
    size += coder_[i].data().size();
  }
  coder_type head(-1);                                                /* :1141-1149 */
  head.setv(n);
  head.set(file.offset(), n + 1);
  head.setv(size);
  for (int i = 0, s = size; i < 7; ++i) {
    head.set(coder_[i].data().size(), s + 1);
    s -= coder_[i].data().size();
  }
  head.flush();
  typename coder_type::value_type data;                               /* :1152-1157 */
  data.push_back(head.data().size());
  data.insert(data.end(), head.data().begin(), head.data().end());
  for (int i = 0; i < 8; ++i) data.insert(data.end(), coder_[i].data().begin(), coder_[i].data().end());
  return data;
}

}  // namespace

namespace unbwt {
/* one more policy_unbwt (the reference's are bitwise :997-1039, bytewise :1041-1103, noop) */
class gpu {
 public:
  std::vector<uint8_t> unbwt(std::array<Rank, 8>& ranks, uint32_t offset, uint32_t n) {
    std::vector<uint8_t> out(n);
    const uint64_t* lv[8];
    for (int j = 0; j < 8; ++j) lv[j] = ranks[j].rank_.data();       /* Rank::words() in INTEGRATION.md */
    bce_gpu_ctx* ctx = nullptr;
    if (bce_gpu_open(0, &ctx) || bce_gpu_unbwt(ctx, lv, offset, n, out.data())) {
      printf("GPU inverse BWT failed: %s\n", ctx ? bce_gpu_last_error(ctx) : "no device");
      std::exit(3);
    }
    bce_gpu_close(ctx);
    return out;
  }
};
}  // namespace unbwt

/* main (bce.cpp:1376-1484) reduced to the two paths that change: -c and -d */
int main(int argc, char** argv) {
  using coder_type = AdaptiveCoder<31>;
  if ((argc == 4 || argc == 5) && argv[1][0] == '-' && argv[1][1] == 'c') {       /* :1403-1427 */
    if (argc == 5) coder_type::load_config(argv[4]);
    GpuFrontEnd file{std::string(argv[3])};
    if (file.status()) { printf("Error loading file\n"); return -1; }
    if (int rc = file.open()) { printf("GPU front end failed: %d\n", rc); return 3; }
    auto data = encode_gpu<coder_type>(file);
    std::ofstream archive(argv[2], std::ios::binary | std::ios::trunc);
    archive.write(reinterpret_cast<const char*>(data.data()), data.size() * sizeof(coder_type::value_type::value_type));
    printf("Compressed from %" PRIuMAX " B -> %" PRIuMAX " B\n", uintmax_t(file.size()),
           uintmax_t(data.size() * sizeof(coder_type::value_type::value_type)));
    return 0;
  }
  if (argc == 4 && argv[1][0] == '-' && argv[1][1] == 'd') {                       /* :1428-1472 */
    std::ifstream archive(argv[3], std::ios::binary | std::ios::ate);
    if (!archive) { printf("Archive not found.\n"); return -1; }
    const std::streamsize size = archive.tellg();
    archive.seekg(0, std::ios::beg);
    coder_type::value_type data(size / sizeof(coder_type::value_type::value_type));
    if (!archive.read(reinterpret_cast<char*>(data.data()), size)) { printf("Could not read Archive.\n"); return -2; }
    BCE<coder_type, unbwt::gpu> bce;                                                /* :1470 with the new policy */
    auto out = bce.decode(data);
    std::ofstream file(argv[2], std::ios::binary | std::ios::trunc);
    file.write(reinterpret_cast<const char*>(out.data()), out.size());
    printf("Decompressed from %" PRIuMAX " B -> %" PRIuMAX " B\n", uintmax_t(size), uintmax_t(out.size()));
    return 0;
  }
  printf("usage: bce_ref_gpu -c archive file [config] | -d file archive\n");
  return 0;
}
