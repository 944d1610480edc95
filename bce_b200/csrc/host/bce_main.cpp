// bce_main.cpp -- the `bce` command line tool with the reference's interface
// (main, bce.cpp:1376-1484): same arguments, same messages, same exit codes, same archive
// bytes -- but RankFile + BCE::code run on the GPU through include/bce_gpu.h.
//
//   bce -c archive.bce file [config.bcc]      compress
//   bce -s config.bcc file                    scan: derive a coder config
//   bce -d file archive.bce                   decompress   (see decode.hpp)
//   bce -ds file archive.bce                  decompress, serial inverse BWT on the host
//
// There is no CPU fallback for the front end: without a usable sm_100 device the tool
// reports the CUDA error and exits with -3.
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "../../../include/bce_host.h"
#include "coders.hpp"
#ifdef BCE_HAVE_DECODER
#include "decode.hpp"
#endif

namespace {

// File::File, bce.cpp:842-856: the whole file in one buffer -- page-locked here, so that the upload to the
// device is one DMA transfer straight out of it (pageable memory would go through staged copies).
struct InputFile {
  bce_gpu_ctx* ctx = nullptr;
  uint8_t* data = nullptr;
  size_t size = 0;
  bool pinned = false;
  ~InputFile() { release(); }
  void release() {
    if (data) { if (pinned) bce_gpu_host_free(ctx, data); else std::free(data); }
    data = nullptr;
  }
  // 0 = ok, 1 = cannot read / empty, 2 = too large for the format
  int load(const std::string& path, bce_gpu_ctx* c) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f) return 1;
    const std::streamoff sz = f.tellg();
    if (sz <= 0) return 1;
    // the format carries n as a varint of at most 31 binary digits (bce.cpp:372-377) and every index is a uint32:
    // the reference silently wraps on larger files, this tool refuses them
    if (uint64_t(sz) > 0x7FFFFFFFull) {
      std::printf("File too large: %" PRIuMAX " B (the BCE v0.4 format holds at most 2147483647 B)\n", uintmax_t(sz));
      return 2;
    }
    ctx = c;
    size = size_t(sz);
    data = c ? static_cast<uint8_t*>(bce_gpu_host_alloc(c, size)) : nullptr;
    pinned = data != nullptr;
    if (!data) data = static_cast<uint8_t*>(std::malloc(size));
    if (!data) return 1;
    f.seekg(0, std::ios::beg);
    return f.read(reinterpret_cast<char*>(data), sz) ? 0 : 1;
  }
};

bool exists_nonempty(const std::string& path) {
  std::ifstream f(path, std::ios::binary | std::ios::ate);
  return f && f.tellg() > 0;
}

int gpu_failure(bce_gpu_ctx* ctx, int rc) {
  std::printf("GPU front end failed: %s (%s)\n", bce_gpu_error_string(rc), ctx ? bce_gpu_last_error(ctx) : "no context");
  return rc;
}

void usage() {                                                           // bce.cpp:1474-1482
  std::printf("Usage:\n");
  std::printf("  bce -c archive.bce file [config.bcc]\n");
  std::printf("   Compresses \"file\" to archive \"archive.bce\" [using config \"config.bcc\"]\n");
  std::printf("\n");
  std::printf("  bce -d file archive.bce\n");
  std::printf("   Decompresses archive \"archive.bce\" to \"file\"\n");
  std::printf("\n");
  std::printf("  bce -s config.bcc file\n");
  std::printf("   Scan \"file\" and generate a config file \"config.bcc\" to improve the AdaptiveCoder (uses a lot of memory)\n");
}

}  // namespace

int main(int argc, char** argv) {
  std::printf("BCE v0.4 Release (B200 front end)\n");
  std::printf("Archive format and coders of BCE v0.4, Copyright (C) 2016  Christoph Diegelmann\n\n");

  const bool flag = argc >= 2 && argv[1][0] == '-';
  using clock = std::chrono::high_resolution_clock;

  if (argc == 4 && flag && argv[1][1] == 's') {                          // bce.cpp:1384-1402
    const auto start = clock::now();
    if (!exists_nonempty(argv[3])) {                                     // before the device is touched (:1390-1393)
      std::printf("Error loading file\n");
      return -1;
    }
    bce_gpu_ctx* ctx = nullptr;
    int rc = bce_gpu_open(0, &ctx);
    if (rc) return gpu_failure(ctx, rc);
    InputFile data;
    if (int lr = data.load(argv[3], ctx)) {
      if (lr == 1) std::printf("Error loading file\n");
      data.release();
      bce_gpu_close(ctx);
      return -1;
    }
    uint8_t cfg[288];
    rc = bce_scan_buffer(ctx, data.data, uint32_t(data.size), cfg);
    if (rc) { gpu_failure(ctx, rc); data.release(); bce_gpu_close(ctx); return rc; }
    const size_t scanned = data.size;
    data.release();
    bce_gpu_close(ctx);
    bcehost::ConfigTable table;
    std::memcpy(table.data(), cfg, 288);
    bcehost::save_config_file(argv[2], table);
    const std::chrono::duration<double> d = clock::now() - start;
    std::printf("Scanned %" PRIuMAX " B in %.1f s\n", uintmax_t(scanned), d.count());
    return 0;
  }

  if ((argc == 4 || argc == 5) && flag && argv[1][1] == 'c') {           // bce.cpp:1403-1427
    const auto start = clock::now();
    bcehost::ConfigTable table = bcehost::default_config();
    if (argc == 5) bcehost::load_config_file(argv[4], table);            // failure is non-fatal (:629-632)
    if (!exists_nonempty(argv[3])) {                                     // before the device is touched (:1412-1415)
      std::printf("Error loading file\n");
      return -1;
    }
    const bool timing = std::getenv("BCE_TIME") != nullptr;
    bce_gpu_ctx* ctx = nullptr;
    int rc = bce_gpu_open(0, &ctx);
    if (rc) return gpu_failure(ctx, rc);
    const auto t_open = clock::now();
    InputFile data;
    if (int lr = data.load(argv[3], ctx)) {
      if (lr == 1) std::printf("Error loading file\n");
      data.release();
      bce_gpu_close(ctx);
      return -1;
    }
    if (timing)
      std::fprintf(stderr, "[bce] open device %.3f s | read into page-locked memory %.3f s\n",
                   std::chrono::duration<double>(t_open - start).count(), std::chrono::duration<double>(clock::now() - t_open).count());
    uint16_t* words = nullptr;
    size_t nwords = 0;
    rc = bce_compress_buffer(ctx, data.data, uint32_t(data.size),
                             reinterpret_cast<const uint8_t*>(table.data()), 8, &words, &nwords);
    if (rc) { gpu_failure(ctx, rc); data.release(); bce_gpu_close(ctx); return rc; }
    const std::chrono::duration<double> d = clock::now() - start;
    std::printf("Compressed from %" PRIuMAX " B -> %zu B in %.1f s\n", uintmax_t(data.size),
                nwords * sizeof(uint16_t), d.count());
    const auto t_write = clock::now();
    {
      std::ofstream archive(argv[2], std::ios::binary | std::ios::trunc);
      archive.write(reinterpret_cast<const char*>(words), std::streamsize(nwords * sizeof(uint16_t)));
    }
    const auto t_close = clock::now();
    if (timing)
      std::fprintf(stderr, "[bce] write %.3f s\n", std::chrono::duration<double>(t_close - t_write).count());
    // The archive is on disk.  Unmapping gigabytes of pinned and device memory one allocation at a
    // time costs up to seconds; leaving that to process exit is what a one-shot tool wants
    // (BCE_CLEAN_EXIT=1 keeps the orderly teardown for leak checkers).
    if (std::getenv("BCE_CLEAN_EXIT")) {
      bce_host_free(words);
      data.release();
      bce_gpu_close(ctx);
      return 0;
    }
    std::fflush(stdout);
    std::fflush(stderr);
    std::_Exit(0);
  }

  if (argc == 4 && flag && argv[1][1] == 'd') {                          // bce.cpp:1428-1472
#ifdef BCE_HAVE_DECODER
    const auto start = clock::now();
    std::ifstream archive(argv[3], std::ios::binary | std::ios::ate);
    const std::streamoff size = archive ? std::streamoff(archive.tellg()) : -1;
    if (size < 0) {
      std::printf("Archive not found.\n");
      return -1;
    }
    archive.seekg(0, std::ios::beg);
    // an archive is a whole number of 16-bit words (bce.cpp:1424-1427): an odd or empty size is a damaged
    // file (the reference reads `size` bytes into size/2 words and overruns by one, bce.cpp:1441-1445)
    std::vector<uint16_t> words(size_t(size) / sizeof(uint16_t));
    if (size < 2 || (size & 1) || !archive.read(reinterpret_cast<char*>(words.data()), size)) {
      std::printf("Could not read Archive.\n");
      return -2;
    }
    std::vector<uint8_t> out;
    const int rc = bcehost::decode_archive(words, argv[1][2] == 's', out);
    if (rc) {
      std::printf("Decoding failed: %s%s\n", bce_gpu_error_string(rc),
                  argv[1][2] == 's' ? "" : " (bce -ds decodes without a GPU)");
      return rc;
    }
    const std::chrono::duration<double> d = clock::now() - start;
    std::printf("Decompressed from %zu B -> %zu B in %.1f s\n", size_t(size), out.size(), d.count());
    std::ofstream file(argv[2], std::ios::binary | std::ios::trunc);
    file.write(reinterpret_cast<const char*>(out.data()), std::streamsize(out.size()));
    return 0;
#else
    std::printf("This build carries the compression front end only; decode with the reference's bce -d\n"
                "(INTEGRATION.md shows its unbwt policy bound to bce_gpu_unbwt).\n");
    return -4;
#endif
  }

  usage();
  return 0;
}
