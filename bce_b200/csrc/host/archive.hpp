// archive.hpp -- archive writer: 8 stream encoders + header, BCE::encode (bce.cpp:1117-1167).
#pragma once

#include <memory>
#include <vector>

#include "../../../include/bce_gpu.h"
#include "coders.hpp"

namespace bcehost {

class ArchiveWriter {
 public:
  ArchiveWriter(uint32_t n, const uint32_t C[8], const ConfigTable& cfg);
  // codes the batch; with threads > 1 every stream runs on its own thread
  void feed(const bce_cse_batch& batch, int threads);
  void feed_words(const bce_cse_words& batch, int threads);      // BCE_EMIT_CODER batches
  std::vector<uint16_t> finish(uint32_t offset);

 private:
  uint32_t n_;
  ConfigTable cfg_;
  std::vector<std::unique_ptr<StreamEncoder>> streams_;
};

class ScanSession {
 public:
  ScanSession();
  void feed(const bce_cse_batch& batch);
  void feed_words(const bce_cse_words& batch);                    // BCE_EMIT_SCAN batches
  ConfigTable finish();

 private:
  std::vector<std::unique_ptr<ScanCollector>> streams_;
};

}  // namespace bcehost
