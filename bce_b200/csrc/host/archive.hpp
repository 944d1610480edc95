// archive.hpp -- archive writer: 8 stream encoders + header, BCE::encode (bce.cpp:1117-1167).
#pragma once

#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
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
  // Overlapped form: begin_words hands the batch to one persistent coder thread per stream
  // and returns at once -- the reference forks its eight coders inside the level loop the same way
  // (bce.cpp:1250-1252, :1302) -- so the caller can fetch the next batch from the GPU meanwhile; wait_words
  // blocks until the batch is coded.  The batch's memory must stay valid until then.
  void begin_words(const bce_cse_words& batch);
  void begin_words20(const bce_cse_words20& batch);               // the same words as 20 bits each
  void wait_words();
  double busy_seconds(int stream) const { return busy_[stream]; }   // time stream's coder spent coding so far
  std::vector<uint16_t> finish(uint32_t offset);
  ~ArchiveWriter();

 private:
  void worker(int stream);
  uint32_t n_;
  ConfigTable cfg_;
  std::vector<std::unique_ptr<StreamEncoder>> streams_;
  // coder threads (started by the first begin_words)
  std::vector<std::thread> pool_;
  std::mutex mu_;
  std::condition_variable cv_work_, cv_done_;
  const void* job_words_[8] = {};
  bool job_20_[8] = {};
  size_t job_count_[8] = {};
  bool job_ready_[8] = {};
  int jobs_open_ = 0;
  bool stop_ = false;
  double busy_[8] = {};
};

class ScanSession {
 public:
  ScanSession();
  void feed(const bce_cse_batch& batch);
  void feed_words(const bce_cse_words& batch);                    // BCE_EMIT_SCAN batches
  void feed_buckets(const bce_scan_buckets& batch);               // batches bucketed on the device, one thread per stream
  ConfigTable finish();

 private:
  std::vector<std::unique_ptr<ScanCollector>> streams_;
};

}  // namespace bcehost
