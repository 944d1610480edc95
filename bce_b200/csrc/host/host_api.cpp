// host_api.cpp -- extern "C" surface of libbce_host (include/bce_host.h).
#include <algorithm>
#include <cstdlib>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <new>

#include "../../../include/bce_host.h"
#include "archive.hpp"
#include "decode.hpp"

using namespace bcehost;

struct bce_archive_writer { ArchiveWriter* w; };
struct bce_scan { ScanSession* s; };

static ConfigTable table_from(const uint8_t* cfg288) {
  ConfigTable t = default_config();
  if (cfg288) std::memcpy(t.data(), cfg288, 288);
  return t;
}

// The C ABI never lets an exception out: allocation failures become BCE_GPU_E_NOMEM.
#define BCE_HOST_GUARD(body)                                   \
  try { body }                                                 \
  catch (const std::bad_alloc&) { return BCE_GPU_E_NOMEM; }    \
  catch (...) { return BCE_GPU_E_INTERNAL; }

static bool config_ok(const uint8_t* cfg288) {                 // context bits are coded in 0..5 (bce.cpp:686)
  if (!cfg288) return true;
  for (int i = 0; i < 288; ++i) if (cfg288[i] > 5) return false;
  return true;
}

extern "C" {

const uint8_t* bce_host_default_config(void) {
  return reinterpret_cast<const uint8_t*>(default_config().data());
}
void bce_host_free(void* p) { std::free(p); }

bce_archive_writer* bce_archive_begin(uint32_t n, const uint32_t C[8], const uint8_t* cfg288) {
  if (!C || n == 0 || n >= 0x80000000u || !config_ok(cfg288)) return nullptr;
  auto* h = new (std::nothrow) bce_archive_writer{nullptr};
  if (!h) return nullptr;
  try { h->w = new ArchiveWriter(n, C, table_from(cfg288)); }
  catch (...) { h->w = nullptr; }
  if (!h->w) { delete h; return nullptr; }
  return h;
}
int bce_archive_feed(bce_archive_writer* h, const bce_cse_batch* batch, int threads) {
  if (!h || !batch) return BCE_GPU_E_ARG;
  BCE_HOST_GUARD(h->w->feed(*batch, threads);)
  return BCE_GPU_OK;
}
int bce_archive_feed_words(bce_archive_writer* h, const bce_cse_words* batch, int threads) {
  if (!h || !batch) return BCE_GPU_E_ARG;
  BCE_HOST_GUARD(h->w->feed_words(*batch, threads);)
  return BCE_GPU_OK;
}
int bce_archive_begin_words(bce_archive_writer* h, const bce_cse_words* batch) {
  if (!h || !batch) return BCE_GPU_E_ARG;
  BCE_HOST_GUARD(h->w->begin_words(*batch);)
  return BCE_GPU_OK;
}
int bce_archive_begin_words20(bce_archive_writer* h, const bce_cse_words20* batch) {
  if (!h || !batch) return BCE_GPU_E_ARG;
  BCE_HOST_GUARD(h->w->begin_words20(*batch);)
  return BCE_GPU_OK;
}
int bce_archive_wait(bce_archive_writer* h) {
  if (!h) return BCE_GPU_E_ARG;
  h->w->wait_words();
  return BCE_GPU_OK;
}
int bce_scan_feed_words(bce_scan* h, const bce_cse_words* batch) {
  if (!h || !batch) return BCE_GPU_E_ARG;
  BCE_HOST_GUARD(h->s->feed_words(*batch);)
  return BCE_GPU_OK;
}
int bce_scan_feed_buckets(bce_scan* h, const bce_scan_buckets* batch) {
  if (!h || !batch) return BCE_GPU_E_ARG;
  BCE_HOST_GUARD(h->s->feed_buckets(*batch);)
  return BCE_GPU_OK;
}
size_t bce_host_pack_counts(int mode, const uint8_t* cfg288, int stream, const bce_tuple* t, size_t count, uint32_t* words) {
  if (!config_ok(cfg288) || stream < 0 || stream > 7) return 0;
  ConfigTable tab = table_from(cfg288);
  size_t at = 0;
  for (size_t i = 0; i < count; ++i)
    at += pack_count(mode, tab[stream].data(), t[i].sym, t[i].k, t[i].c1, t[i].c2, t[i].cs, words + at);
  return at;
}
int bce_archive_finish(bce_archive_writer* h, uint32_t offset, uint16_t** words, size_t* nwords) {
  if (!h || !words || !nwords) return BCE_GPU_E_ARG;
  std::vector<uint16_t> out;
  try { out = h->w->finish(offset); }
  catch (...) { delete h->w; delete h; return BCE_GPU_E_NOMEM; }
  delete h->w;
  delete h;
  uint16_t* mem = static_cast<uint16_t*>(std::malloc(out.size() * sizeof(uint16_t) + 2));
  if (!mem) return BCE_GPU_E_NOMEM;
  std::memcpy(mem, out.data(), out.size() * sizeof(uint16_t));
  *words = mem;
  *nwords = out.size();
  return BCE_GPU_OK;
}
void bce_archive_abort(bce_archive_writer* h) {
  if (!h) return;
  delete h->w;
  delete h;
}

bce_scan* bce_scan_begin(void) {
  auto* h = new (std::nothrow) bce_scan{nullptr};
  if (!h) return nullptr;
  try { h->s = new ScanSession(); }
  catch (...) { h->s = nullptr; }
  if (!h->s) { delete h; return nullptr; }
  return h;
}
int bce_scan_feed(bce_scan* h, const bce_cse_batch* batch) {
  if (!h || !batch) return BCE_GPU_E_ARG;
  BCE_HOST_GUARD(h->s->feed(*batch);)
  return BCE_GPU_OK;
}
int bce_scan_finish(bce_scan* h, uint8_t cfg288_out[288]) {
  if (!h || !cfg288_out) return BCE_GPU_E_ARG;
  int rc = BCE_GPU_OK;
  try {
    ConfigTable t = h->s->finish();
    std::memcpy(cfg288_out, t.data(), 288);
  } catch (const std::bad_alloc&) { rc = BCE_GPU_E_NOMEM; }
  catch (...) { rc = BCE_GPU_E_INTERNAL; }
  delete h->s;
  delete h;
  return rc;
}

int bce_decode_buffer(const uint16_t* words, size_t nwords, int low_memory, uint8_t** out, size_t* nout) {
  if (!words || !out || !nout) return BCE_GPU_E_ARG;
  std::vector<uint8_t> text;
  int rc = BCE_GPU_OK;
  BCE_HOST_GUARD(
    std::vector<uint16_t> archive(words, words + nwords);
    rc = decode_archive(archive, low_memory != 0, text);                         // BCE::decode, bce.cpp:1169-1233
  )
  if (rc) return rc;
  uint8_t* mem = static_cast<uint8_t*>(std::malloc(text.size() + 1));
  if (!mem) return BCE_GPU_E_NOMEM;
  std::memcpy(mem, text.data(), text.size());
  *out = mem;
  *nout = text.size();
  return BCE_GPU_OK;
}

#ifndef BCE_HOST_NO_GPU
int bce_compress_buffer(bce_gpu_ctx* ctx, const uint8_t* T, uint32_t n, const uint8_t* cfg288, int threads,
                        uint16_t** words, size_t* nwords) {
  using clk = std::chrono::steady_clock;
  auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
  const bool timing = std::getenv("BCE_TIME") != nullptr;           // the reference's -DM_TIME lines, per phase
  const auto t0 = clk::now();
  uint32_t offset = 0, C[8];
  // the device emits coder-ready words: context index and k > 31 halving are done there
  int rc = bce_gpu_set_emit_mode(ctx, BCE_EMIT_CODER, cfg288);
  if (rc) return rc;
  rc = bce_gpu_compress_front(ctx, T, n, &offset, C);               // RankFile ctor, bce.cpp:1411
  if (rc) { bce_gpu_set_emit_mode(ctx, BCE_EMIT_RAW, nullptr); return rc; }
  const auto t1 = clk::now();
  bce_archive_writer* w = bce_archive_begin(n, C, cfg288);          // BCE::encode, bce.cpp:1417
  if (!w) { bce_gpu_set_emit_mode(ctx, BCE_EMIT_RAW, nullptr); return BCE_GPU_E_NOMEM; }
  // Double buffering over the context's two pinned batch buffers: while the coder threads work on batch j, the
  // next call copies batch j+1 out and runs the kernels of batch j+2 (a batch stays valid until the call after
  // the next one, include/bce_gpu.h).  threads <= 1 keeps the serial form.
  bce_cse_words20 cur, nxt;                             // 20 bits per word over PCIe
  double gpu_s = 0, wait_s = 0;
  size_t nbatch = 0, nw = 0;
  auto fail = [&](int code) { bce_archive_abort(w); bce_gpu_set_emit_mode(ctx, BCE_EMIT_RAW, nullptr); return code; };
  {
    const auto a = clk::now();
    rc = bce_gpu_cse_next_words20(ctx, &cur);
    gpu_s += secs(a, clk::now());
    if (rc) return fail(rc);
  }
  for (;;) {
    ++nbatch;
    for (int i = 0; i < 8; ++i) nw += cur.count[i];
    if (threads <= 1) {
      const auto a = clk::now();
      try { w->w->begin_words20(cur); w->w->wait_words(); } catch (...) { return fail(BCE_GPU_E_NOMEM); }
      wait_s += secs(a, clk::now());
      if (cur.done) break;
      const auto b = clk::now();
      rc = bce_gpu_cse_next_words20(ctx, &nxt);
      gpu_s += secs(b, clk::now());
      if (rc) return fail(rc);
    } else {
      try { w->w->begin_words20(cur); } catch (...) { return fail(BCE_GPU_E_NOMEM); }
      const bool last = cur.done != 0;
      if (!last) {
        const auto a = clk::now();
        rc = bce_gpu_cse_next_words20(ctx, &nxt);                  // coders of batch j run under this call
        gpu_s += secs(a, clk::now());
      }
      const auto b = clk::now();
      w->w->wait_words();
      wait_s += secs(b, clk::now());
      if (last) break;
      if (rc) return fail(rc);
    }
    cur = nxt;
  }
  bce_gpu_set_emit_mode(ctx, BCE_EMIT_RAW, nullptr);
  const auto t2 = clk::now();
  double busiest = 0;
  for (int i = 0; i < 8; ++i) busiest = std::max(busiest, w->w->busy_seconds(i));
  rc = bce_archive_finish(w, offset, words, nwords);
  if (timing)
    std::fprintf(stderr, "[bce] front (H2D + BWT + wavelet) %.3f s | in cse_next_words %.3f s (%s the coders) | waiting for "
                 "the coders after it %.3f s | busiest coder %.3f s (%d threads, %zu words in %zu batches) | finish %.3f s\n",
                 secs(t0, t1), gpu_s, threads > 1 ? "under" : "not overlapped with", wait_s, busiest, threads, nw, nbatch,
                 secs(t2, clk::now()));
  return rc;
}

int bce_scan_buffer(bce_gpu_ctx* ctx, const uint8_t* T, uint32_t n, uint8_t cfg288_out[288]) {
  uint32_t offset = 0, C[8];
  int rc = bce_gpu_set_emit_mode(ctx, BCE_EMIT_SCAN, nullptr);
  if (rc) return rc;
  rc = bce_gpu_compress_front(ctx, T, n, &offset, C);
  if (rc) { bce_gpu_set_emit_mode(ctx, BCE_EMIT_RAW, nullptr); return rc; }
  bce_scan* s = bce_scan_begin();
  if (!s) { bce_gpu_set_emit_mode(ctx, BCE_EMIT_RAW, nullptr); return BCE_GPU_E_NOMEM; }
  // the counts arrive bucketed by (k, context key) from the device; the host appends runs of symbol bytes and
  // runs ScanCoder's flush unchanged (bce.cpp:751-800)
  bce_scan_buckets batch;
  do {
    rc = bce_gpu_cse_next_buckets(ctx, &batch);
    if (!rc) rc = bce_scan_feed_buckets(s, &batch);
    if (rc) { delete s->s; delete s; bce_gpu_set_emit_mode(ctx, BCE_EMIT_RAW, nullptr); return rc; }
  } while (!batch.done);
  bce_gpu_set_emit_mode(ctx, BCE_EMIT_RAW, nullptr);
  return bce_scan_finish(s, cfg288_out);
}
#endif

}  // extern "C"
