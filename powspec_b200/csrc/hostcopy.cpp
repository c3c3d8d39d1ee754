// Host-side staging copy: pageable memory -> pinned ring, by a persistent pool of threads.
//
// The reference's C host hands genr_mesh() plain malloc'd DATA arrays
// (src/read_cata.c:86-189), which the DMA engine cannot read directly: they are staged
// through a ring of pinned pieces (context.cu: h2d_async).  Round 1 spawned 16 std::threads
// per 64 MB piece and used memcpy: 44 GB/s on the 16-core B200 host, below the 55 GB/s of
// the PCIe link, so the upload of a pageable catalogue was bound by the staging (108.5 ms per
// config-2 step against 90.5 ms from pinned memory).  Here the threads live as long as the
// context (no creation / join per piece: 55.9 GB/s with memcpy, 72.7 GB/s with streaming
// stores, tools/staging_bench.py) and the ring is sized for the last-level cache, which is
// what counts once the DMA engine competes for host DRAM: 94.5 ms per step.

#include <emmintrin.h>

#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace psb_host {

namespace {

// 1: non-temporal stores (option "h2d_nt").  Measured on the B200 host (config 2 from malloc'd
// memory, profiles/r2_v10_pageable_ring.jsonl): with a staging ring small enough to stay in the
// last-level cache (3 x 16 MB) plain stores win — the DMA engine reads the ring from the cache
// and host DRAM only sees the read of the source: 94.5 ms per step against 98.7-100.1 ms with
// streaming stores into 2 x 64 MB (three trips through DRAM per byte), pinned source 90.6 ms
int g_stream_stores = 0;

void stream_copy(char *dst, const char *src, size_t n) {
  if (n < 256 || !g_stream_stores) { memcpy(dst, src, n); return; }
  const size_t head = (64 - (reinterpret_cast<uintptr_t>(dst) & 63)) & 63;
  memcpy(dst, src, head);
  dst += head; src += head; n -= head;
  const size_t lines = n / 64;
  for (size_t i = 0; i < lines; i++, src += 64, dst += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 32));
    const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i *>(src + 48));
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst), a);
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i *>(dst + 48), d);
  }
  _mm_sfence();
  memcpy(dst, src, n - lines * 64);
}

}  // namespace

struct CopyPool {
  int n = 1;                            // participants: n - 1 workers + the caller
  int requested = 1;                    // what the creator asked for (n is smaller if threads ran out)
  std::vector<std::thread> workers;
  std::mutex m;
  std::condition_variable wake, done;
  unsigned long generation = 0;
  int pending = 0;
  bool stop = false;
  // the job of the current generation
  char *dst = nullptr;
  const char *src = nullptr;
  size_t bytes = 0, per = 0;

  void part(int t) const {
    const size_t a = per * (size_t) t < bytes ? per * (size_t) t : bytes;
    const size_t b = a + per < bytes ? a + per : bytes;
    if (b > a) stream_copy(dst + a, src + a, b - a);
  }
  void work(int t) {
    unsigned long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m);
        wake.wait(lk, [&] { return stop || generation != seen; });
        if (stop) return;
        seen = generation;
      }
      part(t);
      std::lock_guard<std::mutex> lk(m);
      if (--pending == 0) done.notify_one();
    }
  }
};

CopyPool *copy_pool_create(int nthreads) {
  CopyPool *p = new CopyPool();
  p->n = p->requested = nthreads < 1 ? 1 : nthreads;
  try {
    for (int t = 1; t < p->n; t++) p->workers.emplace_back([p, t] { p->work(t); });
  } catch (...) {       // out of threads: keep the ones that started
    p->n = (int) p->workers.size() + 1;
  }
  return p;
}

void copy_pool_destroy(CopyPool *p) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(p->m);
    p->stop = true;
  }
  p->wake.notify_all();
  for (auto &t : p->workers) t.join();
  delete p;
}

int copy_pool_threads(const CopyPool *p) { return p ? p->requested : 0; }
void copy_set_stream_stores(int on) { g_stream_stores = on; }
int g_stream_stores_get() { return g_stream_stores; }

// dst[0, bytes) = src[0, bytes); returns when the copy is complete and globally visible.
// One caller at a time per pool.
void copy_pool_run(CopyPool *p, void *dst, const void *src, size_t bytes) {
  if (!bytes) return;
  if (!p || p->n == 1 || bytes < ((size_t) 1 << 20)) {
    stream_copy(static_cast<char *>(dst), static_cast<const char *>(src), bytes);
    return;
  }
  {
    std::lock_guard<std::mutex> lk(p->m);
    p->dst = static_cast<char *>(dst);
    p->src = static_cast<const char *>(src);
    p->bytes = bytes;
    p->per = (bytes / (size_t) p->n + 4095) & ~(size_t) 4095;
    p->pending = p->n - 1;
    p->generation++;
  }
  p->wake.notify_all();
  p->part(0);
  std::unique_lock<std::mutex> lk(p->m);
  p->done.wait(lk, [&] { return p->pending == 0; });
}

}  // namespace psb_host

// test hook (tests/test_library_cpu.py): the staging copy needs no GPU
extern "C" int psb_test_host_copy(void *dst, const void *src, size_t bytes, int nthreads, int repeats) {
  psb_host::CopyPool *p = psb_host::copy_pool_create(nthreads);
  const int n = psb_host::copy_pool_threads(p);
  for (int r = 0; r < (repeats < 1 ? 1 : repeats); r++) psb_host::copy_pool_run(p, dst, src, bytes);
  psb_host::copy_pool_destroy(p);
  return n;
}

// test / measurement hook: stream `bytes` of src through two alternating staging buffers of
// `piece` bytes with ONE pool of `nthreads` — h2d_async (context.cu) without the DMA; returns
// the seconds it took (< 0: allocation failure).  stream_stores: 1 / 0 as option "h2d_nt"
extern "C" double psb_test_host_stage(const void *src, size_t bytes, size_t piece, int nthreads, int stream_stores) {
  if (!piece) return -1.0;
  char *stage[2] = {nullptr, nullptr};
  for (int i = 0; i < 2; i++) {
    if (posix_memalign(reinterpret_cast<void **>(&stage[i]), 4096, piece)) { free(stage[0]); return -1.0; }
    memset(stage[i], 0, piece);
  }
  const int saved = psb_host::g_stream_stores_get();
  psb_host::copy_set_stream_stores(stream_stores);
  psb_host::CopyPool *p = psb_host::copy_pool_create(nthreads);
  const auto t0 = std::chrono::steady_clock::now();
  int slot = 0;
  for (size_t off = 0; off < bytes; off += piece, slot ^= 1)
    psb_host::copy_pool_run(p, stage[slot], static_cast<const char *>(src) + off, bytes - off < piece ? bytes - off : piece);
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  psb_host::copy_pool_destroy(p);
  psb_host::copy_set_stream_stores(saved);
  free(stage[0]); free(stage[1]);
  return s;
}
