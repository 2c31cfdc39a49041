// Host-side label packer of the upload path (plain C++: no device code, compiled by the host compiler).
//
// Replaces the label upload of the training loop and of the consensus notebook on the host side:
//   b_seg_modified = b_seg_modified.to(device=config.device)      main_deep_staple.py:687   (int64, 8 B per voxel)
//   the per-candidate `.to_dense()` / `.cuda()` of consensus.ipynb:382-392
// Label maps are int64 in the reference's containers but hold class indices < 256: eight bytes per voxel cross
// PCIe where one carries the information.  ds_host_narrow_labels packs them to uint8 on the host (worker threads,
// each streaming a contiguous slice at memory speed) into a pinned staging buffer; the kernels read DS_U8 labels
// natively, so the packed bytes are what is copied.  Lossless: any label outside [0, 255] is reported and the
// caller copies the int64 labels instead.
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <immintrin.h>
#include <stdint.h>
#include <pthread.h>

#include "deep_staple_b200.h"

namespace ds {
int fail(int code, const char* fmt, ...);  // ds_api.cu

namespace {

// persistent workers: one job at a time (callers are serialised by `submit_mu`), slices handed out by index
class HostPool {
 public:
  static HostPool& get() {
    // never destroyed: worker threads outlive static destruction order.  A forked child (a DataLoader worker that
    // packs labels itself) inherits the object but none of its threads: the atfork handler makes the child start a
    // pool of its own; the inherited one is left alone (its mutexes may have been held at the time of the fork).
    static std::once_flag once;
    std::call_once(once, [] { pthread_atfork(nullptr, nullptr, [] { instance().store(nullptr); }); });
    HostPool* pool = instance().load(std::memory_order_acquire);
    if (!pool) {
      HostPool* fresh = new HostPool();
      if (instance().compare_exchange_strong(pool, fresh)) pool = fresh;
      else delete fresh;  // another caller was first; `fresh` has no threads yet
    }
    return *pool;
  }
  // runs fn(i) for i in [0, n) on up to `threads` threads (the caller takes part), returns when all are done
  void run(int n, int threads, const std::function<void(int)>& fn) {
    if (n <= 0) return;
    std::lock_guard<std::mutex> submit(submit_mu_);
    const int helpers = (threads < n ? threads : n) - 1;
    grow(helpers);
    {
      std::lock_guard<std::mutex> lk(mu_);
      fn_ = &fn, n_ = n, next_.store(0), pending_ = helpers, allowed_ = helpers, ++generation_;
    }
    cv_.notify_all();
    work();
    std::unique_lock<std::mutex> lk(mu_);
    done_cv_.wait(lk, [&] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  static std::atomic<HostPool*>& instance() {
    static std::atomic<HostPool*> p{nullptr};
    return p;
  }
  void work() {
    for (int i; (i = next_.fetch_add(1)) < n_;) (*fn_)(i);
  }
  void grow(int helpers) {
    while ((int)workers_.size() < helpers) {
      const int id = (int)workers_.size();
      workers_.emplace_back([this, id] {
        uint64_t seen = 0;
        for (;;) {
          std::unique_lock<std::mutex> lk(mu_);
          cv_.wait(lk, [&] { return generation_ != seen; });
          seen = generation_;
          if (id >= allowed_) continue;  // this job asked for fewer threads
          lk.unlock();
          work();
          lk.lock();
          if (--pending_ == 0) done_cv_.notify_one();
        }
      });
      workers_.back().detach();
    }
  }
  std::mutex submit_mu_, mu_;
  std::condition_variable cv_, done_cv_;
  std::vector<std::thread> workers_;
  const std::function<void(int)>* fn_ = nullptr;
  std::atomic<int> next_{0};
  int n_ = 0, pending_ = 0, allowed_ = 0;
  uint64_t generation_ = 0;
};

// one slice; returns the OR of every bit above the low byte of the labels it has seen
uint64_t narrow_scalar(const int64_t* __restrict__ src, uint8_t* __restrict__ dst, int64_t n) {
  uint64_t high = 0;
  for (int64_t i = 0; i < n; ++i) {
    const uint64_t v = (uint64_t)src[i];
    high |= v, dst[i] = (uint8_t)v;
  }
  return high & ~0xffull;
}

// AVX-512: vpmovqb narrows eight labels per instruction; 32 labels (256 B) per iteration
__attribute__((target("avx512f"))) uint64_t narrow_avx512(const int64_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                          int64_t n) {
  __m512i acc = _mm512_setzero_si512();
  int64_t i = 0;
  for (; i + 32 <= n; i += 32) {
    const __m512i a = _mm512_loadu_si512(src + i), b = _mm512_loadu_si512(src + i + 8);
    const __m512i c = _mm512_loadu_si512(src + i + 16), d = _mm512_loadu_si512(src + i + 24);
    acc = _mm512_or_si512(acc, _mm512_or_si512(_mm512_or_si512(a, b), _mm512_or_si512(c, d)));
    _mm_storel_epi64(reinterpret_cast<__m128i*>(dst + i), _mm512_cvtepi64_epi8(a));
    _mm_storel_epi64(reinterpret_cast<__m128i*>(dst + i + 8), _mm512_cvtepi64_epi8(b));
    _mm_storel_epi64(reinterpret_cast<__m128i*>(dst + i + 16), _mm512_cvtepi64_epi8(c));
    _mm_storel_epi64(reinterpret_cast<__m128i*>(dst + i + 24), _mm512_cvtepi64_epi8(d));
  }
  const uint64_t high = (uint64_t)_mm512_reduce_or_epi64(acc) & ~0xffull;
  return high | narrow_scalar(src + i, dst + i, n - i);
}

// AVX2: even dwords of two registers (shuffle_ps), dword -> word -> byte packs, one 16-bit interleave to undo the
// per-lane order; 16 labels per iteration.  The packs saturate, which only matters for labels that are reported anyway.
__attribute__((target("avx2"))) uint64_t narrow_avx2(const int64_t* __restrict__ src, uint8_t* __restrict__ dst,
                                                     int64_t n) {
  __m256i acc = _mm256_setzero_si256();
  int64_t i = 0;
  for (; i + 16 <= n; i += 16) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 4));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 8));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 12));
    acc = _mm256_or_si256(acc, _mm256_or_si256(_mm256_or_si256(a, b), _mm256_or_si256(c, d)));
    const __m256i ab = _mm256_castps_si256(_mm256_shuffle_ps(_mm256_castsi256_ps(a), _mm256_castsi256_ps(b), 0x88));
    const __m256i cd = _mm256_castps_si256(_mm256_shuffle_ps(_mm256_castsi256_ps(c), _mm256_castsi256_ps(d), 0x88));
    const __m256i w = _mm256_packus_epi32(ab, cd);  // lanes: a0 a1 b0 b1 c0 c1 d0 d1 | a2 a3 b2 b3 c2 c3 d2 d3
    const __m256i p = _mm256_packus_epi16(w, w);
    const __m128i r = _mm_unpacklo_epi16(_mm256_castsi256_si128(p), _mm256_extracti128_si256(p, 1));
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i), r);
  }
  alignas(32) uint64_t lanes[4];
  _mm256_store_si256(reinterpret_cast<__m256i*>(lanes), acc);
  const uint64_t high = (lanes[0] | lanes[1] | lanes[2] | lanes[3]) & ~0xffull;
  return high | narrow_scalar(src + i, dst + i, n - i);
}

using NarrowFn = uint64_t (*)(const int64_t*, uint8_t*, int64_t);
// DS_HOST_ISA=scalar|avx2|avx512 (read once): lets the tests run every variant the CPU supports
NarrowFn pick_narrow() {
  __builtin_cpu_init();
  if (__builtin_cpu_supports("avx512f")) return narrow_avx512;
  if (__builtin_cpu_supports("avx2")) return narrow_avx2;
  return narrow_scalar;
}

NarrowFn env_isa_override(NarrowFn best) {
  const char* e = getenv("DS_HOST_ISA");
  if (!e || !*e) return best;
  if (!strcmp(e, "scalar")) return narrow_scalar;
  if (!strcmp(e, "avx2") && __builtin_cpu_supports("avx2")) return narrow_avx2;
  if (!strcmp(e, "avx512") && __builtin_cpu_supports("avx512f")) return narrow_avx512;
  return best;
}

}  // namespace
}  // namespace ds

extern "C" int ds_host_narrow_labels(const int64_t* src, uint8_t* dst, int64_t n, int threads) {
  using namespace ds;
  if (!src || !dst) return fail(DS_ERR_NULL, "ds_host_narrow_labels: src and dst are required");
  if (n < 0) return fail(DS_ERR_SHAPE, "ds_host_narrow_labels: n must be non-negative");
  if (n == 0) return 0;
  if (threads < 1) {
    threads = (int)std::thread::hardware_concurrency();
    if (threads < 1) threads = 1;
  }
  if (threads > 64) threads = 64;
  constexpr int64_t kSlice = 1 << 18;  // 2 MB of int64 per slice: a few slices per thread, balanced tails
  const int64_t slices64 = (n + kSlice - 1) / kSlice;
  const int slices = (int)(slices64 < (1 << 30) ? slices64 : (1 << 30));
  const int64_t per = ((n + slices - 1) / slices + 7) & ~(int64_t)7;
  static const NarrowFn narrow_slice = env_isa_override(pick_narrow());
  std::atomic<uint64_t> high{0};
  const std::function<void(int)> job = [&](int s) {
    const int64_t a = (int64_t)s * per, b = a + per < n ? a + per : n;
    if (a < b) {
      const uint64_t h = narrow_slice(src + a, dst + a, b - a);
      if (h) high.fetch_or(h, std::memory_order_relaxed);
    }
  };
  if (threads == 1 || slices == 1) {
    for (int s = 0; s < slices; ++s) job(s);
  } else {
    HostPool::get().run(slices, threads, job);
  }
  if (high.load())
    return fail(DS_ERR_ARG, "ds_host_narrow_labels: a label lies outside [0, 255]; upload the int64 labels instead");
  return 0;
}
