/*
 * ds_host.h -- host-only helpers shared by the marshalling code (pure C++, no CUDA: the emulation build uses it too).
 */
#ifndef DS_HOST_H_
#define DS_HOST_H_

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <thread>
#include <vector>

namespace ds {

/* Host-side marshalling runs on a few threads when there is enough of it: the arenas are tens of MB per call
 * (124 MB per 2368-frame SfT batch, 70 MB per 480 k-point normals batch) and one core copies ~10 GB/s.
 * fn(begin, end) over [0, n), at least `grain` items per thread, at most 8 threads (DEFSLAM_HOST_THREADS). */
template <class Fn>
static inline void host_parallel_for(size_t n, size_t grain, Fn fn) {
  static const unsigned max_threads = [] {
    unsigned t = std::thread::hardware_concurrency() / 2;
    if (const char *e = getenv("DEFSLAM_HOST_THREADS")) t = (unsigned)atoi(e);
    return std::max(1u, std::min(8u, t));
  }();
  size_t nt = grain ? n / grain : 1;
  if (nt > max_threads) nt = max_threads;
  if (nt <= 1) { fn((size_t)0, n); return; }
  std::vector<std::thread> th;
  th.reserve(nt - 1);
  const size_t per = (n + nt - 1) / nt;
  for (size_t t = 1; t < nt; t++) {
    const size_t b = std::min(n, t * per), e = std::min(n, b + per);
    th.emplace_back([=] { fn(b, e); });
  }
  fn((size_t)0, std::min(n, per));
  for (auto &t : th) t.join();
}
/* memcpy split over threads above 4 MB */
static inline void host_big_memcpy(void *dst, const void *src, size_t bytes) {
  host_parallel_for(bytes, (size_t)4 << 20, [=](size_t b, size_t e) { memcpy((uint8_t *)dst + b, (const uint8_t *)src + b, e - b); });
}

}  // namespace ds
#endif
