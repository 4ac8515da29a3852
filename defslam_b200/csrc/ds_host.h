/*
 * ds_host.h -- host-only helpers shared by the marshalling code (pure C++, no CUDA: the emulation build uses it too).
 */
#ifndef DS_HOST_H_
#define DS_HOST_H_

#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace ds {

/* Host-side marshalling runs on a few threads when there is enough of it: the arenas are tens of MB per call
 * (124 MB per 2368-frame SfT batch, 70 MB per 480 k-point normals batch) and one core copies ~10 GB/s.
 * fn(begin, end) over [0, n), at least `grain` items per thread, at most 8 threads (DEFSLAM_HOST_THREADS). */
/* A few persistent helper threads (created on first use, detached: never joined at process teardown).  One
 * parallel region at a time; a caller that finds the pool busy (another host thread of the application is inside
 * the library) runs its region on its own thread. */
class HostPool {
 public:
  static HostPool &get() { static HostPool *p = new HostPool; return *p; }
  unsigned max_threads() const { return nmax_; }
  /* fn(t) for t in [0, nt), t = 0 on the calling thread; returns false (nothing done) when the pool is busy */
  template <class Fn>
  bool run(unsigned nt, Fn &fn) {
    if (nt <= 1) { fn(0u); return true; }
    std::unique_lock<std::mutex> region(region_, std::try_to_lock);
    if (!region.owns_lock()) return false;
    ensure_workers(nt - 1);
    {
      std::lock_guard<std::mutex> g(m_);
      call_ = [](void *f, unsigned t) { (*(Fn *)f)(t); };
      arg_ = &fn; want_ = nt - 1; next_ = 1; pending_ = nt - 1; epoch_++;
    }
    cv_.notify_all();
    fn(0u);
    std::unique_lock<std::mutex> g(m_);
    done_.wait(g, [&] { return pending_ == 0; });
    return true;
  }

 private:
  HostPool() {
    unsigned t = std::thread::hardware_concurrency() / 2;
    /* one process per GPU on a shared host (torchrun exports LOCAL_WORLD_SIZE): the ranks share the cores */
    if (const char *e = getenv("LOCAL_WORLD_SIZE")) { const int w = atoi(e); if (w > 1) t /= (unsigned)w; }
    if (const char *e = getenv("DEFSLAM_HOST_THREADS")) t = (unsigned)atoi(e);
    nmax_ = std::max(1u, std::min(8u, t));
  }
  void ensure_workers(unsigned n) {
    while (nworkers_ < n) {
      std::thread([this] { worker(); }).detach();
      nworkers_++;
    }
  }
  void worker() {
    unsigned long long seen = 0;
    for (;;) {
      unsigned t;
      void (*call)(void *, unsigned);
      void *arg;
      {
        std::unique_lock<std::mutex> g(m_);
        cv_.wait(g, [&] { return epoch_ != seen && next_ <= want_; });
        t = next_++;
        if (next_ > want_) seen = epoch_;
        call = call_; arg = arg_;
      }
      call(arg, t);
      {
        std::lock_guard<std::mutex> g(m_);
        if (--pending_ == 0) done_.notify_all();
      }
    }
  }
  std::mutex region_, m_;
  std::condition_variable cv_, done_;
  void (*call_)(void *, unsigned) = nullptr;
  void *arg_ = nullptr;
  unsigned want_ = 0, next_ = 1, pending_ = 0, nworkers_ = 0, nmax_ = 1;
  unsigned long long epoch_ = 0;
};

template <class Fn>
static inline void host_parallel_for(size_t n, size_t grain, Fn fn) {
  HostPool &pool = HostPool::get();
  size_t nt = grain ? n / grain : 1;
  if (nt > pool.max_threads()) nt = pool.max_threads();
  if (nt <= 1) { fn((size_t)0, n); return; }
  const size_t per = (n + nt - 1) / nt;
  auto part = [&](unsigned t) {
    const size_t b = std::min(n, (size_t)t * per), e = std::min(n, b + per);
    if (e > b) fn(b, e);
  };
  if (!pool.run((unsigned)nt, part)) fn((size_t)0, n);
}
/* memcpy split over threads above 4 MB */
static inline void host_big_memcpy(void *dst, const void *src, size_t bytes) {
  host_parallel_for(bytes, (size_t)4 << 20, [=](size_t b, size_t e) { memcpy((uint8_t *)dst + b, (const uint8_t *)src + b, e - b); });
}

}  // namespace ds
#endif
