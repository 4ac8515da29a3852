// Host-side marshalling helpers (defslam_b200/csrc/ds_host.h): the persistent thread pool behind host_parallel_for /
// host_big_memcpy -- every index covered exactly once, regions of all sizes, a second application thread that finds
// the pool busy falls back to its own thread, nothing hangs at process exit.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../defslam_b200/csrc/ds_host.h"

int main() {
  setenv("DEFSLAM_HOST_THREADS", "4", 1);
  int bad = 0;
  for (size_t n : {(size_t)0, (size_t)1, (size_t)7, (size_t)1000, (size_t)100003}) {
    std::vector<std::atomic<int>> hit(n ? n : 1);
    for (auto &h : hit) h = 0;
    ds::host_parallel_for(n, 10, [&](size_t lo, size_t hi) { for (size_t i = lo; i < hi; i++) hit[i]++; });
    for (size_t i = 0; i < n; i++) bad += hit[i] != 1;
  }
  std::vector<uint8_t> a(24 << 20), b(24 << 20, 0);
  for (size_t i = 0; i < a.size(); i++) a[i] = (uint8_t)(i * 2654435761u >> 24);
  ds::host_big_memcpy(b.data(), a.data(), a.size());
  bad += a != b;
  // two application threads inside the "library" at the same time
  std::atomic<long> total{0};
  auto work = [&] {
    for (int r = 0; r < 200; r++)
      ds::host_parallel_for(4000, 100, [&](size_t lo, size_t hi) { total += (long)(hi - lo); });
  };
  std::thread t1(work), t2(work);
  t1.join(); t2.join();
  bad += total != 2L * 200 * 4000;
  printf("host pool: %s (max threads %u)\n", bad ? "FAILED" : "ok", ds::HostPool::get().max_threads());
  return bad ? 1 : 0;
}
