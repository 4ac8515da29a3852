/*
 * ds_runtime.h -- host runtime shared by the .cu translation units: per-thread,
 * per-device stream + grow-only pinned/device buffers, launch accounting.
 */
#ifndef DS_RUNTIME_H_
#define DS_RUNTIME_H_

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <map>
#include <memory>

#include "../../include/defslam_b200.h"

namespace ds {

extern std::atomic<long long> g_launches;
extern thread_local double g_last_kernel_ms;

#define DS_CUDA_TRY(expr)                          \
  do {                                             \
    cudaError_t _e = (expr);                       \
    if (_e != cudaSuccess) {                       \
      ds::note_cuda_error(_e, #expr, __FILE__, __LINE__); \
      return DEFSLAM_ECUDA;                        \
    }                                              \
  } while (0)

void note_cuda_error(cudaError_t e, const char *what, const char *file, int line);

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  bool pinned = false;
  int ensure(size_t need) {
    if (need <= cap && p) return 0;
    if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); p = nullptr; cap = 0; }
    size_t want = need + need / 4 + 256;
    cudaError_t e = pinned ? cudaMallocHost(&p, want) : cudaMalloc(&p, want);
    if (e != cudaSuccess) { note_cuda_error(e, "alloc", __FILE__, __LINE__); p = nullptr; return DEFSLAM_ECUDA; }
    cap = want;
    return 0;
  }
  void release() {
    if (p) { if (pinned) cudaFreeHost(p); else cudaFree(p); }
    p = nullptr; cap = 0;
  }
};

struct DevCtx {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int sm_count = 0;
  int smem_optin = 0;
  ~DevCtx();
};

/* Entry points that take a device ordinal switch to it (get_ctx) for their CUDA calls; this guard puts the
 * caller's current device back when the entry point returns. */
struct DeviceGuard {
  int prev = -1;
  DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
  ~DeviceGuard() { if (prev >= 0) { int cur = -1; if (cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev); } }
};

/* cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute of the function, not of the calling thread
 * or stream: it is only ever RAISED, under a lock, so that two host threads launching batches of different sizes
 * cannot undercut each other between set and launch. */
cudaError_t raise_dynamic_smem(const void *func, int device, int bytes);

/* context of the calling thread on `device` (-1: current device); nullptr when
 * no usable CUDA device exists -- callers return DEFSLAM_ECUDA, there is no
 * CPU fallback. */
DevCtx *get_ctx(int device);

}  // namespace ds
#endif
