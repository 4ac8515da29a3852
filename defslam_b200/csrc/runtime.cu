/*
 * runtime.cu -- host runtime + library information entry points.
 */
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <utility>

#include "ds_runtime.h"

namespace ds {

std::atomic<long long> g_launches{0};
thread_local double g_last_kernel_ms = 0.0;

void note_cuda_error(cudaError_t e, const char *what, const char *file, int line) {
  if (getenv("DEFSLAM_QUIET") == nullptr)
    fprintf(stderr, "defslam_b200: CUDA error %d (%s) at %s:%d: %s\n", (int)e, cudaGetErrorString(e), file, line, what);
  cudaGetLastError(); /* clear the sticky-free error state */
}

cudaError_t raise_dynamic_smem(const void *func, int device, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void *, int>, int> cur;
  std::lock_guard<std::mutex> lock(mu);
  int &have = cur[std::make_pair(func, device)];
  if (bytes <= have) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess) have = bytes;
  return e;
}

DevCtx::~DevCtx() {
  /* streams/events are released with the process; destroying them from a
   * thread_local destructor can run after the CUDA runtime shut down. */
}

DevCtx *get_ctx(int device) {
  static thread_local std::map<int, std::unique_ptr<DevCtx>> tl;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return nullptr; }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  }
  if (device >= ndev) return nullptr;
  auto it = tl.find(device);
  if (it != tl.end()) {
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return it->second.get();
  }
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  std::unique_ptr<DevCtx> c(new DevCtx);
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (cudaEventCreate(&c->e0) != cudaSuccess || cudaEventCreate(&c->e1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  DevCtx *r = c.get();
  tl[device] = std::move(c);
  return r;
}

}  // namespace ds

extern "C" {

const char *defslam_version(void) { return "defslam_b200 0.1.0 (sm_100a)"; }

int64_t defslam_kernel_launch_count(void) { return (int64_t)ds::g_launches.load(); }

int defslam_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

double defslam_last_kernel_ms(void) { return ds::g_last_kernel_ms; }

}
