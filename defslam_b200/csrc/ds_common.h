/*
 * ds_common.h -- shared definitions of the device code.
 *
 * The solver kernels are written against a tiny "team" abstraction: every
 * parallel phase is a `DS_FOR(i, n)` loop strided over the threads of one CTA,
 * phases are separated by `team.sync()`.  Under nvcc this is ordinary CUDA
 * (threadIdx-strided loops + __syncthreads).  The same sources also compile
 * with plain g++ (DS_EMULATE) as a one-thread team; that build exists ONLY so
 * that tests can exercise the kernel source on a box without a GPU
 * (tests/emu/, never shipped, never loaded by the package).
 */
#ifndef DS_ISO
#define DS_ISO 0
#endif
#ifndef DS_FAST_RSQRT
#define DS_FAST_RSQRT 0
#endif
#ifndef DS_COMMON_H_
#define DS_COMMON_H_

#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__) && !defined(DS_EMULATE)
#define DS_CUDA 1
#define DS_FN __device__ __forceinline__
#define DS_MFN __device__ __forceinline__
#define DS_FN_NOINLINE __device__ __noinline__
#else
#define DS_CUDA 0
#define DS_FN static inline
#define DS_MFN inline
#define DS_FN_NOINLINE static
#endif

namespace ds {

constexpr int NB = 8;           /* panel width of the banded factorisation   */
constexpr int TILE = 8;         /* tile edge of the trailing update (one DMMA m8n8k4 output) */
constexpr int NFACC = 48;       /* per-facet accumulators (12 barycentric moments + 3 x 12 weighted Jacobian sums) */
constexpr int NMSCR = 18;       /* per-match scratch doubles                 */

/* 16-byte pair of doubles (LDS.128 / STS.128 on the device) */
struct alignas(16) dbl2 {
  double x, y;
};

struct Team {
  int tid, nthr;
  DS_MFN void sync() const {
#if DS_CUDA
    __syncthreads();
#endif
  }
  DS_MFN void warp_sync() const {
#if DS_CUDA
    __syncwarp();
#endif
  }
  /* lane count seen by warp-scoped loops (DS_WARP_FOR) */
  DS_MFN int lane() const {
#if DS_CUDA
    return tid & 31;
#else
    return 0;
#endif
  }
  DS_MFN bool warp0() const { return tid < 32; }
};

#if DS_CUDA
#define DS_FOR(i, n) for (int i = team.tid; i < (n); i += team.nthr)
#define DS_WARP_FOR(l, n) for (int l = team.lane(); l < (n); l += 32)
#else
#define DS_FOR(i, n) for (int i = 0; i < (n); i++)
#define DS_WARP_FOR(l, n) for (int l = 0; l < (n); l++)
#endif

/* Deterministic CTA-wide sum.  red: shared scratch of >= 33 doubles.  All
 * threads must call; all threads get the result. */
DS_FN double team_sum(const Team &team, double v, double *red) {
#if DS_CUDA
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = team.tid >> 5, nw = (team.nthr + 31) >> 5;
  team.sync(); /* red may still be read from a previous call */
  if ((team.tid & 31) == 0) red[w] = v;
  team.sync();
  if (team.tid < 32) {
    double s = team.tid < nw ? red[team.tid] : 0.0;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (team.tid == 0) red[32] = s;
  }
  team.sync();
  return red[32];
#else
  (void)team; (void)red;
  return v;
#endif
}

DS_FN double team_max(const Team &team, double v, double *red) {
#if DS_CUDA
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
  const int w = team.tid >> 5, nw = (team.nthr + 31) >> 5;
  team.sync();
  if ((team.tid & 31) == 0) red[w] = v;
  team.sync();
  if (team.tid < 32) {
    double s = team.tid < nw ? red[team.tid] : -DBL_MAX;
    for (int o = 16; o > 0; o >>= 1) s = fmax(s, __shfl_down_sync(0xffffffffu, s, o));
    if (team.tid == 0) red[32] = s;
  }
  team.sync();
  return red[32];
#else
  (void)team; (void)red;
  return v;
#endif
}

DS_FN int team_sum_int(const Team &team, int v, double *red) {
  return (int)(team_sum(team, (double)v, red) + 0.5);
}

DS_FN int atomic_inc_int(int *p) {
#if DS_CUDA
  return atomicAdd(p, 1);
#else
  return (*p)++;
#endif
}

#if DS_CUDA
__host__ __device__
#endif
static inline int round_up(int a, int m) { return ((a + m - 1) / m) * m; }

}  // namespace ds
#endif
