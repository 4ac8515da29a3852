// Micro-benchmark of the latency chain of the banded factorisation: Cholesky of one 8x8 diagonal block + inverse of
// its factor by ONE warp (every lane redundant), as the chain warp of sft_rows.h runs it.  Variants are compared for
// cycles per block; the other warps of the CTA optionally hammer the FP64 tensor cores (contention as in the solve).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/chainbench tools/chainbench.cu
#include <cstdio>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ double rsq_lib(double d) { return rsqrt(d); }
__device__ __forceinline__ double rsq_fast(double d) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  const double h = y0 * y0;
  const double e = fma(-d, h, 1.0);
  const double p = fma(0.375, e, 0.5);
  const double t = y0 * e;
  return fma(t, p, y0);
}
// second-order only (error ~ e^2 ~ 2^-40): for reference of what the third-order term costs
__device__ __forceinline__ double rsq_fast2(double d) {
  double y0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  const double h = y0 * y0;
  const double e = fma(-d, h, 1.0);
  return fma(y0 * 0.5, e, y0);
}

template <int RS> __device__ __forceinline__ double rsq(double d) {
  return RS == 0 ? rsq_lib(d) : (RS == 1 ? rsq_fast(d) : rsq_fast2(d));
}

// straightforward right-looking Cholesky of the packed lower triangle a[i*(i+1)/2+j], reciprocal diagonals
template <int RS> __device__ __forceinline__ void chol8(double *a) {
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const double inv = rsq<RS>(a[k * (k + 1) / 2 + k]);
    a[k * (k + 1) / 2 + k] = inv;
#pragma unroll
    for (int i = k + 1; i < 8; i++) a[i * (i + 1) / 2 + k] *= inv;
#pragma unroll
    for (int i = k + 1; i < 8; i++)
#pragma unroll
      for (int j = k + 1; j <= i; j++) a[i * (i + 1) / 2 + j] -= a[i * (i + 1) / 2 + k] * a[j * (j + 1) / 2 + k];
  }
}
// third-order reciprocal from the hardware seed (MUFU.RCP64H, ~2^-20): y1 = y0 (1 + e + e^2), e = 1 - d y0
__device__ __forceinline__ double rcp_fast(double d) {
  double y0;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
  const double e = fma(-d, y0, 1.0);
  const double t = y0 * e;
  const double q = 1.0 + e;
  return fma(t, q, y0);
}
// 2x2 block pivots (block LDL^T turned into the Cholesky factor off the chain): per pair of columns the only
// dependent sequence is det -> 1/det -> one fma per entry of the next pivot block; the reciprocal square roots that
// scale the columns of L are not on the path to the next pivot
__device__ __forceinline__ void chol8_blk2(double *a) {
#define A_(i, j) a[(i) * ((i) + 1) / 2 + (j)]
#pragma unroll
  for (int k = 0; k < 8; k += 2) {
    const double p = A_(k, k), b = A_(k + 1, k), c = A_(k + 1, k + 1);
    const double det = fma(p, c, -b * b);
    const double rdet = rcp_fast(det);
    double na[8], nb[8];
#pragma unroll
    for (int i = k + 2; i < 8; i++) {
      const double u = A_(i, k), v = A_(i, k + 1);
      na[i] = fma(u, c, -v * b);
      nb[i] = fma(v, p, -u * b);
    }
#pragma unroll
    for (int i = k + 2; i < 8; i++)
#pragma unroll
      for (int j = k + 2; j <= i; j++) {
        const double w = fma(na[i], A_(j, k), nb[i] * A_(j, k + 1));
        A_(i, j) = fma(-w, rdet, A_(i, j));
      }
    const double r1 = rsq_fast(p), r2 = rsq_fast(p * det);
    A_(k, k) = r1;
    A_(k + 1, k) = b * r1;
    A_(k + 1, k + 1) = p * r2;
#pragma unroll
    for (int i = k + 2; i < 8; i++) { A_(i, k) *= r1; A_(i, k + 1) = nb[i] * r2; }
  }
#undef A_
}
// column j of the inverse
__device__ __forceinline__ void invcol8(const double *a, int j, double *col) {
  double s[8];
#pragma unroll
  for (int i = 0; i < 8; i++) s[i] = i == j ? 1.0 : 0.0;
#pragma unroll
  for (int m = 0; m < 8; m++) {
    const double xm = s[m] * a[m * (m + 1) / 2 + m];
    col[m] = m >= j ? xm : 0.0;
#pragma unroll
    for (int i = m + 1; i < 8; i++) s[i] -= a[i * (i + 1) / 2 + m] * xm;
  }
}
// inverse by columns for a FIXED j (compile time): no wasted work above the diagonal
template <int J> __device__ __forceinline__ void invcol8_fixed(const double *a, double *col) {
  double s[8];
#pragma unroll
  for (int i = 0; i < 8; i++) s[i] = i == J ? 1.0 : 0.0;
#pragma unroll
  for (int m = 0; m < 8; m++) {
    if (m < J) { col[m] = 0.0; continue; }
    const double xm = s[m] * a[m * (m + 1) / 2 + m];
    col[m] = xm;
#pragma unroll
    for (int i = m + 1; i < 8; i++) s[i] -= a[i * (i + 1) / 2 + m] * xm;
  }
}

template <int RS, int INV, bool LOAD>
__global__ void chain(double *gm, long long *cyc, int n, int busy_warps) {
  __shared__ double D[2][64];
  __shared__ double Y[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 128) ((double *)D)[threadIdx.x] = gm[threadIdx.x & 63];
  __syncthreads();
  if (warp == 0) {
    double acc = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < n; it++) {
      double a[36];
      const volatile double *Dv = D[it & 1];
#pragma unroll
      for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j <= i; j++) a[i * (i + 1) / 2 + j] = LOAD ? Dv[i * 8 + j] : (gm[i * 8 + j] + acc);
      a[0] += acc; /* dependency between iterations: the chain */
      if (RS == 3) chol8_blk2(a); else chol8<RS>(a);
      double col[8];
      if (INV == 1) {
        invcol8(a, lane & 7, col);
      } else if (INV == 2) {
        switch (lane & 7) {
          case 0: invcol8_fixed<0>(a, col); break;
          case 1: invcol8_fixed<1>(a, col); break;
          case 2: invcol8_fixed<2>(a, col); break;
          case 3: invcol8_fixed<3>(a, col); break;
          case 4: invcol8_fixed<4>(a, col); break;
          case 5: invcol8_fixed<5>(a, col); break;
          case 6: invcol8_fixed<6>(a, col); break;
          default: invcol8_fixed<7>(a, col); break;
        }
      } else {
#pragma unroll
        for (int m = 0; m < 8; m++) col[m] = a[m * (m + 1) / 2 + (m > 0 ? m - 1 : 0)];
      }
      if (lane < 8) {
#pragma unroll
        for (int m = 0; m < 8; m++) Y[m * 8 + lane] = col[m];
      }
      __syncwarp();
      acc = ((volatile double *)Y)[63] * 1e-30; /* the next block depends on the last entry of the inverse */
    }
    long long t1 = clock64();
    if (lane == 0) { cyc[0] = t1 - t0; gm[64] = acc; }
  } else if (warp <= busy_warps) {
    double c0 = 0, c1 = 0, d0 = 0, d1 = 0, a = gm[0] + lane, b = gm[1];
    for (int it = 0; it < n * 40; it++) {
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
    }
    gm[65 + threadIdx.x] = c0 + c1 + d0 + d1;
  }
}

// two chains on ONE sub-partition (warps 0 and 4 of a 256-thread CTA), as the chain warps of two co-resident CTAs:
// does predicating the redundant lanes off (only ACT lanes factor the block) shorten the pair?
template <int ACT>
__global__ void chain_pair(double *gm, long long *cyc, int n, int second) {
  __shared__ double D[2][2][64];
  __shared__ double Y[2][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ((double *)D)[threadIdx.x] = gm[threadIdx.x & 63];
  __syncthreads();
  if (warp == 0 || (second && warp == 4)) {
    const int w = warp == 0 ? 0 : 1;
    double acc = 0.0;
    long long t0 = clock64();
    for (int it = 0; it < n; it++) {
      if (lane < ACT) {
        double a[36];
        const volatile double *Dv = D[w][it & 1];
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j <= i; j++) a[i * (i + 1) / 2 + j] = Dv[i * 8 + j];
        a[0] += acc;
        chol8<1>(a);
        double col[8];
        invcol8(a, lane & 7, col);
        if (lane < 8) {
#pragma unroll
          for (int m = 0; m < 8; m++) Y[w][m * 8 + lane] = col[m];
        }
      }
      __syncwarp();
      acc = ((volatile double *)Y[w])[63] * 1e-30;
    }
    long long t1 = clock64();
    if (lane == 0) { cyc[w] = t1 - t0; gm[64 + w] = acc; }
  }
}
template <int ACT> void run_pair(double *gm, long long *cyc) {
  const int n = 2000;
  for (int second = 0; second <= 1; second++) {
    chain_pair<ACT><<<1, 256>>>(gm, cyc, n, second);
    cudaDeviceSynchronize();
    chain_pair<ACT><<<1, 256>>>(gm, cyc, n, second);
    cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
    printf("chain with %2d active lanes, %s: %7.1f cycles / block\n", ACT, second ? "two chains on the sub-partition" : "alone", (double)h[0] / n);
  }
}

template <int RS, int INV, bool LOAD> void run(const char *name, double *gm, long long *cyc, int busy) {
  const int n = 2000;
  chain<RS, INV, LOAD><<<1, 256>>>(gm, cyc, n, busy);
  cudaDeviceSynchronize();
  chain<RS, INV, LOAD><<<1, 256>>>(gm, cyc, n, busy);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-46s busy warps %d: %7.1f cycles / block\n", name, busy, (double)h / n);
}

__global__ void cmp_kernel(const double *gm, double *out) {
  double a[36], b[36];
  for (int i = 0; i < 8; i++)
    for (int j = 0; j <= i; j++) a[i * (i + 1) / 2 + j] = b[i * (i + 1) / 2 + j] = gm[i * 8 + j];
  chol8<0>(a);
  chol8_blk2(b);
  double m = 0.0;
  for (int i = 0; i < 36; i++) m = fmax(m, fabs(a[i] - b[i]) / fabs(a[i]));
  out[0] = m;
}

int main() {
  double h[64];
  for (int i = 0; i < 8; i++)
    for (int j = 0; j < 8; j++) h[i * 8 + j] = (i == j ? 10.0 : 0.0) + 1.0 / (1 + i + j);
  double *gm; long long *cyc;
  cudaMalloc(&gm, 8 * 1024); cudaMalloc(&cyc, 64);
  cudaMemcpy(gm, h, sizeof(h), cudaMemcpyHostToDevice);
  for (int busy = 0; busy <= 7; busy += 7) {
    run<0, 0, true>("chol8 lib rsqrt, no inverse", gm, cyc, busy);
    run<1, 0, true>("chol8 fast rsqrt (3rd order), no inverse", gm, cyc, busy);
    run<2, 0, true>("chol8 fast rsqrt (2nd order), no inverse", gm, cyc, busy);
    run<0, 1, true>("chol8 lib rsqrt + inverse (runtime column)", gm, cyc, busy);
    run<1, 1, true>("chol8 fast rsqrt + inverse (runtime column)", gm, cyc, busy);
    run<1, 2, true>("chol8 fast rsqrt + inverse (switch column)", gm, cyc, busy);
    run<3, 0, true>("2x2 block pivots, no inverse", gm, cyc, busy);
    run<3, 1, true>("2x2 block pivots + inverse (runtime column)", gm, cyc, busy);
  }
  run_pair<32>(gm, cyc);
  run_pair<16>(gm, cyc);
  run_pair<8>(gm, cyc);
  {
    double *out; cudaMalloc(&out, 8);
    cmp_kernel<<<1, 1>>>(gm, out);
    double m; cudaMemcpy(&m, out, 8, cudaMemcpyDeviceToHost);
    printf("2x2 block pivots vs library-rsqrt Cholesky: max relative difference of the factor %.3g\n", m);
  }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
