// FP64 micro-benchmarks for B200: DFMA latency/throughput, DMMA (mma.sync m8n8k4 f64) throughput, rsqrt latency.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_latency(double *out, long long *cyc, int n) {
  double a = out[0], b = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) a = fma(a, b, c);
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
__global__ void dfma_tput(double *out, long long *cyc, int n) {
  double a[ILP];
  for (int k = 0; k < ILP; k++) a[k] = out[k] + threadIdx.x;
  double b = 1.0000001, c = 1e-9;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) a[k] = fma(a[k], b, c);
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < ILP; k++) s += a[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
__global__ void dmma_tput(double *out, long long *cyc, int n) {
  double c0[ILP], c1[ILP];
  for (int k = 0; k < ILP; k++) { c0[k] = 0; c1[k] = 0; }
  double a = out[0] + threadIdx.x * 1e-3, b = out[1] + 1e-3;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c0[k]), "+d"(c1[k]) : "d"(a), "d"(b));
  }
  __syncthreads();
  long long t1 = clock64();
  double s = 0;
  for (int k = 0; k < ILP; k++) s += c0[k] + c1[k];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void rsqrt_latency(double *out, long long *cyc, int n) {
  double a = out[0] + 2.0;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) a = rsqrt(a) + 1.5;
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void shfl_latency(double *out, long long *cyc, int n) {
  double a = out[0] + threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) a = __shfl_sync(0xffffffffu, a, (threadIdx.x + 1) & 31);
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void lds_latency(double *out, long long *cyc, int n) {
  __shared__ int idx[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) idx[i] = (i * 37 + 11) & 1023;
  __syncthreads();
  int j = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) j = idx[j];
  long long t1 = clock64();
  out[threadIdx.x] = j;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void sync_latency(double *out, long long *cyc, int n) {
  long long t0 = clock64();
  for (int i = 0; i < n; i++) __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; out[0] = 1; }
}

int main() {
  double *out; long long *cyc, h;
  cudaMalloc(&out, 8192 * 8); cudaMemset(out, 0, 8192 * 8);
  cudaMalloc(&cyc, 64);
  const int n = 4096;
  dfma_latency<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA dependent latency: %.2f cycles\n", (double)h / n);
  rsqrt_latency<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("rsqrt(double)+add dependent latency: %.2f cycles\n", (double)h / n);
  shfl_latency<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("shfl(double) dependent latency: %.2f cycles\n", (double)h / n);
  lds_latency<<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS dependent latency: %.2f cycles\n", (double)h / n);
  for (int thr : {64, 256, 1024}) {
    sync_latency<<<1, thr>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("__syncthreads (%d thr): %.2f cycles\n", thr, (double)h / n);
  }
  for (int thr : {32, 128, 256, 512, 1024}) {
    dfma_tput<8><<<1, thr>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA tput %4d thr ILP8: %.2f FMA/clk/SM\n", thr, (double)thr * 8 * n / h);
  }
  dfma_tput<16><<<1, 256>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DFMA tput  256 thr ILP16: %.2f FMA/clk/SM\n", 256.0 * 16 * n / h);
  for (int thr : {32, 128, 256, 512, 1024}) {
    dmma_tput<4><<<1, thr>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DMMA m8n8k4 %4d thr ILP4: %.2f FMA/clk/SM (%.1f cyc/inst/warp)\n", thr, (double)(thr / 32) * 4 * n * 256 / h,
           (double)h / (4.0 * n));
  }
  dmma_tput<1><<<1, 32>>>(out, cyc, n); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DMMA dependent latency: %.2f cycles\n", (double)h / n);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
