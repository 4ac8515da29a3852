// Which hardware warp slot (%warpid; scheduler / sub-partition = %warpid % 4) do the warps of two co-resident
// 256-thread CTAs get?  Decides where the latency-chain warp of sft_rows.h should sit.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256, 2) k(int *out) {
  extern __shared__ double sm[];
  unsigned smid, wid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  if ((threadIdx.x & 31) == 0) {
    out[(blockIdx.x * 8 + (threadIdx.x >> 5)) * 2] = smid;
    out[(blockIdx.x * 8 + (threadIdx.x >> 5)) * 2 + 1] = wid;
  }
  sm[threadIdx.x] = 1.0;
  long long t0 = clock64();
  while (clock64() - t0 < 2000000) {}
}
int main() {
  int *d; cudaMalloc(&d, 296 * 16 * 4);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 109952);
  k<<<296, 256, 109952>>>(d);
  cudaDeviceSynchronize();
  static int h[296 * 16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  for (int b : {0, 1, 2, 147, 148, 149, 150, 295}) {
    printf("block %3d sm %3d warpids:", b, h[b * 16]);
    for (int w = 0; w < 8; w++) printf(" %2d", h[(b * 8 + w) * 2 + 1]);
    printf("\n");
  }
  // per SM: list blocks
  for (int s = 0; s < 3; s++) { printf("sm %d blocks:", s); for (int b = 0; b < 296; b++) if (h[b * 16] == s) printf(" %d", b); printf("\n"); }
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
}
