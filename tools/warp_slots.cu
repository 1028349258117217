// Which SM sub-partition (%warpid & 3) do the warps of co-resident CTAs land on?  usage: warp_slots <threads> <smem_bytes>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void probe(int* out, int nw) {
  extern __shared__ char sm[];
  unsigned smid, wid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
  if ((threadIdx.x & 31) == 0) {
    int w = threadIdx.x >> 5;
    out[(blockIdx.x * nw + w) * 2] = smid;
    out[(blockIdx.x * nw + w) * 2 + 1] = wid;
  }
  long long t0 = clock64();
  while (clock64() - t0 < 2000000) {}
  if (sm[threadIdx.x] == 77) out[0] = 1;
}
int main(int argc, char** argv) {
  int threads = atoi(argv[1]), smem = atoi(argv[2]), nw = threads / 32, grid = atoi(argv[3]);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int* d; cudaMalloc(&d, grid * nw * 8);
  probe<<<grid, threads, smem>>>(d, nw);
  int* h = (int*)malloc(grid * nw * 8);
  cudaMemcpy(h, d, grid * nw * 8, cudaMemcpyDeviceToHost);
  printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
  for (int b = 0; b < grid; b++) if (h[b * nw * 2] < 2) {
    printf("cta %3d sm %d warpids:", b, h[b * nw * 2]);
    for (int w = 0; w < nw; w++) printf(" %d", h[(b * nw + w) * 2 + 1]);
    printf("\n");
  }
}
