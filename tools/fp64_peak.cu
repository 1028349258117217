// fp64_peak.cu — measures the FP64 pipe on this GPU: peak DFMA / DADD throughput (many independent chains,
// all SMs) and the dependent-issue latency of DADD (one chain, one warp).  The DTW roofline in bench.py
// divides by these numbers (MEASURED_PEAKS.json has no FP64 entry).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>  // 0: DFMA, 1: DADD, 2: DMNMX(min)
__global__ void throughput(double* out, int iters, double seed) {
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-7;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) a[i] = __fma_rn(a[i], b, c);
      else if (MODE == 1) a[i] = __dadd_rn(a[i], c);
      else a[i] = fmin(a[i], seed + it + i);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void latency(double* out, long long* cycles, int iters, double c) {
  double a = 1.0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) a = __dadd_rn(a, c);
  }
  long long t1 = clock64();
  out[threadIdx.x] = a;
  if (threadIdx.x == 0) *cycles = t1 - t0;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 8 * 256);
  cudaMalloc(&cyc, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000, grid = p.multiProcessorCount * 8, block = 256;
  const char* names[3] = {"DFMA", "DADD", "DMNMX"};
  double best[3] = {0, 0, 0};
  for (int mode = 0; mode < 3; mode++) {
    for (int rep = 0; rep < 5; rep++) {
      cudaEventRecord(e0);
      if (mode == 0) throughput<0><<<grid, block>>>(out, iters, 1.0);
      if (mode == 1) throughput<1><<<grid, block>>>(out, iters, 1.0);
      if (mode == 2) throughput<2><<<grid, block>>>(out, iters, 1.0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double ops = (double)grid * block * 8.0 * iters;
      const double rate = ops / (ms * 1e-3);
      if (rep > 0 && rate > best[mode]) best[mode] = rate;
    }
    printf("%s_gops %.1f\n", names[mode], best[mode] / 1e9);
  }
  latency<<<1, 32>>>(out, cyc, 4096, 1e-9);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("DADD_dependent_latency_cycles %.2f\n", (double)h / (4096.0 * 16));
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("sm_clock_khz_max %d sms %d\n", clk, p.multiProcessorCount);
  printf("{\"fp64_dfma_gops\": %.1f, \"fp64_dadd_gops\": %.1f, \"fp64_dmnmx_gops\": %.1f, \"fp64_tflops_fma\": %.2f}\n",
         best[0] / 1e9, best[1] / 1e9, best[2] / 1e9, 2 * best[0] / 1e12);
  return 0;
}
