// LDS.128 issue cost for a single warp: lane-private rows (pitch 272 B) vs lane-consecutive, with 16 loads per iteration.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, long long* cyc, int iters) {
  __shared__ __align__(16) double tile[2 * 32 * 34];
  for (int i = threadIdx.x; i < 2 * 32 * 34; i += 32) tile[i] = i;
  __syncwarp();
  const int lane = threadIdx.x;
  const double* base = MODE == 0 ? tile + lane * 34 : tile + lane * 2;   // private row / consecutive
  const int cstride = MODE == 0 ? 2 : 64;                                // doubles between successive loads
  double acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    double2 v[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const double* p = base + (i % 8) * cstride + (i / 8) * (MODE == 0 ? 32 * 34 : 16 * 64);
      asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[i].x), "=d"(v[i].y) : "r"((unsigned)__cvta_generic_to_shared(p)));
    }
#pragma unroll
    for (int i = 0; i < 16; i++) acc += v[i].x + v[i].y;
  }
  long long t1 = clock64();
  out[lane] = acc;
  if (lane == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 8 * 64); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  k<0><<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("lane-private rows (pitch 272B): %.2f cycles per LDS.128 (incl. 2 DADD each) (%s)\n", (double)h / (iters * 16.0), cudaGetErrorString(cudaGetLastError()));
  k<1><<<1, 32>>>(out, cyc, iters); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("lane-consecutive: %.2f cycles per LDS.128 (incl. 2 DADD each)\n", (double)h / (iters * 16.0));
  return 0;
}
