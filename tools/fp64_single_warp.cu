// Single-warp FP64 issue behaviour on B200: cycles per DP instruction for ILP = 1,2,4,8 independent DADD chains,
// and for the walker's pattern (2 dependent chains of add/sub + 2 squares).
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, long long* cyc, int iters, double c) {
  double a[ILP];
#pragma unroll
  for (int i = 0; i < ILP; i++) a[i] = 1.0 + i + threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int i = 0; i < ILP; i++) a[i] = __dadd_rn(a[i], c);
    }
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < ILP; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
__global__ void walker_like(double* out, long long* cyc, int iters, const double* in) {
  double ex = 0, ex2 = 0;
  double a = in[threadIdx.x], o = in[threadIdx.x + 32];
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 16; r++) {
      ex = __dadd_rn(ex, a);
      ex2 = __dadd_rn(ex2, __dmul_rn(a, a));
      ex = __dsub_rn(ex, o);
      ex2 = __dsub_rn(ex2, __dmul_rn(o, o));
      a += 1.0;  // keep the squares from being hoisted (1 extra DADD, independent)
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = ex + ex2;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double *out, *in; long long* cyc; long long h;
  cudaMalloc(&out, 8 * 1024 * 148); cudaMalloc(&in, 8 * 64); cudaMalloc(&cyc, 8);
  cudaMemset(in, 0, 8 * 64);
  const int iters = 20000;
#define RUN(ILP, W) k<ILP><<<1, 32 * W>>>(out, cyc, iters, 1e-9); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("DADD ILP=%d warps=%d: %.2f cycles per warp-instruction\n", ILP, W, (double)h / (iters * 8.0 * ILP));
  RUN(1, 1) RUN(2, 1) RUN(4, 1) RUN(8, 1) RUN(8, 4) RUN(8, 8) RUN(8, 16)
  for (int w = 1; w <= 8; w *= 2) {
    walker_like<<<1, 32 * w>>>(out, cyc, iters, in); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("walker-like step (7 DP), %d warp(s)/SM: %.2f cycles per step\n", w, (double)h / (iters * 16.0));
  }
  return 0;
}
