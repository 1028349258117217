// Drives the real cnsm_walk_kernel on a small resident series with clock-based timing (developer experiment harness).
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../kvmatch_b200/csrc/cnsm_kernels.cuh"
using namespace kvm;
int main(int argc, char** argv) {
  const int m = 1024, nchains = argc > 1 ? atoi(argv[1]) : 32, len = argc > 2 ? atoi(argv[2]) : 100000;
  const long long n = (long long)nchains * (len - m + 1) + m + 4096;
  std::vector<double> h(n + kFrontPad + kTailPad, 0.0);
  for (long long i = 0; i < n; i++) h[kFrontPad + i] = 100.0 + 3.0 * ((i * 2654435761u) % 1000) / 1000.0;
  double* d; cudaMalloc(&d, h.size() * 8); cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  std::vector<int32_t> cb(nchains), cn(nchains); std::vector<long long> rb((nchains + 31) / 32 + 1);
  for (int c = 0; c < nchains; c++) { cb[c] = c * (len - m + 1); cn[c] = len; }
  for (size_t r = 0; r < rb.size(); r++) rb[r] = (long long)r * 32 * (len - m + 1);
  int32_t *dcb, *dcn, *eoff, *rc; long long* drb; double *eex, *eex2;
  cudaMalloc(&dcb, 4 * nchains); cudaMalloc(&dcn, 4 * nchains); cudaMalloc(&drb, 8 * rb.size()); cudaMalloc(&rc, 4 * rb.size());
  const long long V = (long long)nchains * (len - m + 1);
  cudaMalloc(&eoff, 4 * V); cudaMalloc(&eex, 8 * V); cudaMalloc(&eex2, 8 * V);
  cudaMemcpy(dcb, cb.data(), 4 * nchains, cudaMemcpyHostToDevice); cudaMemcpy(dcn, cn.data(), 4 * nchains, cudaMemcpyHostToDevice);
  cudaMemcpy(drb, rb.data(), 8 * rb.size(), cudaMemcpyHostToDevice);
  WalkParams W{}; W.T = d + kFrontPad; W.cbegin = dcb; W.cnsamp = dcn; W.region_base = drb; W.K = nchains; W.m = m; W.first_global = 1;
  W.idx_hi = (int)((n + kTailPad - 2) & ~1LL); W.mean_klo = 0x7ffffff0; W.var_klo = 0; W.mean_kspan = 0; W.var_kspan = 0; W.dm = m;
  W.e_off = eoff; W.e_ex = eex; W.e_ex2 = eex2; W.region_count = rc;
  cudaFuncSetAttribute(cnsm_walk_kernel<4, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)walk_smem_bytes(4));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int rep = 0; rep < 3; rep++) {
    cudaEventRecord(e0);
    cnsm_walk_kernel<4, 1, 0><<<(nchains + 31) / 32, kWalkThreads, walk_smem_bytes(4)>>>(W);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%s: %d chains x %d steps: %.3f ms = %.1f ns/step = %.1f cycles/step @1.965GHz (%s)\n", EXPNAME, nchains, len, ms, ms * 1e6 / len, ms * 1e6 / len * 1.965, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
