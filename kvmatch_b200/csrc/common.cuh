// common.cuh — shared device helpers for libkvmatch_gpu.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kvm {

constexpr unsigned kFullMask = 0xffffffffu;
constexpr double kDtwInf = 1e20;  // K/utils/DtwUtils.java:24 — the reference's "infinity" is 1e20, not inf

// Every value that is reported, or that decides an answer, is computed with these: round-to-nearest
// binary64 operations that the compiler can never contract into an FMA (Java never fuses).
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double xsqrt(double a) { return __dsqrt_rn(a); }
// (a-b)^2 as the reference's dist(): K/utils/DtwUtils.java:38-40
__device__ __forceinline__ double xsqdist(double a, double b) {
  double d = __dsub_rn(a, b);
  return __dmul_rn(d, d);
}
// DtwUtils.min: (a < b) ? a : b
__device__ __forceinline__ double xmin(double a, double b) { return (a < b) ? a : b; }

__device__ __forceinline__ int warp_max_i32(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(kFullMask, v, o));
  return v;
}

__device__ __forceinline__ int warp_incl_scan_i32(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(kFullMask, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Largest p in [0, n) with prefix[p] <= x  (prefix is non-decreasing, prefix[0] == 0).
template <typename T>
__device__ __forceinline__ int find_segment(const T* __restrict__ prefix, int n, T x) {
  int lo = 0, hi = n;  // answer in [lo, hi)
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (prefix[mid] <= x) lo = mid; else hi = mid;
  }
  return lo;
}

// Sparse answer sink: append with one atomic per answer; entries beyond `cap` are counted, not stored
// (the host grows the buffers and re-runs the call).
struct AnswerSink {
  int32_t* off;
  double* dist;
  unsigned long long* count;
  long long cap;
  __device__ __forceinline__ void emit(int32_t o, double d) const {
    unsigned long long i = atomicAdd(count, 1ULL);
    if ((long long)i < cap) {
      off[i] = o;
      dist[i] = d;
    }
  }
};

// a <= b for doubles that are both >= +0 (or NaN, which then compares false for a and true for b): on the integer pipe
__device__ __forceinline__ bool le_nonneg(double a, double b) { return __double_as_longlong(a) <= __double_as_longlong(b); }
// min of two non-negative doubles on the integer pipe (DMNMX is as slow as DSETP)
__device__ __forceinline__ double min_nonneg(double a, double b) {
  const long long x = __double_as_longlong(a), y = __double_as_longlong(b);
  return __longlong_as_double(x < y ? x : y);
}

// Per-candidate record handed from the statistics / lower-bound stages to the exact stages.
struct CandList {
  int32_t* off;   // 1-based global window start
  double* mean;   // exact chain statistics (0 / 1 for the raw-series engines)
  double* stdv;
  unsigned long long* count;
  long long cap;
  double* lb = nullptr;  // optional: the candidate's LB_Keogh(EQ) total (0 when the producing stage did not form it)
};

}  // namespace kvm
