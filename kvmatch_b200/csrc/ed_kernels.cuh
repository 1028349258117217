// ed_kernels.cuh — RSM-ED phase-2 verification (replaces K/QueryEngine.java:341-363).
//
// One thread per candidate window start; the lanes of a warp hold 32 consecutive starts, so every
// load T[start + j] is a coalesced 256-byte row and the query value q[j] is a warp-uniform
// broadcast.  Each thread accumulates sum (d-q)^2 in the reference's natural order with unfused
// binary64 ops, so an accepted distance is bit-identical to the Java loop's.  The early-abandon test
// `dist <= eps^2` is evaluated after 1, 4, 12, 20, ... terms instead of after every term: partial sums are monotone
// non-decreasing, so this changes neither the accept decision nor any accepted value.
//
// HBM traffic: the series is streamed once (8 B per verified subsequence on a full scan); the
// re-reads by neighbouring lanes hit L1.
#pragma once
#include "common.cuh"

namespace kvm {

constexpr int kEdThreads = 256;
constexpr int kEdAhead = 64;
constexpr int kEdTile = 4096;  // candidates per CTA (16 per thread): amortises the tile -> interval search

struct EdParams {
  const double* __restrict__ T;          // local shard, element 0 = global sample `first_global`
  const double* __restrict__ q;          // raw query, length m
  const int32_t* __restrict__ cbegin;    // per interval: local 0-based index of its first sample
  const int32_t* __restrict__ ncand;     // per interval: number of window starts
  const int32_t* __restrict__ tile_prefix;  // K+1 exclusive prefix of ceil(ncand/kEdTile)
  int K;
  int m;
  double eps2;
  int32_t first_global;  // 1-based global offset of T[0]
  int ahead;             // samples between a long-lived window's current round and the line it prefetches (kEdAhead; >= m: off)
  AnswerSink sink;
};

__global__ void __launch_bounds__(kEdThreads) ed_verify_kernel(EdParams P) {
  __shared__ int s_p;
  if (threadIdx.x == 0) s_p = find_segment<int32_t>(P.tile_prefix, P.K + 1, (int32_t)blockIdx.x);
  __syncthreads();
  const int p = s_p;
  const int c0 = ((int)blockIdx.x - P.tile_prefix[p]) * kEdTile;
  const int ncand = P.ncand[p];
  const int cbegin = P.cbegin[p];
  const double* __restrict__ q = P.q;
  const int m = P.m;
  const double eps2 = P.eps2;
  const double q0 = __ldg(q);
  const int ahead = P.ahead;
  // On a scan almost every window is hopeless after its first sample, so a thread's work per window is one 8-byte load
  // and a compare: the kernel is bound by the bytes in flight.  The first samples of four of the thread's windows are
  // therefore requested together (4 x 256 B per warp in flight instead of one load and a dependent branch).
  const int c_end = min(c0 + kEdTile, ncand);
  auto finish = [&](int start, double first) {
    const double* __restrict__ w = P.T + start;
    double dist = first;
    bool alive = true;
    int j = 1;
    if (m >= 4) {
#pragma unroll
      for (int u = 1; u < 4; u++) dist = xadd(dist, xsqdist(w[u], __ldg(q + u)));
      alive = dist <= eps2;
      j = 4;
    }
    for (; j + 8 <= m && alive; j += 8) {
      // A window that lives on (a match, or on smooth data its neighbours) is a chain of dependent rounds, each waiting
      // for lines that left L1 long ago: ask for the line eight rounds ahead.
      if (j + ahead < m) asm volatile("prefetch.global.L1 [%0];" ::"l"(w + j + ahead));
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; u++) t[u] = xsqdist(w[j + u], __ldg(q + j + u));
#pragma unroll
      for (int u = 0; u < 8; u++) dist = xadd(dist, t[u]);
      alive = dist <= eps2;
    }
    if (alive) {
      for (; j < m; j++) dist = xadd(dist, xsqdist(w[j], __ldg(q + j)));
    }
    if (dist <= eps2) P.sink.emit(P.first_global + start, xsqrt(dist));
  };
  for (int c = c0 + (int)threadIdx.x; c < c_end; c += 4 * kEdThreads) {
    double f[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int cu = c + u * kEdThreads;
      f[u] = (cu < c_end) ? P.T[cbegin + cu] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const double d0 = xsqdist(f[u], q0);
      if (c + u * kEdThreads < c_end && d0 <= eps2) finish(cbegin + c + u * kEdThreads, d0);
    }
  }
}

}  // namespace kvm
