// cnsm_kernels.cuh — constrained z-normalised matching (cNSM), shared by the ED and DTW engines.
// Replaces K/NormQueryEngine.java:432-528 and the statistics half of K/NormQueryEngineDtw.java:457-603.
//
// Why a "chain walker": the reference's window mean/std come from a sequential add-then-subtract
// chain (ex += d; ...; ex -= T[j]) that restarts at every merged interval.  Its rounding errors
// accumulate along the chain, so a prefix-sum or fresh-sum kernel cannot reproduce it to better than
// ~1e-7 relative on long chains — not enough for bit-exact alpha/beta gates or 1e-9 distances.  The
// chain is therefore emulated exactly: ONE THREAD PER CHAIN, the 32 lanes of a warp walking 32
// different chains in lock-step.  Samples reach the lanes through shared-memory tiles that are
// filled with coalesced 256-byte row loads (one row per chain), so HBM sees only sequential streams.
//
//   cnsm_walk_kernel   chain-exact ex/ex2 per window + a cheap conservative alpha/beta pre-gate;
//                      windows that may pass are appended (offset, ex, ex2) to the warp's private
//                      region of the work list (no global atomics in the streaming loop)
//   cnsm_plan_kernel   exclusive scan of per-region tile counts -> flat tile index for the evaluators
//   cnsm_ed_eval_kernel  one thread per work-list entry: exact mean/std/gate (reference arithmetic),
//                      then a fast FMA distance in |zQ|-descending order with early abandon against
//                      eps^2*(1+1e-9); survivors go to the exact list
//   cnsm_ed_exact_kernel one thread per survivor: the reference's sequential, unfused sum -> the
//                      accepted distances are bit-identical to the Java loop's
#pragma once
#include "common.cuh"

namespace kvm {

constexpr int kWalkWarps = 4;     // warps per CTA
constexpr int kWalkTile = 32;     // samples per lane per shared-memory tile
constexpr int kWalkPitch = 33;    // tile row pitch in doubles (conflict-free lane-private rows)
constexpr int kFifoDepth = 8;     // per-lane staging of work-list entries between flushes
constexpr int kWalkSmemDoublesPerWarp = 2 * 32 * kWalkPitch + 2 * kFifoDepth * 32 + (kFifoDepth * 32) / 2;
constexpr int kEvalTile = 128;    // work-list entries per evaluator tile (= evaluator CTA size)

struct WalkParams {
  const double* __restrict__ T;
  const int32_t* __restrict__ cbegin;   // per chain: local 0-based index of its first sample
  const int32_t* __restrict__ cnsamp;   // per chain: samples in the chain (0 = nothing to do)
  const long long* __restrict__ region_base;  // per walker warp: first work-list slot of its region
  int K;
  int m;
  int32_t first_global;
  // conservative pre-gate (superset of the exact gate; see DESIGN.md "cNSM pre-gate")
  double inv_m, meanQ, beta_hi, var_lo, var_hi;
  int32_t* e_off;
  double* e_ex;
  double* e_ex2;
  int32_t* region_count;
};

__global__ void __launch_bounds__(kWalkWarps * 32) cnsm_walk_kernel(WalkParams P) {
  extern __shared__ double walk_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* tA = walk_smem + (size_t)warp * kWalkSmemDoublesPerWarp;  // incoming samples  [32][pitch]
  double* tS = tA + 32 * kWalkPitch;                                // outgoing samples  [32][pitch]
  double* f_ex = tS + 32 * kWalkPitch;                              // [depth][32]
  double* f_ex2 = f_ex + kFifoDepth * 32;
  int32_t* f_off = reinterpret_cast<int32_t*>(f_ex2 + kFifoDepth * 32);

  const int region = blockIdx.x * kWalkWarps + warp;
  const int c = region * 32 + lane;
  int pos = 0, len = 0;
  if (c < P.K) {
    pos = P.cbegin[c];
    len = P.cnsamp[c];
  }
  const int maxlen = warp_max_i32(len);
  if (maxlen == 0) {
    if (lane == 0 && region * 32 < P.K) P.region_count[region] = 0;
    return;
  }
  const long long base = P.region_base[region];
  const int m = P.m;
  const double* __restrict__ T = P.T;
  int rcount = 0, fcnt = 0;
  double ex = 0.0, ex2 = 0.0;

  auto flush = [&]() {
    const int incl = warp_incl_scan_i32(fcnt, lane);
    const int total = __shfl_sync(kFullMask, incl, 31);
    long long e = base + rcount + (incl - fcnt);
    for (int i = 0; i < fcnt; i++) {
      P.e_off[e + i] = f_off[i * 32 + lane];
      P.e_ex[e + i] = f_ex[i * 32 + lane];
      P.e_ex2[e + i] = f_ex2[i * 32 + lane];
    }
    rcount += total;
    fcnt = 0;
  };

  for (int s0 = 0; s0 < maxlen; s0 += kWalkTile) {
    // Fill both tiles: row r = chain of lane r, 32 consecutive samples per row (coalesced).
#pragma unroll 8
    for (int r = 0; r < 32; r++) {
      const int p = __shfl_sync(kFullMask, pos, r);
      const int l = __shfl_sync(kFullMask, len, r);
      const int idx = s0 + lane;
      double a = 0.0, o = 0.0;
      if (idx < l) {
        a = T[p + idx];
        if (idx >= m - 1) o = T[p + idx - (m - 1)];
      }
      tA[r * kWalkPitch + lane] = a;
      tS[r * kWalkPitch + lane] = o;
    }
    __syncwarp();
#pragma unroll 4
    for (int i = 0; i < kWalkTile; i++) {
      const int s = s0 + i;
      if (s < len) {
        const double d = tA[lane * kWalkPitch + i];
        ex = xadd(ex, d);                 // K/NormQueryEngine.java:498
        ex2 = xadd(ex2, xmul(d, d));      // :499
        if (s >= m - 1) {
          // conservative pre-gate on (ex, ex2); FMA is fine here, the exact gate is re-done later
          const double mean_a = ex * P.inv_m;
          const double var_a = __fma_rn(ex2, P.inv_m, -(mean_a * mean_a));
          const bool pass = (fabs(mean_a - P.meanQ) <= P.beta_hi) && (var_a <= P.var_hi) && (var_a >= P.var_lo);
          if (pass) {
            f_ex[fcnt * 32 + lane] = ex;
            f_ex2[fcnt * 32 + lane] = ex2;
            f_off[fcnt * 32 + lane] = P.first_global + pos + s - (m - 1);
            fcnt++;
          }
          const double o = tS[lane * kWalkPitch + i];
          ex = xsub(ex, o);               // :523
          ex2 = xsub(ex2, xmul(o, o));    // :524
        }
      }
      if (__any_sync(kFullMask, fcnt == kFifoDepth)) flush();
    }
    __syncwarp();
  }
  flush();
  if (lane == 0) P.region_count[region] = rcount;
}

// Single-CTA exclusive scan over regions: tile_prefix[r] = sum_{r'<r} ceil(count[r']/kEvalTile).
// totals[0] = #tiles, totals[1] = #entries.
__global__ void __launch_bounds__(1024) cnsm_plan_kernel(const int32_t* __restrict__ region_count, int n_regions,
                                                         int32_t* __restrict__ tile_prefix,
                                                         unsigned long long* __restrict__ totals) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  __shared__ unsigned long long s_entries;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    s_carry = 0;
    s_entries = 0ULL;
  }
  __syncthreads();
  unsigned long long my_entries = 0;
  for (int r0 = 0; r0 < n_regions; r0 += 1024) {
    const int r = r0 + tid;
    const int cnt = (r < n_regions) ? region_count[r] : 0;
    my_entries += (unsigned long long)cnt;
    const int tiles = (cnt + kEvalTile - 1) / kEvalTile;
    const int incl = warp_incl_scan_i32(tiles, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = s_warp[lane];
      const int wi = warp_incl_scan_i32(w, lane);
      s_warp[lane] = wi - w;
    }
    __syncthreads();
    const int excl = s_carry + s_warp[warp] + incl - tiles;
    if (r < n_regions) tile_prefix[r] = excl;
    __syncthreads();
    if (tid == 1023) s_carry = excl + tiles;
    __syncthreads();
  }
  atomicAdd(&s_entries, my_entries);
  __syncthreads();
  if (tid == 0) {
    tile_prefix[n_regions] = s_carry;
    totals[0] = (unsigned long long)s_carry;
    totals[1] = s_entries;
  }
}

struct EvalParams {
  const double* __restrict__ T;
  int32_t first_global;
  int m;
  // work list
  const int32_t* __restrict__ e_off;
  const double* __restrict__ e_ex;
  const double* __restrict__ e_ex2;
  const long long* __restrict__ region_base;
  const int32_t* __restrict__ region_count;
  const int32_t* __restrict__ tile_prefix;
  const unsigned long long* __restrict__ totals;
  int n_regions;
  // query
  const double* __restrict__ zq;       // |z|-descending order (ED) / natural order (DTW)
  const int32_t* __restrict__ order;   // ED only
  double meanQ, stdQ, alpha, inv_alpha, beta;   // exact gate, K/NormQueryEngine.java:511
  double eps2, eps2_hi;                          // eps2_hi = eps2*(1+1e-9)+1e-18: fast-path guard band
  CandList out;
  unsigned long long* gate_pass;
};

// Exact window statistics and gate from the chain sums — the reference's arithmetic, unfused.
__device__ __forceinline__ bool cnsm_exact_gate(double ex, double ex2, int m, double meanQ, double stdQ, double alpha,
                                                double inv_alpha, double beta, double& mean, double& stdv) {
  const double dm = (double)m;
  mean = xdiv(ex, dm);                                          // :508
  stdv = xsqrt(xsub(xdiv(ex2, dm), xmul(mean, mean)));          // :509
  const double ratio = xdiv(stdv, stdQ);
  return (fabs(xsub(mean, meanQ)) <= beta) && (ratio <= alpha) && (ratio >= inv_alpha);  // :511
}

__global__ void __launch_bounds__(kEvalTile) cnsm_ed_eval_kernel(EvalParams P) {
  __shared__ int s_r;
  __shared__ unsigned int s_gate;
  if (threadIdx.x == 0) s_gate = 0;
  const int n_tiles = (int)P.totals[0];
  const int m = P.m;
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) s_r = find_segment<int32_t>(P.tile_prefix, P.n_regions + 1, t);
    __syncthreads();
    const int r = s_r;
    const int i = (t - P.tile_prefix[r]) * kEvalTile + (int)threadIdx.x;
    bool live = i < P.region_count[r];
    double mean = 0.0, stdv = 1.0;
    int32_t off = 0;
    if (live) {
      const long long e = P.region_base[r] + i;
      off = P.e_off[e];
      live = cnsm_exact_gate(P.e_ex[e], P.e_ex2[e], m, P.meanQ, P.stdQ, P.alpha, P.inv_alpha, P.beta, mean, stdv);
    }
    const unsigned gmask = __ballot_sync(kFullMask, live);
    if ((threadIdx.x & 31) == 0 && gmask) atomicAdd(&s_gate, (unsigned)__popc(gmask));
    if (!live) continue;
    // fast distance: x = T*rstd - mean*rstd, FMA allowed (approximate; the guard band absorbs it)
    const double rstd = 1.0 / stdv;
    const double nmr = -mean * rstd;
    const double* __restrict__ w = P.T + (off - P.first_global);
    double dist = 0.0;
    int k = 0;
    bool alive = true;
    for (; k + 4 <= m && alive; k += 4) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const double x = __fma_rn(w[__ldg(P.order + k + u)], rstd, nmr);
        const double df = x - __ldg(P.zq + k + u);
        dist = __fma_rn(df, df, dist);
      }
      alive = dist <= P.eps2_hi;
    }
    if (alive) {
      for (; k < m; k++) {
        const double x = __fma_rn(w[__ldg(P.order + k)], rstd, nmr);
        const double df = x - __ldg(P.zq + k);
        dist = __fma_rn(df, df, dist);
      }
    }
    if (dist <= P.eps2_hi) {
      const unsigned long long slot = atomicAdd(P.out.count, 1ULL);
      if ((long long)slot < P.out.cap) {
        P.out.off[slot] = off;
        P.out.mean[slot] = mean;
        P.out.stdv[slot] = stdv;
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_gate) atomicAdd(P.gate_pass, (unsigned long long)s_gate);
}

struct ExactEdParams {
  const double* __restrict__ T;
  int32_t first_global;
  int m;
  const double* __restrict__ zq;
  const int32_t* __restrict__ order;
  double eps2;
  CandList in;
  AnswerSink sink;
};

// K/NormQueryEngine.java:513-520 verbatim arithmetic: x = (T[order[k]+j]-mean)/std; dist += (x-zQ[k])^2.
__global__ void __launch_bounds__(128) cnsm_ed_exact_kernel(ExactEdParams P) {
  unsigned long long n = *P.in.count;
  if ((long long)n > P.in.cap) n = (unsigned long long)P.in.cap;
  const int m = P.m;
  for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < n;
       e += (unsigned long long)gridDim.x * blockDim.x) {
    const int32_t off = P.in.off[e];
    const double mean = P.in.mean[e], stdv = P.in.stdv[e];
    const double* __restrict__ w = P.T + (off - P.first_global);
    double dist = 0.0;
    bool alive = true;
    int k = 0;
    for (; k + 8 <= m && alive; k += 8) {
      double t[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const double x = xdiv(xsub(w[__ldg(P.order + k + u)], mean), stdv);
        t[u] = xsqdist(x, __ldg(P.zq + k + u));
      }
#pragma unroll
      for (int u = 0; u < 8; u++) dist = xadd(dist, t[u]);
      alive = dist <= P.eps2;
    }
    if (alive) {
      for (; k < m; k++) {
        const double x = xdiv(xsub(w[__ldg(P.order + k)], mean), stdv);
        dist = xadd(dist, xsqdist(x, __ldg(P.zq + k)));
      }
    }
    if (dist <= P.eps2) P.sink.emit(off, xsqrt(dist));
  }
}

}  // namespace kvm
