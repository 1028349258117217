// dtw_kernels.cuh — Sakoe-Chiba banded DTW verification with the LB_Kim / LB_Keogh cascade.
// Replaces K/QueryEngineDtw.java:349-452, K/NormQueryEngineDtw.java:457-603 and
// K/utils/DtwUtils.java:149-337.
//
// Stage 1 (one thread per candidate, lanes = consecutive window starts -> coalesced):
//   LB_KimFL (3 points at each end) and LB_Keogh against the query envelope, in fast FMA arithmetic,
//   pruning only when the bound exceeds eps^2*(1+1e-9).  Both are valid lower bounds of the band DTW
//   (m >= 6), so pruning never changes the answer set; survivors are appended to a candidate list.
// Stage 2 (one warp per survivor): the DTW matrix is swept along anti-diagonals.  Cell (i,j) lives on
//   diagonal d = i+j at band coordinate u = i-j+rho in [0, 2 rho]; on one diagonal only every second u
//   is populated, so a lane owns R consecutive (even,odd) u-pairs in registers and needs one 64-bit
//   shuffle per step for the pair that straddles its neighbour.  Every cell is
//   min(min(x,y),z) + (a-b)^2 with unfused binary64 ops on the same operands as the reference's
//   row sweep, so the result is bit-identical to DtwUtils.dtw() whenever that does not abandon
//   (out-of-band / out-of-matrix neighbours are the reference's INF = 1e20).
#pragma once
#include "cnsm_kernels.cuh"
#include "ed_kernels.cuh"

namespace kvm {

struct LbQuery {
  const double* __restrict__ q;   // query in natural order (raw for RSM, z-normalised for cNSM)
  const double* __restrict__ uq;  // upper envelope of q, radius rho
  const double* __restrict__ lq;  // lower envelope
  int m;
  double eps2_hi;
};

__device__ __forceinline__ double fsq(double a, double b) {
  const double d = a - b;
  return d * d;
}

// true = the candidate may still be an answer
// x = (w - mean) * rstd (well conditioned for any |mean| / std; exact for the raw engines' mean = 0, rstd = 1)
__device__ __forceinline__ bool lb_cascade(const double* __restrict__ w, const LbQuery& Q, double rstd, double mean) {
  const int m = Q.m;
  const double* __restrict__ q = Q.q;
  double lb = 0.0;
  if (m >= 6) {  // LB_KimFL, K/utils/DtwUtils.java:149-189, with the reference's early returns: on a raw-series scan the two
                 // end points alone reject almost every window (one load per end, 5 FP64 operations); the minima run on the
                 // integer pipe (DMNMX issues at a fifth of the DADD rate)
    const double q0 = __ldg(q), p0 = __ldg(q + m - 1);
    const double x0 = ((w[0] - mean) * rstd), y0 = ((w[m - 1] - mean) * rstd);
    lb = fsq(x0, q0) + fsq(y0, p0);
    if (!le_nonneg(lb, Q.eps2_hi)) return false;
    auto mn = [](double a, double b) { return min_nonneg(a, b); };
    const double x1 = ((w[1] - mean) * rstd), y1 = ((w[m - 2] - mean) * rstd);
    const double q1 = __ldg(q + 1), p1 = __ldg(q + m - 2);
    lb += mn(mn(fsq(x1, q0), fsq(x0, q1)), fsq(x1, q1));
    lb += mn(mn(fsq(y1, p0), fsq(y0, p1)), fsq(y1, p1));
    if (!le_nonneg(lb, Q.eps2_hi)) return false;
    const double x2 = ((w[2] - mean) * rstd), y2 = ((w[m - 3] - mean) * rstd);
    const double q2 = __ldg(q + 2), p2 = __ldg(q + m - 3);
    lb += mn(mn(mn(fsq(x0, q2), fsq(x1, q2)), fsq(x2, q2)), mn(fsq(x2, q1), fsq(x2, q0)));
    lb += mn(mn(mn(fsq(y0, p2), fsq(y1, p2)), fsq(y2, p2)), mn(fsq(y2, p1), fsq(y2, p0)));
    if (!le_nonneg(lb, Q.eps2_hi)) return false;
  }
  // LB_Keogh on the query envelope, K/utils/DtwUtils.java:206-222
  lb = 0.0;
  bool alive = true;
  int k = 0;
  for (; k + 4 <= m && alive; k += 4) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const double x = ((w[k + u] - mean) * rstd);
      const double up = __ldg(Q.uq + k + u), lo = __ldg(Q.lq + k + u);
      const double d = (x > up) ? (x - up) : ((x < lo) ? (x - lo) : 0.0);
      lb = __fma_rn(d, d, lb);
    }
    alive = lb <= Q.eps2_hi;
  }
  if (alive) {
    for (; k < m; k++) {
      const double x = ((w[k] - mean) * rstd);
      const double up = __ldg(Q.uq + k), lo = __ldg(Q.lq + k);
      const double d = (x > up) ? (x - up) : ((x < lo) ? (x - lo) : 0.0);
      lb = __fma_rn(d, d, lb);
    }
  }
  return lb <= Q.eps2_hi;
}

__device__ __forceinline__ void cand_append(const CandList& L, int32_t off, double mean, double stdv, double lb = 0.0) {
  const unsigned long long slot = atomicAdd(L.count, 1ULL);
  if ((long long)slot < L.cap) {
    L.off[slot] = off;
    L.mean[slot] = mean;
    L.stdv[slot] = stdv;
    if (L.lb) L.lb[slot] = lb;
  }
}

// RSM-DTW stage 1: candidates enumerated from the interval tiles (same tiling as ed_verify_kernel).
struct LbRawParams {
  const double* __restrict__ T;
  const int32_t* __restrict__ cbegin;
  const int32_t* __restrict__ ncand;
  const int32_t* __restrict__ tile_prefix;
  int K;
  int32_t first_global;
  LbQuery Q;
  CandList out;
};

__global__ void __launch_bounds__(kEdThreads) dtw_lb_raw_kernel(LbRawParams P) {
  __shared__ int s_p;
  if (threadIdx.x == 0) s_p = find_segment<int32_t>(P.tile_prefix, P.K + 1, (int32_t)blockIdx.x);
  __syncthreads();
  const int p = s_p;
  const int c0 = ((int)blockIdx.x - P.tile_prefix[p]) * kEdTile;
  const int ncand = P.ncand[p], cbegin = P.cbegin[p];
  // Almost every window of a raw-series scan fails on its two end points: those are requested for four of the thread's
  // windows together (eight loads in flight per thread), only the windows that pass go through the cascade.
  const int c_end = min(c0 + kEdTile, ncand);
  const int m = P.Q.m;
  const double q0 = __ldg(P.Q.q), p0 = __ldg(P.Q.q + m - 1);
  for (int c = c0 + (int)threadIdx.x; c < c_end; c += 4 * kEdThreads) {
    double a[4], b[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int cu = c + u * kEdThreads;
      a[u] = (cu < c_end) ? P.T[cbegin + cu] : 0.0;
      b[u] = (cu < c_end) ? P.T[cbegin + cu + m - 1] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int cu = c + u * kEdThreads;
      if (cu >= c_end) continue;
      if (m >= 6 && !le_nonneg(fsq(a[u], q0) + fsq(b[u], p0), P.Q.eps2_hi)) continue;
      const int start = cbegin + cu;
      if (lb_cascade(P.T + start, P.Q, 1.0, 0.0)) cand_append(P.out, P.first_global + start, 0.0, 1.0);
    }
  }
}

// cNSM-DTW stage 1: work-list entries from cnsm_walk_kernel -> exact gate -> lower bounds.
struct LbNormParams {
  EvalParams E;  // work list + exact gate parameters (zq/order/eps fields unused)
  LbQuery Q;
};

__global__ void __launch_bounds__(kEvalTile) cnsm_dtw_lb_kernel(LbNormParams P) {
  __shared__ int s_r;
  __shared__ unsigned int s_gate;
  if (threadIdx.x == 0) s_gate = 0;
  const EvalParams& E = P.E;
  const int n_tiles = (int)E.totals[0];
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) s_r = find_segment<int32_t>(E.tile_prefix, E.n_regions + 1, t);
    __syncthreads();
    const int r = s_r;
    const int i = (t - E.tile_prefix[r]) * kEvalTile + (int)threadIdx.x;
    bool live = i < E.region_count[r];
    double mean = 0.0, stdv = 1.0;
    int32_t off = 0;
    if (live) {
      const long long e = E.region_base[r] + i;
      off = E.e_off[e];
      live = cnsm_exact_gate(E.e_ex[e], E.e_ex2[e], E.m, E.meanQ, E.stdQ, E.alpha, E.inv_alpha, E.beta, mean, stdv);
    }
    const unsigned gmask = __ballot_sync(kFullMask, live);
    if ((threadIdx.x & 31) == 0 && gmask) atomicAdd(&s_gate, (unsigned)__popc(gmask));
    if (!live) continue;
    const double rstd = 1.0 / stdv;
    if (lb_cascade(E.T + (off - E.first_global), P.Q, rstd, mean)) cand_append(E.out, off, mean, stdv);
  }
  __syncthreads();
  if (threadIdx.x == 0 && s_gate) atomicAdd(E.gate_pass, (unsigned long long)s_gate);
}

// cNSM-DTW stage 1 behind the streaming statistics pass (stream_kernels.cuh): entries are the flagged windows with
// their exact chain sums -> exact gate (counted) -> lower bounds -> candidate list.
struct LbListParams {
  const double* __restrict__ T;
  int32_t first_global;
  int m;
  XList in;
  double meanQ, stdQ, alpha, inv_alpha, beta;
  LbQuery Q;
  CandList out;
  unsigned long long* gate_pass;
};

__global__ void __launch_bounds__(128) cnsm_dtw_lb_list_kernel(LbListParams P) {
  unsigned long long n = *P.in.count;
  if ((long long)n > P.in.cap) n = (unsigned long long)P.in.cap;
  unsigned my_gate = 0;
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (unsigned long long)gridDim.x * blockDim.x) {
    double mean, stdv;
    const int32_t off = P.in.off[e];
    if (!cnsm_exact_gate(P.in.ex[e], P.in.ex2[e], P.m, P.meanQ, P.stdQ, P.alpha, P.inv_alpha, P.beta, mean, stdv)) continue;
    my_gate++;
    if (lb_cascade(P.T + (off - P.first_global), P.Q, 1.0 / stdv, mean)) cand_append(P.out, off, mean, stdv);
  }
  if (my_gate) atomicAdd(P.gate_pass, (unsigned long long)my_gate);
}

// ---------------------------------------------------------------------------------------------------------------
// LB_Keogh on the DATA envelope (K/utils/DtwUtils.java:238-257, lbKeoghDataCumulative; fed by
// lowerUpperLemire(data, ...) at K/QueryEngineDtw.java:397-399 / K/NormQueryEngineDtw.java:522-524): the third bound of
// the reference's cascade.  One warp per survivor of the first stage: the (normalised) window goes to shared memory,
// its sliding max / min of radius rho are formed with the van Herk / Gil-Werman block decomposition (block = 2 rho + 1:
// a prefix maximum from each block start and a suffix maximum to each block end, then
// U[i] = max(suffix[i - rho], prefix[i + rho])), and sum_i dist(q_i, [L_i, U_i])^2 is compared with eps^2.
// The envelope is taken inside the window (the reference's runs over its whole read buffer, a superset: ours is the
// tighter valid bound) and kept in single precision rounded outwards (a slightly wider, still valid envelope).
struct Lb2Params {
  const double* __restrict__ T;
  int32_t first_global;
  int m, rho;
  const double* __restrict__ q;  // query in natural order (raw / z-normalised)
  double eps2_hi;
  CandList in, out;
};

constexpr size_t lb2_warp_bytes(int m) { return sizeof(double) * (size_t)m + 2 * sizeof(float) * (size_t)m; }

__global__ void __launch_bounds__(256) dtw_lb_data_kernel(Lb2Params P) {
  extern __shared__ __align__(16) unsigned char lb2_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int m = P.m, rho = P.rho;
  double* A = reinterpret_cast<double*>(lb2_smem + (size_t)warp * lb2_warp_bytes(m));
  float* Pf = reinterpret_cast<float*>(A + m);  // prefix extremum from the block start
  float* Sf = Pf + m;                            // suffix extremum to the block end
  unsigned long long n = *P.in.count;
  if ((long long)n > P.in.cap) n = (unsigned long long)P.in.cap;
  const int blk = 2 * rho + 1;
  const int n_blk = (m + blk - 1) / blk;
  for (unsigned long long e = (unsigned long long)blockIdx.x * n_warps + warp; e < n;
       e += (unsigned long long)gridDim.x * n_warps) {
    const int32_t off = P.in.off[e];
    const double mean = P.in.mean[e], stdv = P.in.stdv[e];
    const double rstd = 1.0 / stdv;
    const double* __restrict__ w = P.T + (off - P.first_global);
    __syncwarp();
    for (int k = lane; k < m; k += 32) A[k] = (w[k] - mean) * rstd;
    __syncwarp();
    double lb = 0.0;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++) {  // pass 0: upper envelope (max), pass 1: lower envelope (min)
      // blocks are independent: lane = (block, direction); forward prefix / backward suffix scans
      for (int t = lane; t < 2 * n_blk; t += 32) {
        const int b = t >> 1, lo = b * blk, hi = min(m, lo + blk) - 1;
        if ((t & 1) == 0) {
          float run = pass == 0 ? -INFINITY : INFINITY;
          for (int k = lo; k <= hi; k++) {
            const float v = pass == 0 ? __double2float_ru(A[k]) : __double2float_rd(A[k]);
            run = pass == 0 ? fmaxf(run, v) : fminf(run, v);
            Pf[k] = run;
          }
        } else {
          float run = pass == 0 ? -INFINITY : INFINITY;
          for (int k = hi; k >= lo; k--) {
            const float v = pass == 0 ? __double2float_ru(A[k]) : __double2float_rd(A[k]);
            run = pass == 0 ? fmaxf(run, v) : fminf(run, v);
            Sf[k] = run;
          }
        }
      }
      __syncwarp();
      for (int i = lane; i < m; i += 32) {
        const int lo = max(0, i - rho), hi = min(m - 1, i + rho);
        const double qv = __ldg(P.q + i);
        if (pass == 0) {
          const double up = (double)fmaxf(Sf[lo], Pf[hi]);
          const double d = qv > up ? qv - up : 0.0;
          lb = __fma_rn(d, d, lb);
        } else {
          const double dn = (double)fminf(Sf[lo], Pf[hi]);
          const double d = qv < dn ? dn - qv : 0.0;
          lb = __fma_rn(d, d, lb);
        }
      }
      __syncwarp();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) lb += __shfl_xor_sync(kFullMask, lb, o);
    if (lane == 0 && lb <= P.eps2_hi) cand_append(P.out, off, mean, stdv);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Both LB_Keogh bounds as ONE streaming kernel with survivor compaction.
//   LB_Keogh(EQ): candidate against the query envelope   K/utils/DtwUtils.java:206-222
//   LB_Keogh(EC): query against the DATA envelope        K/utils/DtwUtils.java:238-257
// The data envelope does not depend on the window: l[i] / u[i] = min / max of the raw samples within rho of i
// (the reference forms it over each read buffer, K/QueryEngineDtw.java:397-399 / K/NormQueryEngineDtw.java:522-524, and
// normalises it per window, (l - mean) / std).  It is therefore built ONCE per (series, rho) by envelope_kernel over the
// whole resident shard (16 bytes per sample, cached in the ctx) and every candidate reads it like it reads the series.
// Taken over the shard rather than one read buffer it is never narrower than the reference's: still a valid bound.
//
// Thread per candidate, CTA = 256 candidates that walk the m terms together in chunks of kLbChunk: lanes hold
// neighbouring window starts, so the series / envelope loads of a warp are coalesced and neighbouring lanes re-use each
// other's lines out of L1, and the query terms of the chunk are broadcast from shared memory.  After every chunk the
// candidates whose bound exceeds eps^2 (1 + 1e-9) are dropped and the survivors are COMPACTED to the front of the CTA,
// so warps stay full of live candidates (a warp of the thread-per-candidate kernels this replaces ran as long as its
// longest-lived lane — with 3 % survivors almost every warp ran all m terms — and the data-envelope bound rebuilt a
// van Herk envelope per candidate: together 50 ns per candidate; this kernel: see DESIGN.md).
// Arithmetic: x = (w - mean) * (1 / std) (well conditioned for any |mean| / std), excess outside [lo, up] as
// max(|x - centre| - half, 0) with the clamp on the sign bit (no DSETP / DMNMX, which issue at a fifth of the DADD rate).
constexpr int kLbChunk = 64;
constexpr int kLbThreads = 256;
constexpr int kLbWarpPathMax = 32768;  // lists up to this length take the warp-per-candidate path

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

struct LbFusedParams {
  const double* __restrict__ T;
  const double* __restrict__ envL;  // data envelope, indexed like T
  const double* __restrict__ envU;
  int32_t first_global;
  int m;
  XList xin;    // kFromSums: (offset, ex, ex2) -> exact gate here
  CandList cin; // else: (offset, mean, std)
  double meanQ, stdQ, alpha, inv_alpha, beta;
  const double* __restrict__ q;   // natural order (raw / z-normalised)
  const double* __restrict__ uq;  // query envelope
  const double* __restrict__ lq;
  double eps2_hi;
  CandList out;
  unsigned long long* gate_pass;
  int warp_path_max;  // lists up to this length take the warp-per-candidate path (kLbWarpPathMax; a developer knob lowers it)
};

// Block-wide stable compaction of the threads with `alive`: returns the number of live threads, `idx` = this thread's
// new position.  Two barriers.
__device__ __forceinline__ int lb_compact(bool alive, int& idx, int* s_wcount) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(kFullMask, alive);
  __syncthreads();  // the previous round's readers of s_wcount are done
  if (lane == 0) s_wcount[warp] = __popc(bal);
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kLbThreads / 32; w++) {
    const int c = s_wcount[w];
    before += (w < warp) ? c : 0;
    total += c;
  }
  idx = before + __popc(bal & ((1u << lane) - 1u));
  return total;
}

template <bool kFromSums, bool kEq>
__global__ void __launch_bounds__(kLbThreads) dtw_lb_fused_kernel(LbFusedParams P) {
  __shared__ double s_q[kLbChunk], s_c[kLbChunk], s_h[kLbChunk];
  __shared__ int s_off[kLbThreads];
  __shared__ double s_mean[kLbThreads], s_std[kLbThreads], s_rstd[kLbThreads], s_lbq[kLbThreads], s_lbc[kLbThreads];
  __shared__ int s_wcount[kLbThreads / 32];
  const int tid = threadIdx.x;
  const int m = P.m;
  unsigned long long n = kFromSums ? *P.xin.count : *P.cin.count;
  const long long cap = kFromSums ? P.xin.cap : P.cin.cap;
  if ((long long)n > cap) n = (unsigned long long)cap;
  unsigned my_gate = 0;
  // Short lists (a selective query: a few hundred to a few thousand flagged windows): one WARP per candidate, lanes
  // split the m terms (coalesced loads, 64 iterations for m = 2048) — microseconds, where a cohort would walk its 2048
  // terms in 32 barrier-separated rounds on a single SM (0.2 ms).  Long lists take the cohort path below: there the
  // warp-per-candidate form would re-read 48 KB per candidate from L2.
  if (n <= (unsigned long long)P.warp_path_max) {
    const int lane = tid & 31;
    const unsigned long long gwarp = ((unsigned long long)blockIdx.x * kLbThreads + tid) >> 5;
    const unsigned long long n_warps = ((unsigned long long)gridDim.x * kLbThreads) >> 5;
    for (unsigned long long e = gwarp; e < n; e += n_warps) {
      int off;
      double mean = 0.0, stdv = 1.0;
      if (kFromSums) {
        off = P.xin.off[e];
        if (!cnsm_exact_gate(P.xin.ex[e], P.xin.ex2[e], m, P.meanQ, P.stdQ, P.alpha, P.inv_alpha, P.beta, mean, stdv)) continue;
        my_gate += (lane == 0) ? 1u : 0u;
      } else {
        off = P.cin.off[e];
        mean = P.cin.mean[e];
        stdv = P.cin.stdv[e];
      }
      const double rstd = 1.0 / stdv;
      const long long i0 = (long long)(off - P.first_global);
      const double* __restrict__ w = P.T + i0;
      if (kEq && m >= 6) {  // LB_KimFL (every lane computes it: six loads per end, all of them broadcasts)
        const double* __restrict__ q = P.q;
        const double x0 = (w[0] - mean) * rstd, x1 = (w[1] - mean) * rstd, x2 = (w[2] - mean) * rstd;
        const double y0 = (w[m - 1] - mean) * rstd, y1 = (w[m - 2] - mean) * rstd, y2 = (w[m - 3] - mean) * rstd;
        const double q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
        const double p0 = __ldg(q + m - 1), p1 = __ldg(q + m - 2), p2 = __ldg(q + m - 3);
        auto mn = [](double a, double b) { return min_nonneg(a, b); };
        double lb = fsq(x0, q0) + fsq(y0, p0);
        lb += mn(mn(fsq(x1, q0), fsq(x0, q1)), fsq(x1, q1));
        lb += mn(mn(fsq(y1, p0), fsq(y0, p1)), fsq(y1, p1));
        lb += mn(mn(mn(fsq(x0, q2), fsq(x1, q2)), fsq(x2, q2)), mn(fsq(x2, q1), fsq(x2, q0)));
        lb += mn(mn(mn(fsq(y0, p2), fsq(y1, p2)), fsq(y2, p2)), mn(fsq(y2, p1), fsq(y2, p0)));
        if (!le_nonneg(lb, P.eps2_hi)) continue;
      }
      double lbq = 0.0, lbc = 0.0;
      bool alive = true;
      for (int k0 = 0; k0 < m && alive; k0 += 256) {  // 8 terms per lane between two abandon tests
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int k = k0 + 32 * u + lane;
          if (k < m) {
            const double qk = __ldg(P.q + k);
            if (kEq) {
              const double up = __ldg(P.uq + k), lo = __ldg(P.lq + k);
              const double c = 0.5 * (up + lo);
              const double h = 0.5 * (up - lo) + 4.0 * 1.1102230246251565e-16 * (fabs(up) + fabs(lo) + 1.0);
              const double x = (w[k] - mean) * rstd;
              const double ex = fabs(x - c) - h;
              const double d = (__double2hiint(ex) < 0) ? 0.0 : ex;
              lbq = __fma_rn(d, d, lbq);
            }
            const double lo2 = (P.envL[i0 + k] - mean) * rstd, up2 = (P.envU[i0 + k] - mean) * rstd;
            const double a = qk - up2, b = lo2 - qk;
            const double d2 = (__double2hiint(a) >= 0) ? a : ((__double2hiint(b) >= 0) ? b : 0.0);
            lbc = __fma_rn(d2, d2, lbc);
          }
        }
        double tq = lbq, tc = lbc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          tq += __shfl_xor_sync(kFullMask, tq, o);
          tc += __shfl_xor_sync(kFullMask, tc, o);
        }
        alive = le_nonneg(tq, P.eps2_hi) && le_nonneg(tc, P.eps2_hi);
        if (k0 + 256 >= m && alive && lane == 0) cand_append(P.out, off, mean, stdv, kEq ? tq : 0.0);
      }
    }
    if (kFromSums) {
      const unsigned tot = __reduce_add_sync(kFullMask, my_gate);
      if ((tid & 31) == 0 && tot) atomicAdd(P.gate_pass, (unsigned long long)tot);
    }
    return;
  }
  // Cohort size: 256 candidates when the list fills the grid, down to 32 when it is short — a cohort lives on one SM and
  // its 2048-term walk is bound by that SM's FP64 pipe (256 candidates x 2048 terms x 12 operations = 0.1 M cycles), so
  // a short list is spread over as many SMs as it has warps.
  const unsigned long long per_cta = (n + gridDim.x - 1) / gridDim.x;
  const int B = (int)min((unsigned long long)kLbThreads, max(32ULL, (per_cta + 31ULL) & ~31ULL));
  for (unsigned long long base = (unsigned long long)blockIdx.x * B; base < n; base += (unsigned long long)gridDim.x * B) {
    const unsigned long long e = base + tid;
    bool alive = tid < B && e < n;
    int off = 0;
    double mean = 0.0, stdv = 1.0;
    if (alive) {
      if (kFromSums) {
        off = P.xin.off[e];
        alive = cnsm_exact_gate(P.xin.ex[e], P.xin.ex2[e], m, P.meanQ, P.stdQ, P.alpha, P.inv_alpha, P.beta, mean, stdv);
        my_gate += alive ? 1u : 0u;
      } else {
        off = P.cin.off[e];
        mean = P.cin.mean[e];
        stdv = P.cin.stdv[e];
      }
    }
    double rstd = 1.0;
    if (alive) {
      rstd = 1.0 / stdv;
      if (kEq && m >= 6) {  // LB_KimFL, K/utils/DtwUtils.java:149-189 (all five stages, no early return)
        const double* __restrict__ w = P.T + (off - P.first_global);
        const double* __restrict__ q = P.q;
        const double x0 = (w[0] - mean) * rstd, x1 = (w[1] - mean) * rstd, x2 = (w[2] - mean) * rstd;
        const double y0 = (w[m - 1] - mean) * rstd, y1 = (w[m - 2] - mean) * rstd, y2 = (w[m - 3] - mean) * rstd;
        const double q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
        const double p0 = __ldg(q + m - 1), p1 = __ldg(q + m - 2), p2 = __ldg(q + m - 3);
        auto mn = [](double a, double b) { return min_nonneg(a, b); };
        double lb = fsq(x0, q0) + fsq(y0, p0);
        lb += mn(mn(fsq(x1, q0), fsq(x0, q1)), fsq(x1, q1));
        lb += mn(mn(fsq(y1, p0), fsq(y0, p1)), fsq(y1, p1));
        lb += mn(mn(mn(fsq(x0, q2), fsq(x1, q2)), fsq(x2, q2)), mn(fsq(x2, q1), fsq(x2, q0)));
        lb += mn(mn(mn(fsq(y0, p2), fsq(y1, p2)), fsq(y2, p2)), mn(fsq(y2, p1), fsq(y2, p0)));
        alive = le_nonneg(lb, P.eps2_hi);
      }
    }
    if (alive) {  // the first two chunks of this candidate's window (series and envelope)
      const long long i0 = (long long)(off - P.first_global);
#pragma unroll
      for (int l = 0; l < 2 * kLbChunk * 8 / 128; l++) {
        if (16 * l < m) {
          if (kEq) prefetch_l2(P.T + i0 + 16 * l);
          prefetch_l2(P.envL + i0 + 16 * l);
          prefetch_l2(P.envU + i0 + 16 * l);
        }
      }
    }
    double lbq = 0.0, lbc = 0.0;
    int n_live = kLbThreads;  // (first compaction always moves)
    for (int k0 = 0;; k0 += kLbChunk) {
      // ---- compaction: survivors to the front (state travels through shared memory only when somebody dropped out)
      int idx;
      const int total = lb_compact(alive, idx, s_wcount);
      if (total != n_live) {
        if (alive) {
          s_off[idx] = off;
          s_mean[idx] = mean;
          s_std[idx] = stdv;
          s_rstd[idx] = rstd;
          s_lbq[idx] = lbq;
          s_lbc[idx] = lbc;
        }
        __syncthreads();
        n_live = total;
        alive = tid < n_live;
        if (alive) {
          off = s_off[tid];
          mean = s_mean[tid];
          stdv = s_std[tid];
          rstd = s_rstd[tid];
          lbq = s_lbq[tid];
          lbc = s_lbc[tid];
        }
      }
      if (n_live == 0 || k0 >= m) break;  // (CTA-uniform)
      // ---- the chunk's query terms: q, centre and half width of its envelope (half width rounded up)
      const int kc = min(kLbChunk, m - k0);
      if (tid < kLbChunk) {
        double qv = 0.0, c = 0.0, h = 0.0;
        if (tid < kc) {
          qv = __ldg(P.q + k0 + tid);
          const double up = __ldg(P.uq + k0 + tid), lo = __ldg(P.lq + k0 + tid);
          c = 0.5 * (up + lo);
          h = 0.5 * (up - lo) + 4.0 * 1.1102230246251565e-16 * (fabs(up) + fabs(lo) + 1.0);
        }
        s_q[tid] = qv;
        s_c[tid] = c;
        s_h[tid] = h;
      }
      __syncthreads();
      if (alive) {
        const long long i0 = (long long)(off - P.first_global) + k0;
        const double* __restrict__ w = P.T + i0;
        const double* __restrict__ el = P.envL + i0;
        const double* __restrict__ eu = P.envU + i0;
        // The windows were streamed long ago: their lines come from DRAM.  Ask for the lines of the chunk after the
        // next one now (two chunks = ~2 us of work for a lone warp), so that a sparse candidate list does not pay one
        // DRAM latency per group of loads (it did: 0.24 ms for a thousand candidates).
        if (k0 + 2 * kLbChunk < m) {
#pragma unroll
          for (int l = 0; l < kLbChunk * 8 / 128; l++) {
            if (kEq) prefetch_l2(w + 2 * kLbChunk + 16 * l);
            prefetch_l2(el + 2 * kLbChunk + 16 * l);
            prefetch_l2(eu + 2 * kLbChunk + 16 * l);
          }
        }
        auto term = [&](int k, double wv, double lv, double uv) {
          if (kEq) {
            const double x = (wv - mean) * rstd;
            const double ex = fabs(x - s_c[k]) - s_h[k];
            const double d = (__double2hiint(ex) < 0) ? 0.0 : ex;
            lbq = __fma_rn(d, d, lbq);
          }
          const double lo = (lv - mean) * rstd, up = (uv - mean) * rstd;
          const double qk = s_q[k];
          const double a = qk - up, b = lo - qk;
          const double d2 = (__double2hiint(a) >= 0) ? a : ((__double2hiint(b) >= 0) ? b : 0.0);
          lbc = __fma_rn(d2, d2, lbc);
        };
        if (kc == kLbChunk) {
#pragma unroll 1
          for (int kk = 0; kk < kLbChunk; kk += 16) {  // 48 loads in flight per thread: a lone cohort is latency-bound
            double wv[16], lv[16], uv[16];
#pragma unroll
            for (int u = 0; u < 16; u++) {
              wv[u] = kEq ? w[kk + u] : 0.0;
              lv[u] = el[kk + u];
              uv[u] = eu[kk + u];
            }
#pragma unroll
            for (int u = 0; u < 16; u++) term(kk + u, wv[u], lv[u], uv[u]);
          }
        } else {
          for (int kk = 0; kk < kc; kk++) term(kk, kEq ? w[kk] : 0.0, el[kk], eu[kk]);
        }
        alive = le_nonneg(lbq, P.eps2_hi) && le_nonneg(lbc, P.eps2_hi);
      }
    }
    if (alive) cand_append(P.out, off, mean, stdv, kEq ? lbq : 0.0);  // (after the last compaction: tid < n_live)
    __syncthreads();
  }
  if (kFromSums) {
    const unsigned tot = __reduce_add_sync(kFullMask, my_gate);
    if ((tid & 31) == 0 && tot) atomicAdd(P.gate_pass, (unsigned long long)tot);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Corner probe in front of the band DTW.  Most candidates that pass the lower bounds still abandon their DTW early:
// measured on BASELINE configs[3]-shaped queries, 10 M DTWs per query gave up after ~6000 of 409 000 cells, i.e. right
// after the top-left corner — and the wavefront kernels are at their worst there (a barrier or a shuffle per
// anti-diagonal that holds a handful of cells).  The probe computes exactly that corner, the K x K square
// (K = min(rho + 1, 128): every cell of it lies inside the band), one warp per candidate as a systolic array: lane l
// owns columns 4l .. 4l+3 and works on row t - l at step t, so a step needs two shuffles and K + K/4 steps finish
// the square.  Every warping path leaves the square through its last row or last column, and the rows it has not
// matched by then each cost at least their LB_Keogh(EQ) term, so
//     min( min_j D(K-1, j) + S(K),  min_i D(i, K-1) + S(i+1) ),   S(i) = sum of the Keogh terms of rows >= i,
// is a lower bound of the band DTW.  S comes from the candidate's Keogh total (formed by the fused LB kernel) minus
// the prefix over the first K rows, with the rounding slack on the safe side.  Candidates above eps^2 (1 + 1e-9) are
// dropped; the rest go to the full DTW kernels unchanged (which start over: K^2 of m (2 rho + 1) cells).
// Pruning only: answers and distances still come from dtw_band_*_kernel.
constexpr int kProbeWarps = 8;
constexpr int kProbeMaxK = 128;

struct ProbeParams {
  const double* __restrict__ T;
  int32_t first_global;
  int m, K;
  const double* __restrict__ q;
  const double* __restrict__ uq;
  const double* __restrict__ lq;
  double eps2_hi;
  CandList in, out;
  unsigned long long* n_cells;
  int min_count;  // lists shorter than this are passed through
  int skip_sub;   // the 64 x 64 sub-square was already probed (dtw_probe64_kernel): no early exit on it
  unsigned long long* handed_on;  // set by the first stage when it handed a short list on unprobed: the second stage does the same
};

__global__ void __launch_bounds__(kProbeWarps * 32) dtw_probe_kernel(ProbeParams P) {
  __shared__ double s_a[kProbeWarps][kProbeMaxK];
  __shared__ double s_S[kProbeWarps][kProbeMaxK + 4];
  __shared__ double s_exit[kProbeWarps][2 * kProbeMaxK + 2 * 64];  // exit cells: last column / last row of the square and of the sub-square
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = P.K;
  float* Af = reinterpret_cast<float*>(s_a[warp]);  // normalised rows, FP32
  double* S = s_S[warp];
  double* COL = s_exit[warp];
  double* ROW = COL + kProbeMaxK;
  double* COL1 = ROW + kProbeMaxK;
  double* ROW1 = COL1 + 64;
  double bu[4], bl[4];
  float bf[4];
  float bmax = 0.f;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const int col = min(4 * lane + c, K - 1);
    bf[c] = (float)__ldg(P.q + col);
    bmax = fmaxf(bmax, fabsf(bf[c]));
    bu[c] = __ldg(P.uq + col);  // (rows and columns share the index range 0..K-1: the lane's four Keogh terms)
    bl[c] = __ldg(P.lq + col);
  }
  bmax = __uint_as_float(__reduce_max_sync(kFullMask, __float_as_uint(bmax)));  // (non-negative floats order like their bit patterns)
  unsigned long long n = *P.in.count;
  if ((long long)n > P.in.cap) n = (unsigned long long)P.in.cap;
  if (n < (unsigned long long)P.min_count || (P.skip_sub && P.handed_on && *P.handed_on)) {  // a short list (mostly true matches): hand it on unprobed
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (unsigned long long)gridDim.x * blockDim.x)
      cand_append(P.out, P.in.off[e], P.in.mean[e], P.in.stdv[e], P.in.lb ? P.in.lb[e] : 0.0);
    return;
  }
  const int steps = K + (K + 3) / 4;
  const int last_lane = (K - 1) >> 2, last_c = (K - 1) & 3;
  unsigned long long probed_cells = 0;
  for (unsigned long long e = (unsigned long long)blockIdx.x * kProbeWarps + warp; e < n; e += (unsigned long long)gridDim.x * kProbeWarps) {
    const int32_t off = P.in.off[e];
    const double mean = P.in.mean[e], stdv = P.in.stdv[e];
    const double total = P.in.lb ? P.in.lb[e] : 0.0;
    const double rstd = 1.0 / stdv;
    const double* __restrict__ w = P.T + (off - P.first_global);
    __syncwarp();
    // rows 4 lane .. 4 lane + 3: normalised samples and their Keogh terms
    double kt[4];
    float amax = 0.f;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int k = 4 * lane + c;
      kt[c] = 0.0;
      if (k < K) {
        const double a = (w[k] - mean) * rstd;
        Af[k] = (float)a;
        amax = fmaxf(amax, fabsf((float)a));
        const double dd = (a > bu[c]) ? (a - bu[c]) : ((a < bl[c]) ? (a - bl[c]) : 0.0);
        kt[c] = dd * dd;
      }
    }
    amax = __uint_as_float(__reduce_max_sync(kFullMask, __float_as_uint(amax)));
    // The square itself runs in FP32 with every rounding on the safe side (a 4-column systolic step is issue bound: FP32
    // min / fused multiply-add are one instruction where the binary64 forms were five).  |a - b| is lowered by
    // e >= the two conversion errors plus the subtraction's own rounding ((|a| + |b|) 2^-22 covers 2 * 2^-24 * (|a| + |b|)
    // four times over), the product and the sum round towards zero, the minimum is exact: every cell is a lower bound
    // of the binary64 cell, and so is the exit bound built from them.
    const float e_sub = __fmul_ru(__fadd_ru(amax, bmax), 2.3841858e-7f);  // 2^-22
    // S(i) = total - sum_{k < i} term_k, kept below the true remainder: total shrunk, prefix grown, clamped at 0
    {
      const double mine = (kt[0] + kt[1]) + (kt[2] + kt[3]);
      double incl = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(kFullMask, incl, o);
        if (lane >= o) incl += t;
      }
      double pre = incl - mine;  // prefix before row 4 lane
      const double tot = total * (1.0 - 1e-10);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int k = 4 * lane + c;
        if (k <= K) S[k] = fmax(tot - pre * (1.0 + 1e-10), 0.0);
        pre += kt[c];
      }
      if (lane == 31 && K == kProbeMaxK) S[K] = fmax(tot - pre * (1.0 + 1e-10), 0.0);
    }
    __syncwarp();
    // The systolic sweep.  Nothing but the recurrence sits in the loop: the exit cells (last row / last column of the
    // square and of the K1 sub-square) are parked in shared memory with predicated stores and reduced afterwards by
    // all lanes.  Cell (0, 0) needs no special case: its diagonal predecessor is seeded with 0.
    const float kInfF = 9.9999e19f;  // (below the reference INF = 1e20: a lower bound of it)
    float prev[4] = {kInfF, kInfF, kInfF, kInfF};
    float last3 = kInfF, last3_prev = kInfF;
    // The same bound on the sub-square K1 x K1 (K1 = 64 when K >= 96) is complete after step K1 - 1 + K1/4 - 1: most
    // candidates are already over eps^2 there (86 % on the measured queries) and skip the remaining 40 % of the steps.
    const int K1 = (K >= 96 && !P.skip_sub) ? 64 : 0;
    const int k1_lane = (K1 >> 2) - 1, k1_step = K1 - 1 + k1_lane;
    bool dead = false;
    auto exit_bound = [&](const double* col, const double* row, int kk) {  // kk x kk square: min over its exit cells
      double bb = kDtwInf;
      const double s_last = S[kk];
      for (int i = lane; i < kk; i += 32) bb = min_nonneg(bb, min_nonneg(col[i] + S[i + 1], row[i] + s_last));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) bb = min_nonneg(bb, __shfl_xor_sync(kFullMask, bb, o));
      return bb;
    };
    for (int t = 0; t < steps; t++) {
      if (K1 && t == k1_step + 1) {  // (warp-uniform)
        __syncwarp();
        if (!le_nonneg(exit_bound(COL1, ROW1, K1), P.eps2_hi)) {
          dead = true;
          break;
        }
      }
      float x_left = __shfl_up_sync(kFullMask, last3, 1), x_diag = __shfl_up_sync(kFullMask, last3_prev, 1);
      if (lane == 0) {
        x_left = kInfF;
        x_diag = (t == 0) ? 0.f : kInfF;
      }
      const int r = t - lane;
      if (r >= 0 && r < K) {
        const float ar = Af[r];
        float cur[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const float d = fmaxf(__fadd_rd(fabsf(ar - bf[c]), -e_sub), 0.f);
          const float up = prev[c];
          const float v = __fmaf_rz(d, d, fminf(fminf(x_left, up), x_diag));
          x_diag = up;
          x_left = v;
          cur[c] = v;
          prev[c] = v;
        }
        last3_prev = last3;
        last3 = cur[3];
        if (lane == last_lane) COL[r] = (double)((last_c == 0) ? cur[0] : (last_c == 1) ? cur[1] : (last_c == 2) ? cur[2] : cur[3]);
        if (r == K - 1) {
#pragma unroll
          for (int c = 0; c < 4; c++) ROW[min(4 * lane + c, kProbeMaxK - 1)] = (double)cur[c];
        }
        if (K1) {
          if (lane == k1_lane && r < K1) COL1[r] = (double)cur[3];
          if (r == K1 - 1 && lane <= k1_lane) {
#pragma unroll
            for (int c = 0; c < 4; c++) ROW1[4 * lane + c] = (double)cur[c];
          }
        }
      }
    }
    __syncwarp();
    probed_cells += dead ? (unsigned long long)K1 * K1 : (unsigned long long)K * K;
    if (!dead) {
      const double best = exit_bound(COL, ROW, K);
      if (lane == 0 && le_nonneg(best, P.eps2_hi)) cand_append(P.out, off, mean, stdv, total);
    }
  }
  if (lane == 0 && probed_cells) atomicAdd(P.n_cells, probed_cells);
}

// First stage of the probe for wide bands (K >= 96): the 64 x 64 sub-square alone, TWO candidates per warp (a half-warp
// of 16 lanes x 4 columns each).  86 % of the measured candidates are already over eps^2 on this sub-square; they cost
// half a warp for 79 steps here instead of a whole warp for 79 of its 129 steps, and only the survivors go through
// dtw_probe_kernel's full square.  Same arithmetic and the same bound as there (FP32 cells rounded to the safe side,
// exit cells + remaining Keogh).
__global__ void __launch_bounds__(kProbeWarps * 32) dtw_probe64_kernel(ProbeParams P) {
  constexpr int K1 = 64;
  __shared__ float s_af[kProbeWarps][2][K1];
  __shared__ double s_S[kProbeWarps][2][K1 + 2];
  __shared__ double s_exit[kProbeWarps][2][2 * K1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, hl = lane & 15;
  float* Af = s_af[warp][half];
  double* S = s_S[warp][half];
  double* COL = s_exit[warp][half];
  double* ROW = COL + K1;
  double bu[4], bl[4];
  float bf[4];
  float bmax = 0.f;
#pragma unroll
  for (int c = 0; c < 4; c++) {
    const int col = 4 * hl + c;
    bf[c] = (float)__ldg(P.q + col);
    bmax = fmaxf(bmax, fabsf(bf[c]));
    bu[c] = __ldg(P.uq + col);
    bl[c] = __ldg(P.lq + col);
  }
  bmax = __uint_as_float(__reduce_max_sync(kFullMask, __float_as_uint(bmax)));
  unsigned long long n = *P.in.count;
  if ((long long)n > P.in.cap) n = (unsigned long long)P.in.cap;
  if (n < (unsigned long long)P.min_count) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && P.handed_on) *P.handed_on = 1ULL;
    for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (unsigned long long)gridDim.x * blockDim.x)
      cand_append(P.out, P.in.off[e], P.in.mean[e], P.in.stdv[e], P.in.lb ? P.in.lb[e] : 0.0);
    return;
  }
  const float kInfF = 9.9999e19f;
  unsigned long long probed = 0;
  for (unsigned long long e0 = 2ULL * ((unsigned long long)blockIdx.x * kProbeWarps + warp); e0 < n; e0 += 2ULL * gridDim.x * kProbeWarps) {
    const bool valid = e0 + half < n;
    const unsigned long long e = valid ? e0 + half : n - 1;  // (an odd tail: the idle half repeats the last candidate)
    const int32_t off = P.in.off[e];
    const double mean = P.in.mean[e], stdv = P.in.stdv[e];
    const double total = P.in.lb ? P.in.lb[e] : 0.0;
    const double rstd = 1.0 / stdv;
    const double* __restrict__ w = P.T + (off - P.first_global);
    __syncwarp();
    double kt[4];
    float amax = 0.f;
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int k = 4 * hl + c;
      const double a = (w[k] - mean) * rstd;
      Af[k] = (float)a;
      amax = fmaxf(amax, fabsf((float)a));
      const double dd = (a > bu[c]) ? (a - bu[c]) : ((a < bl[c]) ? (a - bl[c]) : 0.0);
      kt[c] = dd * dd;
    }
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(kFullMask, amax, o, 16));
    const float e_sub = __fmul_ru(__fadd_ru(amax, bmax), 2.3841858e-7f);  // 2^-22, as in dtw_probe_kernel
    {
      const double mine = (kt[0] + kt[1]) + (kt[2] + kt[3]);
      double incl = mine;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        const double t = __shfl_up_sync(kFullMask, incl, o, 16);
        if (hl >= o) incl += t;
      }
      double pre = incl - mine;
      const double tot = total * (1.0 - 1e-10);
#pragma unroll
      for (int c = 0; c < 4; c++) {
        S[4 * hl + c] = fmax(tot - pre * (1.0 + 1e-10), 0.0);
        pre += kt[c];
      }
      if (hl == 15) S[K1] = fmax(tot - pre * (1.0 + 1e-10), 0.0);
    }
    __syncwarp();
    float prev[4] = {kInfF, kInfF, kInfF, kInfF};
    float last3 = kInfF, last3_prev = kInfF;
#pragma unroll 1
    for (int t = 0; t < K1 + 15; t++) {
      float x_left = __shfl_up_sync(kFullMask, last3, 1, 16), x_diag = __shfl_up_sync(kFullMask, last3_prev, 1, 16);
      if (hl == 0) {
        x_left = kInfF;
        x_diag = (t == 0) ? 0.f : kInfF;
      }
      const int r = t - hl;
      if (r >= 0 && r < K1) {
        const float ar = Af[r];
        float cur[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          const float d = fmaxf(__fadd_rd(fabsf(ar - bf[c]), -e_sub), 0.f);
          const float up = prev[c];
          const float v = __fmaf_rz(d, d, fminf(fminf(x_left, up), x_diag));
          x_diag = up;
          x_left = v;
          cur[c] = v;
          prev[c] = v;
        }
        last3_prev = last3;
        last3 = cur[3];
        if (hl == 15) COL[r] = (double)cur[3];
        if (r == K1 - 1) {
#pragma unroll
          for (int c = 0; c < 4; c++) ROW[4 * hl + c] = (double)cur[c];
        }
      }
    }
    __syncwarp();
    double bb = kDtwInf;
    const double s_last = S[K1];
    for (int i = hl; i < K1; i += 16) bb = min_nonneg(bb, min_nonneg(COL[i] + S[i + 1], ROW[i] + s_last));
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) bb = min_nonneg(bb, __shfl_xor_sync(kFullMask, bb, o, 16));
    if (valid && hl == 0) {
      probed++;
      if (le_nonneg(bb, P.eps2_hi)) cand_append(P.out, off, mean, stdv, total);
    }
  }
  probed = __reduce_add_sync(kFullMask, (unsigned)probed);
  if (lane == 0 && probed) atomicAdd(P.n_cells, probed * (unsigned long long)(K1 * K1));
}

// lowerUpperLemire on the device (K/utils/DtwUtils.java:50-91): l[i] = min, u[i] = max of t[max(0,i-r) .. min(len-1,i+r)]
// over a region of the series, exact doubles.  One CTA per tile of kEnvTile outputs: the tile plus its halo goes to
// shared memory as order-preserving integer keys, a sparse table is built level by level (doubling), and every output
// combines two overlapping power-of-two ranges.
constexpr int kEnvTile = 2048;
__device__ __forceinline__ long long env_key(double x) {  // monotone: a <= b  <=>  key(a) <= key(b)  (no NaNs)
  const long long b = __double_as_longlong(x);
  return b ^ ((b >> 63) & 0x7fffffffffffffffLL);
}
__device__ __forceinline__ double env_unkey(long long k) { return __longlong_as_double(k ^ ((k >> 63) & 0x7fffffffffffffffLL)); }

__global__ void __launch_bounds__(256) envelope_kernel(const double* __restrict__ t, int len, int r, double* __restrict__ lo_out,
                                                       double* __restrict__ up_out) {
  extern __shared__ long long env_smem[];
  const int span = kEnvTile + 2 * r;  // keys of samples [i0 - r, i0 + kEnvTile + r)
  long long* mx = env_smem;
  long long* mn = env_smem + span;
  const int i0 = (int)blockIdx.x * kEnvTile;
  for (int k = threadIdx.x; k < span; k += blockDim.x) {
    const int g = i0 - r + k;
    const bool in = g >= 0 && g < len;
    const long long key = in ? env_key(t[g]) : 0;
    mx[k] = in ? key : LLONG_MIN;  // outside the region: neutral elements (the reference clamps the range)
    mn[k] = in ? key : LLONG_MAX;
  }
  __syncthreads();
  const int L = 2 * r + 1;
  int levels = 0;
  while ((2 << levels) <= L) levels++;  // 2^levels <= L < 2^(levels+1)
  for (int lv = 0; lv < levels; lv++) {  // after level lv: mx[k] = max over [k, k + 2^(lv+1))
    const int step = 1 << lv;
    long long a[ (kEnvTile + 1024 + 255) / 256 ], b[ (kEnvTile + 1024 + 255) / 256 ];
    int c = 0;
    for (int k = threadIdx.x; k < span; k += blockDim.x, c++) {
      const int k2 = min(k + step, span - 1);
      a[c] = max(mx[k], (k + step < span) ? mx[k2] : LLONG_MIN);
      b[c] = min(mn[k], (k + step < span) ? mn[k2] : LLONG_MAX);
    }
    __syncthreads();
    c = 0;
    for (int k = threadIdx.x; k < span; k += blockDim.x, c++) {
      mx[k] = a[c];
      mn[k] = b[c];
    }
    __syncthreads();
  }
  const int p2 = 1 << levels;
  for (int k = threadIdx.x; k < kEnvTile; k += blockDim.x) {
    const int i = i0 + k;
    if (i >= len) break;
    // window [i - r, i + r] = smem indices [k, k + 2r]: two ranges of length p2
    const long long u = max(mx[k], mx[k + L - p2]);
    const long long l = min(mn[k], mn[k + L - p2]);
    up_out[i] = env_unkey(u);
    lo_out[i] = env_unkey(l);
  }
}

// min of two non-negative doubles through their bit patterns (for x, y >= +0 the IEEE order is the unsigned integer
// order).  DTW costs are sums of squares, never negative and never -0.0, so this equals DtwUtils.min — and it runs on
// the integer pipe: DMNMX issues at only ~1/5 of the DADD rate on B200 (tools/fp64_peak.cu) and two of them per cell
// were the kernel's bottleneck.
__device__ __forceinline__ double umin_pos(double a, double b) {
  const unsigned long long x = (unsigned long long)__double_as_longlong(a), y = (unsigned long long)__double_as_longlong(b);
  return __longlong_as_double((long long)(x < y ? x : y));
}

struct DtwParams {
  const double* __restrict__ T;
  int32_t first_global;
  int m;
  int rho;
  const double* __restrict__ q;   // natural order
  const double* __restrict__ uq;  // query envelope (radius rho), for the cumulative LB_Keogh remainder
  const double* __restrict__ lq;
  double eps2, eps2_hi;
  CandList in;
  AnswerSink sink;
  unsigned long long* n_abandoned;
  unsigned long long* n_cells;  // band cells evaluated (for the FP64 roofline: 5 flops per cell)
  // candidates below `coop_limit` belong to the CTA-cooperative kernel, the rest to the warp-per-candidate kernel
  // (the host sends a query's whole list to one of the two: 0 or "all")
  long long coop_limit;
};

template <int R>
__global__ void __launch_bounds__(256) dtw_band_kernel(DtwParams P) {
  extern __shared__ double dtw_smem[];
  const int m = P.m, rho = P.rho;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  // per warp: the (normalised) window A[m] and the cumulative LB_Keogh remainder sampled every 8th position
  const int cbs_len = (m >> 3) + 2;
  const int warp_doubles = m + ((cbs_len + 1) & ~1);
  double* B = dtw_smem;                                        // query
  double* A = dtw_smem + m + (size_t)warp * warp_doubles;      // this warp's window
  double* CBS = A + m;                                         // CBS[t] = sum_{k >= 8t} contribution of A[k]
  for (int k = threadIdx.x; k < m; k += blockDim.x) B[k] = P.q[k];
  __syncthreads();

  unsigned long long n = *P.in.count;
  if ((long long)n > P.in.cap) n = (unsigned long long)P.in.cap;
  const int tgt_u = rho;  // final cell (m-1, m-1): i-j = 0
  const int tgt_pair = tgt_u >> 1;

  // candidate e -> (warp, block) with the block index fastest: a short list spreads over all SMs, one warp per
  // scheduler, instead of filling the first few CTAs
  for (unsigned long long e = (unsigned long long)P.coop_limit + (unsigned long long)warp * gridDim.x + blockIdx.x; e < n;
       e += (unsigned long long)gridDim.x * n_warps) {
    const int32_t off = P.in.off[e];
    const double mean = P.in.mean[e], stdv = P.in.stdv[e];
    const double* __restrict__ w = P.T + (off - P.first_global);
    __syncwarp();
    // window -> shared memory; each lane also sums the LB_Keogh contributions (K/utils/DtwUtils.java:206-222) of the
    // 8-sample groups it owns (group t = samples 8t..8t+7, lane t % 32)
    for (int t = lane; t < cbs_len; t += 32) {
      double grp = 0.0;
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int k = 8 * t + u;
        if (k < m) {
          const double a = xdiv(xsub(w[k], mean), stdv);  // NormQueryEngineDtw.java:564-567
          A[k] = a;
          const double up = __ldg(P.uq + k), lo = __ldg(P.lq + k);
          const double dd = (a > up) ? (a - up) : ((a < lo) ? (a - lo) : 0.0);
          grp += dd * dd;
        }
      }
      CBS[t] = grp;
    }
    __syncwarp();
    {  // suffix sums over the groups (the reference's cb, K/QueryEngineDtw.java:430-441, at every 8th position)
      const int per = (cbs_len + 31) / 32;
      const int t0 = lane * per, t1 = min(cbs_len, t0 + per);
      double run = 0.0;
      for (int t = t1 - 1; t >= t0; t--) {
        run += CBS[t];
        CBS[t] = run;
      }
      double tot = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double tt = __shfl_down_sync(kFullMask, tot, o);
        if (lane + o < 32) tot += tt;
      }
      const double right = tot - run;
      for (int t = t0; t < t1; t++) CBS[t] += right;
    }
    __syncwarp();

    double Ev[R], Od[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
      Ev[r] = kDtwInf;
      Od[r] = kDtwInf;
    }
    const int u0 = 2 * lane * R;  // band coordinate of this lane's first even cell
    const int last = 2 * m - 2;
    // Operands in registers: av[r] = A[i0 + r], bv[r] = B[j0 - r] for the lane's R cells of the current step.  Going
    // from an even to an odd diagonal every i grows by one (av shifts, one new element), from odd to even every j
    // grows by one (bv shifts) — one LDS per step instead of 2R.
    double av[R], bv[R];
    auto ld = [&](const double* base, int idx) { return base[min(max(idx, 0), m - 1)]; };  // clamped: unused when invalid
    {
      const int par0 = rho & 1;  // parity of the first step's cells (d = 0)
      const int i0 = (u0 + par0 - rho) >> 1, j0 = -i0;
#pragma unroll
      for (int r = 0; r < R; r++) {
        av[r] = ld(A, i0 + r);
        bv[r] = ld(B, j0 - r);
      }
    }
    unsigned band_even = 0, band_odd = 0;  // this lane's cells that lie inside the band
#pragma unroll
    for (int r = 0; r < R; r++) {
      band_even |= (u0 + 2 * r <= 2 * rho) ? (1u << r) : 0u;
      band_odd |= (u0 + 2 * r + 1 <= 2 * rho) ? (1u << r) : 0u;
    }
    bool abandoned = false;
    unsigned long long cells = 0;  // (warp-uniform) cells inside matrix and band on the diagonals walked so far
    // Interior diagonals (rho + 2 <= d, d + 1 < 2m - 2 - rho) are walked two at a time by the tight loop below: every
    // band cell lies inside the matrix there, so no index test survives; cells beyond the band hold values >= INF
    // (INF plus non-negative costs), which a band cell's min never selects because one of its three predecessors is
    // always a band cell — exactly what the reference's INF neighbours do.
    const int fast_begin = rho + 2, fast_pairs = m - 2 - rho;
    for (int d = 0; d <= last; d++) {
      if (d == fast_begin && fast_pairs > 0) {
        const int i_base = (d + u0 - rho) >> 1;  // (d + rho) is even here
        int ia = i_base + R, jb = d - i_base;    // next A element (odd diagonal), next B element (even diagonal)
        double a_new = A[min(ia, m - 1)], b_new = B[max(jb, 0)];
        int p = 0;
        for (; p < fast_pairs; p++) {
          if ((p & 7) == 0 && p > 0) {
            const int dc = d + 2 * p;
            double mn = kDtwInf;
#pragma unroll
            for (int r = 0; r < R; r++) mn = umin_pos(mn, umin_pos(Ev[r], Od[r]));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mn = umin_pos(mn, __shfl_xor_sync(kFullMask, mn, o));
            const int imax = min(m - 1, (dc - 1 + rho) >> 1);
            if (mn + CBS[(imax + 1 + 7) >> 3] > P.eps2_hi) {
              abandoned = true;
              break;
            }
          }
          // even diagonal: j grows by one
          double left = __shfl_up_sync(kFullMask, Od[R - 1], 1);
          if (lane == 0) left = kDtwInf;
#pragma unroll
          for (int r = R - 1; r > 0; r--) bv[r] = bv[r - 1];
          bv[0] = b_new;
          jb++;
          b_new = B[min(max(jb, 0), m - 1)];
#pragma unroll
          for (int r = 0; r < R; r++) {
            const double own = umin_pos(Od[r], Ev[r]);  // ready before the neighbour arrives
            const double x = (r == 0) ? left : Od[r - 1];
            Ev[r] = xadd(umin_pos(x, own), xsqdist(av[r], bv[r]));
          }
          // odd diagonal: i grows by one
          double right = __shfl_down_sync(kFullMask, Ev[0], 1);
          if (lane == 31) right = kDtwInf;
#pragma unroll
          for (int r = 0; r < R - 1; r++) av[r] = av[r + 1];
          av[R - 1] = a_new;
          ia++;
          a_new = A[min(ia, m - 1)];
#pragma unroll
          for (int r = 0; r < R; r++) {
            const double own = umin_pos(Ev[r], Od[r]);
            const double y = (r == R - 1) ? right : Ev[r + 1];
            const double v = xadd(umin_pos(y, own), xsqdist(av[r], bv[r]));
            Od[r] = ((band_odd >> r) & 1u) ? v : kDtwInf;
          }
        }
        cells += (unsigned long long)p * (unsigned long long)(2 * rho + 1);
        if (abandoned) break;
        d += 2 * p;  // p == fast_pairs: the edge diagonals follow
      }
      {
        const int i_lo = max(max(0, d - (m - 1)), (d - rho + 1) >> 1), i_hi = min(min(m - 1, d), (d + rho) >> 1);
        cells += (unsigned long long)max(0, i_hi - i_lo + 1);
      }
      // Early abandon (the reference abandons per row with min_cost + cb[i+r+1], DtwUtils.java:324-326).  On the
      // wavefront: every warping path crosses one of the two most recent anti-diagonals, cell values only grow
      // along a path, and data points beyond imax = (d-1+rho)/2 have not been matched yet, so
      // min(cells on the last two diagonals) + cb[imax+1] is a lower bound of the final distance.
      if ((d & 7) == 0 && d > 0) {
        double mn = kDtwInf;
#pragma unroll
        for (int r = 0; r < R; r++) mn = umin_pos(mn, umin_pos(Ev[r], Od[r]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = umin_pos(mn, __shfl_xor_sync(kFullMask, mn, o));
        // (rows on diagonal d - 1: i <= d - 1 while the wavefront is still in the top-left corner, i <= (d - 1 + rho) / 2
        // once the band binds)
        const int imax = min(m - 1, min(d - 1, (d - 1 + rho) >> 1));
        const double rest = CBS[(imax + 1 + 7) >> 3];  // first sampled position >= imax+1: a (slightly smaller) valid remainder
        if (mn + rest > P.eps2_hi) {
          abandoned = true;
          break;
        }
      }
      // In the interior (rho <= d <= 2m-2-rho) every band cell lies inside the matrix, so validity is the lane-constant
      // band bit and the per-cell index arithmetic drops out of the dependent chain; edges use the general form.
      const bool interior = (d > rho) && (d < last - rho);
      if (((d + rho) & 1) == 0) {
        double left = __shfl_up_sync(kFullMask, Od[R - 1], 1);
        if (lane == 0) left = kDtwInf;
        // i = (d + u - rho)/2, j = d - i ; consecutive pairs: i+1, j-1
        const int s = d + u0 - rho;
        const int i0 = s >> 1;  // s is even here
        if (d > 0) {            // odd -> even: j grows by one
#pragma unroll
          for (int r = R - 1; r > 0; r--) bv[r] = bv[r - 1];
          bv[0] = ld(B, d - i0);
        }
        if (interior) {
#pragma unroll
          for (int r = 0; r < R; r++) {
            const double c = xsqdist(av[r], bv[r]);
            const double x = (r == 0) ? left : Od[r - 1];
            const double v = xadd(umin_pos(umin_pos(x, Od[r]), Ev[r]), c);
            Ev[r] = ((band_even >> r) & 1u) ? v : kDtwInf;
          }
        } else {
#pragma unroll
          for (int r = 0; r < R; r++) {
            const int i = i0 + r, j = d - i;
            const bool valid = (u0 + 2 * r <= 2 * rho) && i >= 0 && j >= 0 && i < m && j < m;
            double v = kDtwInf;
            if (valid) {
              const double c = xsqdist(av[r], bv[r]);
              const double x = (r == 0) ? left : Od[r - 1];
              const double y = Od[r];
              const double z = Ev[r];
              v = (d == 0) ? c : xadd(umin_pos(umin_pos(x, y), z), c);
            }
            Ev[r] = v;
          }
        }
      } else {
        double right = __shfl_down_sync(kFullMask, Ev[0], 1);
        if (lane == 31) right = kDtwInf;
        const int s = d + u0 + 1 - rho;
        const int i0 = s >> 1;
        if (d > 0) {            // even -> odd: i grows by one
#pragma unroll
          for (int r = 0; r < R - 1; r++) av[r] = av[r + 1];
          av[R - 1] = ld(A, i0 + R - 1);
        }
        if (interior) {
#pragma unroll
          for (int r = 0; r < R; r++) {
            const double c = xsqdist(av[r], bv[r]);
            const double y = (r == R - 1) ? right : Ev[r + 1];
            const double v = xadd(umin_pos(umin_pos(Ev[r], y), Od[r]), c);
            Od[r] = ((band_odd >> r) & 1u) ? v : kDtwInf;
          }
        } else {
#pragma unroll
          for (int r = 0; r < R; r++) {
            const int i = i0 + r, j = d - i;
            const bool valid = (u0 + 2 * r + 1 <= 2 * rho) && i >= 0 && j >= 0 && i < m && j < m;
            double v = kDtwInf;
            if (valid) {
              const double c = xsqdist(av[r], bv[r]);
              const double x = Ev[r];
              const double y = (r == R - 1) ? right : Ev[r + 1];
              const double z = Od[r];
              v = (d == 0) ? c : xadd(umin_pos(umin_pos(x, y), z), c);
            }
            Od[r] = v;
          }
        }
      }
    }
    if (lane == 0) atomicAdd(P.n_cells, cells);
    if (abandoned) {
      if (lane == 0) atomicAdd(P.n_abandoned, 1ULL);
      continue;
    }
    double res = kDtwInf;
#pragma unroll
    for (int r = 0; r < R; r++) {
      if (lane * R + r == tgt_pair) res = (tgt_u & 1) ? Od[r] : Ev[r];
    }
    res = __shfl_sync(kFullMask, res, tgt_pair / R);
    if (lane == 0 && res <= P.eps2) P.sink.emit(off, xsqrt(res));
  }
}


// ---------------------------------------------------------------------------------------------------------------
// CTA-cooperative band DTW: the same anti-diagonal sweep with the band split over kCoopThreads threads, one DTW per
// CTA at a time.  Thread t owns RC consecutive (even, odd) band pairs; the two values that cross a thread boundary
// per diagonal go through shared memory (one block barrier per diagonal).  A diagonal then costs one or two cells per
// thread instead of R = 4..16 per lane: measured 0.33 ms per m = 2048 / rho = 102 DTW (153 cycles per diagonal:
// barrier + LDS + two integer mins + DADD) against 0.52 ms on one warp.  Same operands, same unfused operations, same
// INF = 1e20 conventions as dtw_band_kernel: bit-identical results.
constexpr int kCoopThreads = 128;

template <int RC>
__global__ void __launch_bounds__(kCoopThreads) dtw_band_coop_kernel(DtwParams P) {
  extern __shared__ double coop_smem[];
  __shared__ double s_xo[kCoopThreads + 2], s_xe[kCoopThreads + 2];  // boundary cells: odd diagonals -> next even, even -> next odd
  __shared__ double s_red[kCoopThreads / 32];
  const int m = P.m, rho = P.rho;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cbs_len = (m >> 3) + 2;
  const double* __restrict__ B = P.q;  // query: read through L1 (every CTA of the grid reads the same 8 m bytes; a private
                                       // shared-memory copy halved the CTAs per SM)
  double* A = coop_smem;           // normalised window
  double* CBS = A + m;             // cumulative LB_Keogh remainder sampled every 8th position
  unsigned long long n = *P.in.count;
  if ((long long)n > P.in.cap) n = (unsigned long long)P.in.cap;
  if ((long long)n > P.coop_limit) n = (unsigned long long)P.coop_limit;
  const int tgt_u = rho, tgt_pair = tgt_u >> 1;
  const int u0 = 2 * tid * RC;
  const int last = 2 * m - 2;
  for (unsigned long long e = blockIdx.x; e < n; e += gridDim.x) {
    const int32_t off = P.in.off[e];
    const double mean = P.in.mean[e], stdv = P.in.stdv[e];
    const double* __restrict__ w = P.T + (off - P.first_global);
    __syncthreads();
    // window -> shared memory (the reference's arithmetic, NormQueryEngineDtw.java:564-567) + LB_Keogh group sums
    for (int t = tid; t < cbs_len; t += kCoopThreads) {
      double grp = 0.0;
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int k = 8 * t + u;
        if (k < m) {
          const double a = xdiv(xsub(w[k], mean), stdv);
          A[k] = a;
          const double up = __ldg(P.uq + k), lo = __ldg(P.lq + k);
          const double dd = (a > up) ? (a - up) : ((a < lo) ? (a - lo) : 0.0);
          grp += dd * dd;
        }
      }
      CBS[t] = grp;
    }
    s_xo[tid + 1] = kDtwInf;
    s_xe[tid] = kDtwInf;
    if (tid == 0) {
      s_xo[0] = kDtwInf;
      s_xe[kCoopThreads] = kDtwInf;
    }
    __syncthreads();
    if (warp == 0) {  // suffix sums over the groups (sequential chunks per lane + a shuffle scan)
      const int per = (cbs_len + 31) / 32;
      const int t0 = lane * per, t1 = min(cbs_len, t0 + per);
      double run = 0.0;
      for (int t = t1 - 1; t >= t0; t--) {
        run += CBS[t];
        CBS[t] = run;
      }
      double tot = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double tt = __shfl_down_sync(kFullMask, tot, o);
        if (lane + o < 32) tot += tt;
      }
      const double right = tot - run;
      for (int t = t0; t < t1; t++) CBS[t] += right;
    }
    __syncthreads();

    double Ev[RC], Od[RC];
#pragma unroll
    for (int r = 0; r < RC; r++) Ev[r] = Od[r] = kDtwInf;
    unsigned band_odd = 0;  // this thread's odd cells that lie inside the band
#pragma unroll
    for (int r = 0; r < RC; r++) band_odd |= (u0 + 2 * r + 1 <= 2 * rho) ? (1u << r) : 0u;
    bool abandoned = false;
    unsigned long long cells = 0;
    // lower bound of the final distance from the two most recent diagonals (see dtw_band_kernel); block-uniform
    auto hopeless = [&](int d) {
      double mn = kDtwInf;
#pragma unroll
      for (int r = 0; r < RC; r++) mn = umin_pos(mn, umin_pos(Ev[r], Od[r]));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mn = umin_pos(mn, __shfl_xor_sync(kFullMask, mn, o));
      if (lane == 0) s_red[warp] = mn;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kCoopThreads / 32; k++) mn = umin_pos(mn, s_red[k]);
      const int imax = min(m - 1, min(d - 1, (d - 1 + rho) >> 1));  // rows on diagonal d - 1 (corner: i <= d - 1)
      return mn + CBS[(imax + 1 + 7) >> 3] > P.eps2_hi;
    };
    // one edge diagonal (some band cells fall outside the matrix): the general form with per-cell index tests
    auto edge_step = [&](int d) {
      const int par = (d + rho) & 1;
      const int i0 = (d + u0 + par - rho) >> 1;
      if (par == 0) {
        const double left = s_xo[tid];  // the previous thread's last odd cell of diagonal d - 1
#pragma unroll
        for (int r = 0; r < RC; r++) {
          const int i = i0 + r, j = d - i;
          const bool valid = (u0 + 2 * r <= 2 * rho) && i >= 0 && j >= 0 && i < m && j < m;
          double v = kDtwInf;
          if (valid) {
            const double c = xsqdist(A[i], B[j]);
            const double x = (r == 0) ? left : Od[r - 1];
            v = (d == 0) ? c : xadd(umin_pos(umin_pos(x, Od[r]), Ev[r]), c);
          }
          Ev[r] = v;
        }
        s_xe[tid] = Ev[0];
      } else {
        const double right = s_xe[tid + 1];  // the next thread's first even cell of diagonal d - 1
#pragma unroll
        for (int r = 0; r < RC; r++) {
          const int i = i0 + r, j = d - i;
          const bool valid = (u0 + 2 * r + 1 <= 2 * rho) && i >= 0 && j >= 0 && i < m && j < m;
          double v = kDtwInf;
          if (valid) {
            const double c = xsqdist(A[i], B[j]);
            const double y = (r == RC - 1) ? right : Ev[r + 1];
            v = (d == 0) ? c : xadd(umin_pos(umin_pos(Ev[r], y), Od[r]), c);
          }
          Od[r] = v;
        }
        s_xo[tid + 1] = Od[RC - 1];
      }
      const int i_lo = max(max(0, d - (m - 1)), (d - rho + 1) >> 1), i_hi = min(min(m - 1, d), (d + rho) >> 1);
      cells += (unsigned long long)max(0, i_hi - i_lo + 1);
      __syncthreads();
    };
    const int fast_begin = rho + 2, fast_pairs = max(0, m - 2 - rho);
    int d = 0;
    for (; d < min(fast_begin, last + 1) && !abandoned; d++) {
      if ((d & 7) == 0 && d > 0 && hopeless(d)) abandoned = true;
      else edge_step(d);
    }
    if (!abandoned && d == fast_begin && fast_pairs > 0) {
      // Interior: two diagonals per iteration, operands carried in registers (even -> odd: every i grows by one;
      // odd -> even: every j grows by one), no index tests; cells beyond the band hold values >= INF, which a band
      // cell's min never selects.  Two block barriers per pair.
      const int i_base = (d + u0 - rho) >> 1;
      double av[RC], bv[RC];
#pragma unroll
      for (int r = 0; r < RC; r++) {
        av[r] = A[min(i_base + r, m - 1)];
        bv[r] = B[min(max(d - 1 - i_base - r, 0), m - 1)];  // as of diagonal d - 1; shifted on entry
      }
      int ia = i_base + RC, jb = d - i_base;
      double a_new = A[min(ia, m - 1)], b_new = B[min(max(jb, 0), m - 1)];
      int p = 0;
      for (; p < fast_pairs; p++) {
        const bool probe = (p & 15) == 15;
        if (probe) {  // warp minima now, decision after this pair's first barrier (no barrier of its own)
          double mn = kDtwInf;
#pragma unroll
          for (int r = 0; r < RC; r++) mn = umin_pos(mn, umin_pos(Ev[r], Od[r]));
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) mn = umin_pos(mn, __shfl_xor_sync(kFullMask, mn, o));
          if (lane == 0) s_red[warp] = mn;
        }
        // even diagonal
        const double left = s_xo[tid];
#pragma unroll
        for (int r = RC - 1; r > 0; r--) bv[r] = bv[r - 1];
        bv[0] = b_new;
        jb++;
        b_new = B[min(max(jb, 0), m - 1)];
#pragma unroll
        for (int r = 0; r < RC; r++) {
          const double own = umin_pos(Od[r], Ev[r]);
          const double x = (r == 0) ? left : Od[r - 1];
          Ev[r] = xadd(umin_pos(x, own), xsqdist(av[r], bv[r]));
        }
        s_xe[tid] = Ev[0];
        __syncthreads();
        if (probe) {
          double mn = s_red[0];
#pragma unroll
          for (int k = 1; k < kCoopThreads / 32; k++) mn = umin_pos(mn, s_red[k]);
          const int dc = d + 2 * p;  // the minima were taken over diagonals dc - 1 and dc - 2
          const int imax = min(m - 1, (dc - 1 + rho) >> 1);
          if (mn + CBS[(imax + 1 + 7) >> 3] > P.eps2_hi) {
            abandoned = true;
            break;
          }
        }
        // odd diagonal
        const double right = s_xe[tid + 1];
#pragma unroll
        for (int r = 0; r < RC - 1; r++) av[r] = av[r + 1];
        av[RC - 1] = a_new;
        ia++;
        a_new = A[min(ia, m - 1)];
#pragma unroll
        for (int r = 0; r < RC; r++) {
          const double own = umin_pos(Ev[r], Od[r]);
          const double y = (r == RC - 1) ? right : Ev[r + 1];
          const double v = xadd(umin_pos(y, own), xsqdist(av[r], bv[r]));
          Od[r] = ((band_odd >> r) & 1u) ? v : kDtwInf;
        }
        s_xo[tid + 1] = Od[RC - 1];
        __syncthreads();
      }
      cells += (unsigned long long)p * (unsigned long long)(2 * rho + 1);
      d += 2 * p;
    }
    for (; d <= last && !abandoned; d++) {
      if ((d & 15) == 0 && hopeless(d)) abandoned = true;
      else edge_step(d);
    }
    if (tid == 0) {
      atomicAdd(P.n_cells, cells);
      if (abandoned) atomicAdd(P.n_abandoned, 1ULL);
    }
    if (!abandoned && tid == tgt_pair / RC) {
      const int r = tgt_pair % RC;
      double res = kDtwInf;
#pragma unroll
      for (int k = 0; k < RC; k++)
        if (k == r) res = (tgt_u & 1) ? Od[k] : Ev[k];
      if (res <= P.eps2) P.sink.emit(off, xsqrt(res));
    }
  }
}

}  // namespace kvm
