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
  if (m >= 6) {  // LB_KimFL, K/utils/DtwUtils.java:149-189 (all five stages, no early return)
    const double x0 = ((w[0] - mean) * rstd), x1 = ((w[1] - mean) * rstd), x2 = ((w[2] - mean) * rstd);
    const double y0 = ((w[m - 1] - mean) * rstd), y1 = ((w[m - 2] - mean) * rstd),
                 y2 = ((w[m - 3] - mean) * rstd);
    const double q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
    const double p0 = __ldg(q + m - 1), p1 = __ldg(q + m - 2), p2 = __ldg(q + m - 3);
    lb = fsq(x0, q0) + fsq(y0, p0);
    lb += fmin(fmin(fsq(x1, q0), fsq(x0, q1)), fsq(x1, q1));
    lb += fmin(fmin(fsq(y1, p0), fsq(y0, p1)), fsq(y1, p1));
    lb += fmin(fmin(fmin(fsq(x0, q2), fsq(x1, q2)), fsq(x2, q2)), fmin(fsq(x2, q1), fsq(x2, q0)));
    lb += fmin(fmin(fmin(fsq(y0, p2), fsq(y1, p2)), fsq(y2, p2)), fmin(fsq(y2, p1), fsq(y2, p0)));
    if (!(lb <= Q.eps2_hi)) return false;
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

__device__ __forceinline__ void cand_append(const CandList& L, int32_t off, double mean, double stdv) {
  const unsigned long long slot = atomicAdd(L.count, 1ULL);
  if ((long long)slot < L.cap) {
    L.off[slot] = off;
    L.mean[slot] = mean;
    L.stdv[slot] = stdv;
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
  for (int c = c0 + (int)threadIdx.x; c < min(c0 + kEdTile, ncand); c += kEdThreads) {
    const int start = cbegin + c;
    if (lb_cascade(P.T + start, P.Q, 1.0, 0.0)) cand_append(P.out, P.first_global + start, 0.0, 1.0);
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
      if ((d & 15) == 0 && d > 0) {
        double mn = kDtwInf;
#pragma unroll
        for (int r = 0; r < R; r++) mn = umin_pos(mn, umin_pos(Ev[r], Od[r]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = umin_pos(mn, __shfl_xor_sync(kFullMask, mn, o));
        const int imax = min(m - 1, (d - 1 + rho) >> 1);
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
  double* B = coop_smem;           // query
  double* A = coop_smem + m;       // normalised window
  double* CBS = A + m;             // cumulative LB_Keogh remainder sampled every 8th position
  for (int k = tid; k < m; k += kCoopThreads) B[k] = P.q[k];
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
      const int imax = min(m - 1, (d - 1 + rho) >> 1);
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
      if ((d & 15) == 0 && d > 0 && hopeless(d)) abandoned = true;
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
