// stream_kernels.cuh — the cNSM statistics pass as an HBM stream with an exact sparse fix-up.
// Replaces the statistics half of K/NormQueryEngine.java:487-526 and K/NormQueryEngineDtw.java:512-603.
//
// The reference's window sums (ex, ex2) come from one sequential add/subtract chain per merged interval, so their
// rounding history depends on the chain's whole past.  Walking every chain exactly (cnsm_relay_kernel, round 1) is
// latency-bound on the dependent DADDs.  But only a vanishing fraction of the windows needs the chain's exact value:
//   * a window whose gate decision (alpha/beta) could flip within the chain's provable rounding drift, and
//   * a window that survives a cheap lower bound of its distance (its mean/std feed a reported distance).
// Everything else is decided from window sums computed any way we like, provided the decision keeps a guard band that
// covers (drift of the reference chain) + (error of our own sums).  So:
//
//   cnsm_stream_kernel   one CTA per tile of up to 33*NT consecutive window starts.  The tile's samples
//                        (nwin + m - 1 doubles) arrive in shared memory with TMA bulk copies (cp.async.bulk, one
//                        elected thread, completion on an mbarrier).  Threads form sums of 33-sample groups, a block
//                        scan turns them into a prefix, each thread takes the sums of its first window from two
//                        prefix entries and then slides over its 33 consecutive windows (stride 33 doubles between
//                        lanes: conflict-free LDS.64).  Per window: two compares on ex; only windows inside the outer
//                        mean band compute m*ex2 - ex^2 and compare it with the variance band; windows inside both
//                        run a 32-term lower bound of the distance (ED: the 32 largest |zQ| terms; DTW: LB_KimFL +
//                        32 envelope terms) from the samples already in shared memory.  A window is FLAGGED when it
//                        is ambiguous (inside a guard band of any gate threshold) or survives the lower bound; a
//                        window that is certainly inside the gate and certainly not an answer is only counted.
//   chain_rewalk_kernel  one warp per flagged chain: the reference's chain, exactly (xadd/xmul, its operation
//                        order), from the chain's first sample to its last flagged window; emits (offset, ex, ex2)
//                        for the flagged windows.
//   the exact stages     (cnsm_ed_exact_kernel<true>, cnsm_dtw_lb_list_kernel) recompute mean/std/gate with the
//                        reference's arithmetic from those sums and go on as before.
//
// Every reported number (answers, distances, n_gate_pass) therefore still comes from the reference's arithmetic; the
// stream only decides which windows can be skipped, and the guard (host: stream_guard()) is a worst-case bound.
#pragma once
#include "cnsm_kernels.cuh"

namespace kvm {

constexpr int kGroup = 33;        // windows per thread = samples per group (odd: lane stride 33 doubles is conflict-free)
constexpr int kScreenTerms = 32;  // terms of the in-stream lower bound

struct StreamTile {
  int32_t s0;     // local index of the tile's first sample (= its first window start)
  int32_t nwin;   // window starts in the tile (<= 33 * NT)
  int32_t chain;  // live-chain index of the first window
  int32_t v0;     // ordinal (over all live chains, list order) of the first window
};

struct StreamGate {  // on ex (= m * mean) and t2 = m*ex2 - ex^2 (= m^2 * variance); out = may pass, in = certainly passes
  double e_lo_out, e_lo_in, e_hi_in, e_hi_out;
  double v_lo_out, v_lo_in, v_hi_in, v_hi_out;
};

// Guard bands as a function of A = max |sample| over everything a tile's chains have summed (stream_guard() on the
// host explains every term).  Evaluated per tile on the device with the tile's own A, so that a series whose
// amplitude varies by orders of magnitude (random walks) keeps tight bands where the values are small.
struct GuardCoef {
  double cd1, cd2;          // |ex_stream - ex_chain| <= cd1*A, |ex2_stream - ex2_chain| <= cd2*A^2
  double dm, abs_mean_beta; // m, |meanQ| + |beta|
  double c_lo, c_hi;        // m*(meanQ -+ beta)
  double lo, hi;            // variance band (stdQ/alpha)^2, (stdQ*alpha)^2
  double std_lo, xm, sqrt_terms, eps_abs;
};

__host__ __device__ inline void stream_guard_eval(const GuardCoef& C, double A, StreamGate* G, double* thr) {
  const double u = 1.1102230246251565e-16;  // 2^-53
  const double dm = C.dm;
  const double d1 = C.cd1 * A, d2 = C.cd2 * A * A;
  const double Mb = C.abs_mean_beta + 2.0 * d1 / dm + 1e-300;  // |mean| of a window in the outer mean band
  const double g1 = d1 + dm * 16.0 * u * Mb;
  G->e_lo_out = C.c_lo - g1;
  G->e_lo_in = C.c_lo + g1;
  G->e_hi_in = C.c_hi - g1;
  G->e_hi_out = C.c_hi + g1;
  const double rv = 16.0 * u * (C.hi + 2.0 * Mb * Mb);
  const double g2 = dm * d2 + 2.0 * dm * Mb * d1 + d1 * d1 + dm * dm * (rv + 64.0 * u * C.hi);
  const double lo2 = dm * dm * C.lo, hi2 = dm * dm * C.hi;
  G->v_lo_out = lo2 - g2;
  G->v_lo_in = lo2 + g2;
  G->v_hi_in = hi2 - g2;
  G->v_hi_out = hi2 + g2;
  const double g_mean = g1 / dm + 4.0 * u * Mb;
  const double g_var = g2 / (dm * dm);
  double t = 1.0 / 0.0;
  if (C.std_lo > 0.0 && g_var < 0.01 * C.std_lo * C.std_lo) {
    const double g_std = g_var / C.std_lo;
    const double dx = (g_mean + C.xm * g_std) / (0.99 * C.std_lo) + 16.0 * u * C.xm;
    const double D = C.sqrt_terms * dx;
    t = (C.eps_abs + D) * (C.eps_abs + D) * (1.0 + 1e-12) + 1e-300;
  }
  *thr = t;
}

constexpr int kBmaxBlock = 1024;  // samples per entry of the block-maximum table

struct StreamParams {
  const double* __restrict__ T;
  const StreamTile* __restrict__ tiles;
  const int32_t* __restrict__ cbegin;  // live chains: local index of the first sample
  const int32_t* __restrict__ ncand;   //              window starts
  int n_chains;
  int m;
  double dm, inv_m, inv_m2;
  GuardCoef C;
  const double* __restrict__ bmax;  // max |sample| per kBmaxBlock samples of the shard
  int n_bmax;
  int l_max;                        // samples of the longest chain
  // in-stream lower bound: ED: idx = order[k], a = zq (sorted order); DTW: idx = sampled positions, a = upper, b = lower envelope
  const int32_t* __restrict__ scr_idx;
  const double* __restrict__ scr_a;
  const double* __restrict__ scr_b;
  // the rest of the bound, for table survivors.  ED: q_full = zQ in |z|-descending order, order_full = its
  // permutation; DTW: q_full = zQ in natural order, uq_full / lq_full = its envelope
  const double* __restrict__ q_full;
  const int32_t* __restrict__ order_full;
  const double* __restrict__ uq_full;
  const double* __restrict__ lq_full;
  int n_screen;
  int force_all;     // debug: flag every window inside the outer gate (the exact stages then decide everything)
  // outputs
  unsigned* need_bits;       // bit v = window ordinal v is flagged (all zero between calls)
  int32_t* chain_last;       // per live chain: last flagged window index, -1 = none (all -1 between calls)
  int32_t* flagged;          // compact list of flagged chains
  unsigned long long* n_flagged;
  unsigned long long* gate_pass;  // windows certainly inside the gate and not flagged
  unsigned long long* n_need;     // diagnostics: flagged windows
};

// bmax[i] = max |sample| over samples [i*1024, (i+1)*1024) of the shard, and the global maximum, as bit patterns of
// non-negative doubles (integer order = IEEE order; a NaN sorts above +inf, so a series with NaNs reports a
// non-finite maximum and the host keeps such series off the stream path).  One warp per block.
__global__ void __launch_bounds__(256) blockmax_kernel(const double* __restrict__ x, long long n, double* __restrict__ bmax,
                                                       long long n_blocks, unsigned long long* out) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  unsigned long long all = 0ULL;
  for (long long b = warp; b < n_blocks; b += n_warps) {
    unsigned long long mx = 0ULL;
    const long long base = b * kBmaxBlock;
#pragma unroll 8
    for (int i = lane; i < kBmaxBlock; i += 32)
      if (base + i < n) mx = max(mx, (unsigned long long)__double_as_longlong(fabs(x[base + i])));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(kFullMask, mx, o));
    if (lane == 0) bmax[b] = __longlong_as_double((long long)mx);
    all = max(all, mx);
  }
  if (lane == 0 && all) atomicMax(out, all);
}

// ---- TMA bulk copy (global -> shared, 1-D) ----------------------------------------------------------------------
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

constexpr size_t stream_smem_bytes(int nt, int m) {
  const size_t n_xs = ((size_t)kGroup * nt + m + 42) & ~size_t(1);
  const size_t ng = ((size_t)kGroup * nt + m + kGroup) / 11 + 4;
  return 16 + sizeof(double) * (n_xs + 2 * ng + 2 * kScreenTerms) + sizeof(int32_t) * kScreenTerms + 16;
}

// Flag the window that starts at local sample s (ordinal v); p = a live chain at or before its chain (rare path;
// arguments by value so that the kernel's parameter block never needs an address).
__device__ __noinline__ void stream_flag(const int32_t* __restrict__ cbegin, const int32_t* __restrict__ ncand,
                                         unsigned* need_bits, int32_t* chain_last, int32_t* flagged,
                                         unsigned long long* n_flagged, unsigned long long* n_need, int s, int p,
                                         unsigned v) {
  while (s >= __ldg(cbegin + p) + __ldg(ncand + p)) p++;  // chains of one segment are consecutive
  atomicOr(need_bits + (v >> 5), 1u << (v & 31));
  const int old = atomicMax(chain_last + p, s - __ldg(cbegin + p));
  if (old < 0) {
    const unsigned long long slot = atomicAdd(n_flagged, 1ULL);
    flagged[slot] = p;
  }
  atomicAdd(n_need, 1ULL);
}

// Shared-memory queue of the windows that survived the table tier of the in-stream lower bound; the CTA's warps finish
// their bound cooperatively after the slide phase.  It reuses the group-prefix arrays (dead by then).
struct StreamQueue {
  int* count;
  int cap;
  int* w;       // window index in the tile; bit 30 = certainly inside the gate
  double* ex;
  double* ex2;
};

// The rare path of the stream: a window inside the outer mean band.  Returns 1 when the window certainly passes the
// gate and certainly is no answer (it is only counted).  Otherwise the window is either flagged for the exact re-walk
// (ambiguous gate and already pruned) or queued for the second tier of the bound.
// Table tier: ED: the 32 largest-|zQ| terms of the distance; DTW: LB_KimFL (K/utils/DtwUtils.java:149-189) and 32
// evenly spread terms of LB_Keogh on the query envelope.  `wv` = the window's first sample in shared memory.
template <int kMode>
__device__ __forceinline__ unsigned stream_window(const StreamParams& P, const StreamTile& tile, const double* __restrict__ wv,
                                                  int w, double ex, double ex2, const int32_t* __restrict__ scr_i,
                                                  const double* __restrict__ scr_a, const double* __restrict__ scr_b,
                                                  const StreamQueue& Q, const StreamGate& G, const double thr) {
  if (!(ex >= G.e_lo_out && ex <= G.e_hi_out)) return 0u;
  const double t2 = __fma_rn(P.dm, ex2, -__dmul_rn(ex, ex));
  if (!(t2 >= G.v_lo_out && t2 <= G.v_hi_out)) return 0u;
  const bool sure = ex >= G.e_lo_in && ex <= G.e_hi_in && t2 >= G.v_lo_in && t2 <= G.v_hi_in;
  bool survive = true;
  const int m = P.m;
  if (t2 > 0.0 && !P.force_all) {
    const double mean = ex * P.inv_m;
    const double rstd = rsqrt(t2 * P.inv_m2);
    if (kMode == 1 && m >= 6) {
      const double* __restrict__ q = P.q_full;
      const double x0 = (wv[0] - mean) * rstd, x1 = (wv[1] - mean) * rstd, x2 = (wv[2] - mean) * rstd;
      const double y0 = (wv[m - 1] - mean) * rstd, y1 = (wv[m - 2] - mean) * rstd, y2 = (wv[m - 3] - mean) * rstd;
      const double q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
      const double p0 = __ldg(q + m - 1), p1 = __ldg(q + m - 2), p2 = __ldg(q + m - 3);
      auto sq = [](double a, double b) { const double d = a - b; return d * d; };
      double kim = sq(x0, q0) + sq(y0, p0);
      kim += fmin(fmin(sq(x1, q0), sq(x0, q1)), sq(x1, q1));
      kim += fmin(fmin(sq(y1, p0), sq(y0, p1)), sq(y1, p1));
      kim += fmin(fmin(fmin(sq(x0, q2), sq(x1, q2)), sq(x2, q2)), fmin(sq(x2, q1), sq(x2, q0)));
      kim += fmin(fmin(fmin(sq(y0, p2), sq(y1, p2)), sq(y2, p2)), fmin(sq(y2, p1), sq(y2, p0)));
      survive = kim <= thr;
    }
    const int n_scr = P.n_screen;
    double dist = 0.0;
    for (int kk = 0; kk < n_scr && survive; kk += 4) {
#pragma unroll
      for (int u = 0; u < 4; u++) {
        if (kk + u < n_scr) {
          const double x = (wv[scr_i[kk + u]] - mean) * rstd;
          double d;
          if (kMode == 0) {
            d = x - scr_a[kk + u];
          } else {
            const double up = scr_a[kk + u], lo = scr_b[kk + u];
            d = (x > up) ? (x - up) : ((x < lo) ? (x - lo) : 0.0);
          }
          dist = __fma_rn(d, d, dist);
        }
      }
      survive = dist <= thr;
    }
    if (survive && n_scr < m) {  // second tier: queue for the warps (a full queue falls through to the flag)
      const int slot = atomicAdd(Q.count, 1);
      if (slot < Q.cap) {
        Q.w[slot] = w | (sure ? (1 << 30) : 0);
        Q.ex[slot] = ex;
        Q.ex2[slot] = ex2;
        return 0u;
      }
    }
  }
  if (!survive && sure) return 1u;
  stream_flag(P.cbegin, P.ncand, P.need_bits, P.chain_last, P.flagged, P.n_flagged, P.n_need, tile.s0 + w, tile.chain,
              (unsigned)(tile.v0 + w));
  return 0u;
}

// Second tier, one warp per queued window: the whole bound with the lanes striding over its terms — ED: the full
// |zQ|-ordered sum (a permutation of the distance's terms); DTW: LB_Keogh on the query envelope over all m positions.
// Returns (lane 0) 1 when the window is certainly a gate pass and certainly no answer.
template <int kMode>
__device__ __forceinline__ unsigned stream_tier2(const StreamParams& P, const StreamTile& tile, const double* __restrict__ xs,
                                                 int wq, double ex, double ex2, int lane, const double thr) {
  const int w = wq & ((1 << 30) - 1);
  const bool sure = (wq >> 30) & 1;
  const int m = P.m;
  const double t2 = __fma_rn(P.dm, ex2, -__dmul_rn(ex, ex));
  const double mean = ex * P.inv_m;
  const double rstd = rsqrt(t2 * P.inv_m2);
  const double* __restrict__ wv = xs + w;
  double part = 0.0;
  bool survive = true;
  for (int k0 = 0; k0 < m && survive; k0 += 128) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int k = k0 + u * 32 + lane;
      if (k < m) {
        double d;
        if (kMode == 0) {
          d = (wv[__ldg(P.order_full + k)] - mean) * rstd - __ldg(P.q_full + k);
        } else {
          const double x = (wv[k] - mean) * rstd;
          const double up = __ldg(P.uq_full + k), lo = __ldg(P.lq_full + k);
          d = (x > up) ? (x - up) : ((x < lo) ? (x - lo) : 0.0);
        }
        part = __fma_rn(d, d, part);
      }
    }
    double tot = part;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFullMask, tot, o);
    survive = tot <= thr;
  }
  if (lane != 0) return 0u;
  if (!survive && sure) return 1u;
  stream_flag(P.cbegin, P.ncand, P.need_bits, P.chain_last, P.flagged, P.n_flagged, P.n_need, tile.s0 + w, tile.chain,
              (unsigned)(tile.v0 + w));
  return 0u;
}

#ifdef KVM_STREAM_PROF
__device__ unsigned long long g_stream_prof[16];  // thread 0 of every CTA: cycles per phase, CTA count
__device__ __forceinline__ long long stream_clock() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
#define STREAM_T(var) const long long var = stream_clock()
#else
#define STREAM_T(var)
#endif

constexpr int kSub = 11;            // a thread's 33 windows are three independent chains of 11 (instruction-level parallelism)
constexpr int kSubs = kGroup / kSub;

// kMode 0: cNSM-ED, 1: cNSM-DTW (they differ in the in-stream lower bound only)
template <int NT, int kMode>
__global__ void __launch_bounds__(NT, NT <= 192 ? 3 : (NT <= 256 ? 2 : 1)) cnsm_stream_kernel(StreamParams P) {
  extern __shared__ __align__(16) unsigned char stream_smem[];
  __shared__ double s_w1[NT / 32], s_w2[NT / 32];
  __shared__ int s_qcount;
  __shared__ StreamGate s_gate;
  __shared__ double s_thr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  STREAM_T(c0);
  const StreamTile tile = P.tiles[blockIdx.x];
  const int m = P.m;
  const int nwin = tile.nwin;
  const int ns = nwin + m - 1;                 // samples the tile's windows cover
  const int s0a = tile.s0 & ~1;                // 16-byte aligned copy start
  const int lead = tile.s0 - s0a;
  const int nload = (lead + ns + 1) & ~1;      // even number of doubles
  const int ng = (ns + kSub - 1) / kSub;       // 11-sample groups
  const size_t n_xs = ((size_t)kGroup * NT + m + 42) & ~size_t(1);
  const size_t ng_cap = ((size_t)kGroup * NT + m + kGroup) / kSub + 4;
  double* xs_raw = reinterpret_cast<double*>(stream_smem + 16);
  double* xs = xs_raw + lead;
  double* gp1 = xs_raw + n_xs;
  double* gp2 = gp1 + ng_cap;
  double* scr_a = gp2 + ng_cap;
  double* scr_b = scr_a + kScreenTerms;
  int32_t* scr_i = reinterpret_cast<int32_t*>(scr_b + kScreenTerms);
  const uint32_t bar = smem_u32(stream_smem);

  if (tid == 0) {
    s_qcount = 0;
    mbar_init(bar, 1);
    fence_mbar_init();
    const uint32_t bytes = (uint32_t)nload * 8u;
    mbar_expect_tx(bar, bytes);
    const unsigned char* src = reinterpret_cast<const unsigned char*>(P.T + s0a);
    const uint32_t dst = smem_u32(xs_raw);
    constexpr uint32_t kChunk = 16384;
    for (uint32_t o = 0; o < bytes; o += kChunk) tma_bulk_g2s(dst + o, src + o, min(kChunk, bytes - o), bar);
  }
  if (tid < P.n_screen) {
    scr_i[tid] = __ldg(P.scr_idx + tid);
    scr_a[tid] = __ldg(P.scr_a + tid);
    scr_b[tid] = __ldg(P.scr_b + tid);
  }
  if (warp == NT / 32 - 1) {
    // guard bands of this tile: A = max |sample| over every sample a chain can have summed before or inside one of the
    // tile's windows, i.e. from l_max samples before the tile to its end (block-maximum table, kBmaxBlock granularity)
    const int b_lo = max(0, tile.s0 - P.l_max) / kBmaxBlock;
    const int b_hi = min(P.n_bmax - 1, (tile.s0 + ns) / kBmaxBlock);
    unsigned long long mx = 0ULL;
    for (int b = b_lo + lane; b <= b_hi; b += 32) mx = max(mx, (unsigned long long)__double_as_longlong(__ldg(P.bmax + b)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(kFullMask, mx, o));
    if (lane == 0) {
      StreamGate g;
      double t;
      stream_guard_eval(P.C, __longlong_as_double((long long)mx), &g, &t);
      s_gate = g;
      s_thr = t;
    }
  }
  // zero what follows the copied samples (never part of a valid window; keeps the group prefix finite)
  for (int i = nload + tid; i < lead + ns + kGroup + 2; i += NT) xs_raw[i] = 0.0;
  __syncthreads();  // barrier initialised; screen table, guard bands and zero tail staged
  STREAM_T(c1);
  mbar_wait(bar, 0);
  STREAM_T(c2);

  // ---- sums of 11-sample groups (lane stride 11 doubles: conflict-free), three independent groups at a time
  for (int g0 = tid; g0 <= ng; g0 += kSubs * NT) {
    double a1[kSubs], a2[kSubs];
#pragma unroll
    for (int b = 0; b < kSubs; b++) a1[b] = a2[b] = 0.0;
#pragma unroll
    for (int j = 0; j < kSub; j++) {
#pragma unroll
      for (int b = 0; b < kSubs; b++) {
        const int g = g0 + b * NT;
        if (g < ng) {
          const double v = xs[g * kSub + j];
          a1[b] += v;
          a2[b] = __fma_rn(v, v, a2[b]);
        }
      }
    }
#pragma unroll
    for (int b = 0; b < kSubs; b++) {
      const int g = g0 + b * NT;
      if (g <= ng) {
        gp1[g] = a1[b];
        gp2[g] = a2[b];
      }
    }
  }
  __syncthreads();
  STREAM_T(c3);
  // ---- exclusive prefix over the groups, in place (blocked: thread t owns elements [t*E, (t+1)*E))
  {
    const int E = (ng + 1 + NT - 1) / NT;
    const int e0 = min(tid * E, ng + 1), e1 = min(e0 + E, ng + 1);
    double t1 = 0.0, t2 = 0.0;
    for (int e = e0; e < e1; e++) {
      t1 += gp1[e];
      t2 += gp2[e];
    }
    double i1 = t1, i2 = t2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u1 = __shfl_up_sync(kFullMask, i1, o), u2 = __shfl_up_sync(kFullMask, i2, o);
      if (lane >= o) {
        i1 += u1;
        i2 += u2;
      }
    }
    if (lane == 31) {
      s_w1[warp] = i1;
      s_w2[warp] = i2;
    }
    __syncthreads();
    double b1 = i1 - t1, b2 = i2 - t2;  // exclusive within the warp
    for (int w = 0; w < warp; w++) {
      b1 += s_w1[w];
      b2 += s_w2[w];
    }
    for (int e = e0; e < e1; e++) {
      const double v1 = gp1[e], v2 = gp2[e];
      gp1[e] = b1;
      gp2[e] = b2;
      b1 += v1;
      b2 += v2;
    }
  }
  __syncthreads();
  STREAM_T(c4);

  // ---- this thread's 33 consecutive windows = three chains of 11.  A chain's first window takes its sums from two
  // prefix entries (+ m % 11 samples); the sums then slide branch-free (one dependent DADD / DFMA per window and chain)
  // while a bit records the windows inside the outer mean band; the rare chains with a bit set are replayed (same
  // operations, same values) window by window through stream_window().
  const int w0 = tid * kGroup;
  unsigned my_gate = 0;
  const int qa = m / kSub, qb = m - qa * kSub;
  double ex[kSubs], ex2[kSubs];
#pragma unroll
  for (int b = 0; b < kSubs; b++) ex[b] = ex2[b] = 0.0;
  const bool active = w0 < nwin;
  if (active) {
#pragma unroll
    for (int b = 0; b < kSubs; b++) {
      const int g = kSubs * tid + b;
      ex[b] = gp1[g + qa] - gp1[g];
      ex2[b] = gp2[g + qa] - gp2[g];
    }
    for (int j = 0; j < qb; j++) {
#pragma unroll
      for (int b = 0; b < kSubs; b++) {
        const double v = xs[(kSubs * tid + b + qa) * kSub + j];
        ex[b] += v;
        ex2[b] = __fma_rn(v, v, ex2[b]);
      }
    }
  }
  __syncthreads();  // every thread has read its prefix entries: the arrays now hold the second-tier queue
  STREAM_T(c5);
  StreamQueue Q;
  Q.count = &s_qcount;
  Q.ex = gp1;
  Q.ex2 = gp2;
  Q.cap = (int)((ng_cap * 2) / 3);
  Q.w = reinterpret_cast<int*>(gp2 + Q.cap);  // the last third of gp2 holds the 4-byte indices
  if (active) {
    const int nw = min(kGroup, nwin - w0);
    const double* __restrict__ xo = xs + w0;
    const double* __restrict__ xi = xs + w0 + m;
    // Outer mean band as an integer test (DSETP issues at a fraction of the DADD rate on sm_100): |ex - mid| <= rad,
    // compared on the high words — monotone, so it can only admit more windows than the exact test, which
    // stream_window() repeats in double precision.
    const double e_mid = 0.5 * (s_gate.e_lo_out + s_gate.e_hi_out);
    const double e_rad = 0.5 * (s_gate.e_hi_out - s_gate.e_lo_out) * (1.0 + 1e-15) + 1e-300;
    const int r_hi = (e_rad >= 0.0) ? __double2hiint(e_rad) : -1;
    double ex_s[kSubs], ex2_s[kSubs];
    unsigned mask[kSubs];
#pragma unroll
    for (int b = 0; b < kSubs; b++) {
      ex_s[b] = ex[b];
      ex2_s[b] = ex2[b];
      mask[b] = 0u;
    }
#pragma unroll
    for (int j = 0; j < kSub; j++) {
#pragma unroll
      for (int b = 0; b < kSubs; b++) {
        const double a = xi[b * kSub + j], o = xo[b * kSub + j];
        mask[b] |= ((__double2hiint(ex[b] - e_mid) & 0x7fffffff) <= r_hi) ? (1u << j) : 0u;
        const double dl = a - o, sm = a + o;
        ex[b] += dl;
        ex2[b] = __fma_rn(dl, sm, ex2[b]);
      }
    }
#pragma unroll
    for (int b = 0; b < kSubs; b++) {
      const int left = nw - b * kSub;  // valid windows of this chain
      unsigned mk = mask[b];
      if (left < kSub) mk &= (left > 0) ? ((1u << left) - 1u) : 0u;
      if (mk) {
        double rex = ex_s[b], rex2 = ex2_s[b];
        const int kb = b * kSub;
        for (int j = 0; (mk >> j) != 0u; j++) {
          if ((mk >> j) & 1u)
            my_gate += stream_window<kMode>(P, tile, xo + kb + j, w0 + kb + j, rex, rex2, scr_i, scr_a, scr_b, Q, s_gate, s_thr);
          const double av = xi[kb + j], ov = xo[kb + j];
          const double dl = av - ov, sm = av + ov;
          rex += dl;
          rex2 = __fma_rn(dl, sm, rex2);
        }
      }
    }
  }
  // ---- second tier of the bound for the queued windows, one warp each
  STREAM_T(c6);
  __syncthreads();
  STREAM_T(c7);
  {
    const int nq = min(s_qcount, Q.cap);
    for (int e = warp; e < nq; e += NT / 32) my_gate += stream_tier2<kMode>(P, tile, xs, Q.w[e], Q.ex[e], Q.ex2[e], lane, s_thr);
  }
  // ---- certain gate passes
  unsigned tot = my_gate;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFullMask, tot, o);
  if (lane == 0 && tot) atomicAdd(P.gate_pass, (unsigned long long)tot);
#ifdef KVM_STREAM_PROF
  if (tid == 0) {
    const long long c8 = stream_clock();
    const long long d[9] = {c1 - c0, c2 - c1, c3 - c2, c4 - c3, c5 - c4, c6 - c5, c7 - c6, c8 - c7, 1};
    for (int i = 0; i < 9; i++) atomicAdd(&g_stream_prof[i], (unsigned long long)d[i]);
  }
#endif
}

// ---------------------------------------------------------------------------------------------------------------
struct RewalkParams {
  const double* __restrict__ T;
  const int32_t* __restrict__ cbegin;
  const int32_t* __restrict__ vbase;  // live chains: ordinal of the chain's first window
  int m;
  int32_t first_global;
  unsigned* need_bits;
  int32_t* chain_last;
  const int32_t* __restrict__ flagged;
  const unsigned long long* __restrict__ n_flagged;
  XList out;
};

__device__ __forceinline__ void cp_async8_zfill(uint32_t dst, const void* src, bool valid) {
  const int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}

constexpr int kRewalkStages = 6;  // 32-position tiles in flight per warp

// One warp per flagged chain.  All lanes run the same recurrence (the chain is sequential; what the warp buys is
// coalesced, deeply prefetched sample rows and nothing but LDS + DADD/DMUL on the dependent path):
//   K/NormQueryEngine.java:498-499  ex += d; ex2 += d*d        (every sample)
//   K/NormQueryEngine.java:523-524  ex -= T[j]; ex2 -= T[j]^2  (after each complete window)
// Out-of-range slots of a 32-position tile are zero-filled: adding / subtracting +0.0 is the identity here (the sums
// start at +0.0 and x + y is -0.0 only when both are).
__global__ void __launch_bounds__(32) chain_rewalk_kernel(RewalkParams P) {
  __shared__ __align__(16) double s_in[kRewalkStages][32];
  __shared__ __align__(16) double s_out[kRewalkStages][32];
  __shared__ double s_ex[32], s_ex2[32];
  const int lane = threadIdx.x;
  const int m = P.m;
  const unsigned long long n = *P.n_flagged;
  for (unsigned long long ci = blockIdx.x; ci < n; ci += gridDim.x) {
    const int p = P.flagged[ci];
    const int cb = P.cbegin[p];
    const int last = P.chain_last[p];
    const int vb = P.vbase[p];
    __syncwarp();
    if (lane == 0) P.chain_last[p] = -1;  // ready for the next call
    const int qend = (m - 1) + last;       // last sample position (0-based within the chain) the walk consumes
    const int n_tiles = qend / 32 + 1;
    const double* __restrict__ base = P.T + cb;
    auto issue = [&](int t) {
      if (t < n_tiles) {
        const int q = t * 32 + lane;
        const int st = t % kRewalkStages;
        cp_async8_zfill(smem_u32(&s_in[st][lane]), base + min(q, qend), q <= qend);
        const int j = q - (m - 1);
        cp_async8_zfill(smem_u32(&s_out[st][lane]), base + max(j, 0), j >= 0 && q <= qend);
      }
      cp_async_commit();
    };
#pragma unroll
    for (int t = 0; t < kRewalkStages - 1; t++) issue(t);
    // flag words of a tile: windows jA..jB of the chain, ordinals vb+jA.. (loaded one tile ahead of their use)
    auto flag_words = [&](int t, unsigned& wlo, unsigned& whi) {
      const int j0 = t * 32 - (m - 1);
      const int jA = max(j0, 0), jB = min(j0 + 31, last);
      wlo = whi = 0u;
      if (t < n_tiles && jA <= jB) {
        const unsigned vA = (unsigned)(vb + jA);
        wlo = P.need_bits[vA >> 5];
        whi = P.need_bits[(vA >> 5) + 1];
      }
    };
    double ex = 0.0, ex2 = 0.0;
    unsigned nlo, nhi;
    flag_words(0, nlo, nhi);
    for (int t = 0; t < n_tiles; t++) {
      issue(t + kRewalkStages - 1);
      const unsigned wlo = nlo, whi = nhi;
      flag_words(t + 1, nlo, nhi);
      cp_async_wait<kRewalkStages - 1>();
      __syncwarp();
      const int st = t % kRewalkStages;
      const int q0 = t * 32;
      const int j0 = q0 - (m - 1);
      const int jA = max(j0, 0), jB = min(j0 + 31, last);
      unsigned mask = 0;
      unsigned vA = 0;
      if (jA <= jB) {
        vA = (unsigned)(vb + jA);
        mask = __funnelshift_r(wlo, whi, vA & 31);
        const int cnt = jB - jA + 1;
        if (cnt < 32) mask &= (1u << cnt) - 1u;
      }
      const double2* __restrict__ pin = reinterpret_cast<const double2*>(s_in[st]);
      const double2* __restrict__ pout = reinterpret_cast<const double2*>(s_out[st]);
      if (j0 + 31 < 0) {  // warm-up tile: no window completes, nothing leaves
#pragma unroll
        for (int h = 0; h < 16; h++) {
          const double2 a = pin[h];
          ex = xadd(ex, a.x);
          ex2 = xadd(ex2, xmul(a.x, a.x));
          ex = xadd(ex, a.y);
          ex2 = xadd(ex2, xmul(a.y, a.y));
        }
      } else if (mask == 0) {
#pragma unroll
        for (int h = 0; h < 16; h++) {
          const double2 a = pin[h], o = pout[h];
          ex = xadd(ex, a.x);
          ex2 = xadd(ex2, xmul(a.x, a.x));
          ex = xsub(ex, o.x);
          ex2 = xsub(ex2, xmul(o.x, o.x));
          ex = xadd(ex, a.y);
          ex2 = xadd(ex2, xmul(a.y, a.y));
          ex = xsub(ex, o.y);
          ex2 = xsub(ex2, xmul(o.y, o.y));
        }
      } else {
        // a tile with flagged windows: the same walk, every post-add state parked in shared memory; afterwards lane pp
        // emits position pp if its window is flagged (one counter update per warp, coalesced stores)
#pragma unroll
        for (int h = 0; h < 16; h++) {
          const double2 a = pin[h], o = pout[h];
          ex = xadd(ex, a.x);
          ex2 = xadd(ex2, xmul(a.x, a.x));
          s_ex[2 * h] = ex;
          s_ex2[2 * h] = ex2;
          ex = xsub(ex, o.x);
          ex2 = xsub(ex2, xmul(o.x, o.x));
          ex = xadd(ex, a.y);
          ex2 = xadd(ex2, xmul(a.y, a.y));
          s_ex[2 * h + 1] = ex;
          s_ex2[2 * h + 1] = ex2;
          ex = xsub(ex, o.y);
          ex2 = xsub(ex2, xmul(o.y, o.y));
        }
        __syncwarp();
        const int j = j0 + lane;
        const bool hit = j >= jA && j <= jB && ((mask >> (j - jA)) & 1u);
        const unsigned bal = __ballot_sync(kFullMask, hit);
        unsigned long long slot0 = 0;
        if (lane == 0) slot0 = atomicAdd(P.out.count, (unsigned long long)__popc(bal));
        slot0 = __shfl_sync(kFullMask, slot0, 0);
        if (hit) {
          const unsigned long long slot = slot0 + __popc(bal & ((1u << lane) - 1u));
          if ((long long)slot < P.out.cap) {
            P.out.off[slot] = P.first_global + cb + j;
            P.out.ex[slot] = s_ex[lane];
            P.out.ex2[slot] = s_ex2[lane];
          }
        }
        if (lane == 0) {  // consume the bits
          const unsigned sh = vA & 31;
          atomicAnd(P.need_bits + (vA >> 5), ~(mask << sh));
          if (sh && (mask >> (32 - sh))) atomicAnd(P.need_bits + (vA >> 5) + 1, ~(mask >> (32 - sh)));
        }
      }
      __syncwarp();
    }
    cp_async_wait<0>();
  }
}

}  // namespace kvm
