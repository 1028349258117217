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
#include <type_traits>
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
  // the same bands in centred form, |x - mid| <= rad, for the per-candidate tests: one DADD and an integer compare of
  // the (non-negative) bit patterns instead of two DSETPs, which issue at about a fifth of the DADD rate on sm_100.
  // rad_out is rounded outwards and rad_in inwards by more than the rounding of x - mid, so the centred tests admit
  // every window the interval tests admit (out) and call a window certain only when the interval tests do (in).
  double e_mid, e_rad_out, e_rad_in;
  double v_mid, v_rad_out, v_rad_in;
};

// Guard bands as a function of A = max |sample| over everything a tile's chains have summed (stream_guard() on the
// host explains every term).  Evaluated per tile on the device with the tile's own A, so that a series whose
// amplitude varies by orders of magnitude (random walks) keeps tight bands where the values are small.
struct GuardCoef {
  double cd1, cd2;          // |ex_stream - ex_chain| <= cd1*A, |ex2_stream - ex2_chain| <= cd2*A^2
  double dm, inv_dm;        // m, 1/m (the guard keeps 4u of slack for the rounded reciprocals)
  double abs_mean_beta;     // |meanQ| + |beta|
  double c_lo, c_hi;        // m*(meanQ -+ beta)
  double lo, hi;            // variance band (stdQ/alpha)^2, (stdQ*alpha)^2
  double std_lo, inv_std_lo;  // lowest std of a gate pass, and 1.02 / it
  double xm, sqrt_terms, eps_abs;
};

__host__ __device__ inline void stream_guard_eval(const GuardCoef& C, double A, StreamGate* G, double* thr) {
  const double u = 1.1102230246251565e-16;  // 2^-53
  const double dm = C.dm;
  const double d1 = C.cd1 * A, d2 = C.cd2 * A * A;
  const double Mb = C.abs_mean_beta + 2.0 * d1 * C.inv_dm * (1.0 + 8.0 * u) + 1e-300;  // |mean| of a window in the outer band
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
  {
    const double up = 1.0 + 16.0 * u, dn = 1.0 - 16.0 * u;
    G->e_mid = 0.5 * (G->e_lo_out + G->e_hi_out);
    G->e_rad_out = fmax(G->e_hi_out - G->e_mid, G->e_mid - G->e_lo_out) * up + 1e-300;
    G->e_rad_in = fmin(G->e_hi_in - G->e_mid, G->e_mid - G->e_lo_in) * dn - 1e-300;   // negative = no window is certain
    G->v_mid = 0.5 * (G->v_lo_out + G->v_hi_out);
    G->v_rad_out = fmax(G->v_hi_out - G->v_mid, G->v_mid - G->v_lo_out) * up + 1e-300;
    G->v_rad_in = fmin(G->v_hi_in - G->v_mid, G->v_mid - G->v_lo_in) * dn - 1e-300;
  }
  const double g_mean = g1 * C.inv_dm * (1.0 + 8.0 * u) + 4.0 * u * Mb;
  const double g_var = g2 * C.inv_dm * C.inv_dm * (1.0 + 8.0 * u);
  double t = 1.0 / 0.0;
  if (C.std_lo > 0.0 && g_var < 0.01 * C.std_lo * C.std_lo) {
    const double g_std = g_var * C.inv_std_lo;                        // >= g_var / std_lo
    const double dx = (g_mean + C.xm * g_std) * C.inv_std_lo + 16.0 * u * C.xm;  // >= (...) / (0.99 std_lo)
    const double D = C.sqrt_terms * dx;
    t = (C.eps_abs + D) * (C.eps_abs + D) * (1.0 + 1e-12) + 1e-300;
  }
  *thr = t;
}

constexpr int kBmaxBlock = 1024;  // samples per entry of the block-maximum table

// Geometry of the live chains (merged intervals with at least one window).  A regular grid — one run of adjacent
// window starts cut every `chunk` (an index-free scan) — needs no tables: everything follows from the chain index.
struct ChainTable {
  const int32_t* __restrict__ cbegin;  // local index of the chain's first sample
  const int32_t* __restrict__ ncand;   // window starts
  const int32_t* __restrict__ vbase;   // ordinal of the chain's first window over all live chains (list order)
  int n_chains;
  int regular;
  int32_t s_base, chunk, total_win;    // regular grids only
  __device__ __forceinline__ int begin(int p) const { return regular ? s_base + p * chunk : __ldg(cbegin + p); }
  __device__ __forceinline__ int count(int p) const { return regular ? min(chunk, total_win - p * chunk) : __ldg(ncand + p); }
  __device__ __forceinline__ int ordinal(int p) const { return regular ? p * chunk : __ldg(vbase + p); }
  __device__ __forceinline__ int chain_of(unsigned v) const {  // largest p with ordinal(p) <= v
    if (regular) return (int)(v / (unsigned)chunk);
    int lo = 0, hi = n_chains;
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if ((unsigned)__ldg(vbase + mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
  }
};

struct StreamParams {
  const double* __restrict__ T;
  const StreamTile* __restrict__ tiles;
  ChainTable chains;
  int m;
  double dm, inv_m, inv_m2;
  GuardCoef C;
  const double* __restrict__ bmax;  // max |sample| per kBmaxBlock samples of the shard
  int n_bmax;
  int l_max;                        // samples of the longest chain
  // uniform tiling (one segment of total_win adjacent window starts cut every 33*NT): the tile is computed from the
  // CTA index instead of being loaded (no dependent global load in front of the TMA issue)
  int uniform;
  int32_t s_base;
  int32_t total_win;
  // in-stream lower bound table (in the parameter block = constant memory).  ED: idx = order[k], a = zQ (sorted
  // order); DTW: idx = evenly spread positions, a = centre, b = half width of the query envelope there
  int32_t scr_idx[kScreenTerms];
  double scr_a[kScreenTerms];
  double scr_b[kScreenTerms];
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

constexpr int kQ1Cap = 64, kQ1Drain = 32;  // candidate queue per warp: drained 32 at a time (full lanes) as soon as it holds 32

constexpr size_t stream_xs_doubles(int nt, int m) { return ((size_t)kGroup * nt + m + 42) & ~size_t(1); }
constexpr size_t stream_gs_doubles(int nt, int m) { return ((size_t)nt + (m + kGroup - 1) / kGroup + 8) & ~size_t(1); }
constexpr size_t stream_smem_bytes(int nt, int m) {
  return 16 + sizeof(double) * (stream_xs_doubles(nt, m) + 2 * stream_gs_doubles(nt, m)) +
         (size_t)(nt / 32) * kQ1Cap * (2 * sizeof(double) + sizeof(int32_t)) + 16;
}

// Flagging, warp-aggregated.  A warp's windows span at most 33*32 adjacent starts, i.e. a handful of chains, and a
// near match flags hundreds of them: the per-chain maximum (chain_last) is therefore kept in warp-uniform registers
// and written with ONE returning atomic when the chain changes or the warp is done, instead of one per window (32
// lanes hammering one address per drain made the CTAs that hold a near match run for ~100 us — the tail of the whole
// launch).  need_bits takes one fire-and-forget atomicOr per window; the window count is reduced at the end.
struct FlagCache {
  int p;        // chain whose maximum is pending (-1 = none); warp-uniform
  int mx;       // pending maximum of (window start - chain begin); warp-uniform
  unsigned n;   // this lane's flagged windows
};

__device__ __forceinline__ void flag_flush(const StreamParams& P, FlagCache& F, int lane) {
  if (F.p >= 0 && lane == 0) {
    const int old = atomicMax(P.chain_last + F.p, F.mx);
    if (old < 0) {
      const unsigned long long slot = atomicAdd(P.n_flagged, 1ULL);
      P.flagged[slot] = F.p;
    }
  }
  F.p = -1;
}

// All 32 lanes call this; `mine` = this lane flags window w of the tile; fl = ballot of `mine` (non-zero).
__device__ __forceinline__ void flag_warp(const StreamParams& P, const StreamTile& tile, int w, bool mine, unsigned fl,
                                          FlagCache& F, int lane) {
  int p = -1, rel = -1;
  if (mine) {
    const int s = tile.s0 + w;
    const unsigned v = (unsigned)(tile.v0 + w);
    if (tile.chain < 0) {
      p = P.chains.chain_of(v);
    } else {
      p = tile.chain;
      while (s >= P.chains.begin(p) + P.chains.count(p)) p++;  // chains of one segment are consecutive
    }
    atomicOr(P.need_bits + (v >> 5), 1u << (v & 31));
    rel = s - P.chains.begin(p);
    F.n++;
  }
  unsigned rem = fl;
  while (rem) {  // one round per distinct chain among the flagged lanes
    const int pl = __shfl_sync(kFullMask, p, __ffs(rem) - 1);
    const bool same = mine && p == pl;
    const int mx = __reduce_max_sync(kFullMask, same ? rel : -1);
    if (pl != F.p) {
      flag_flush(P, F, lane);
      F.p = pl;
      F.mx = mx;
    } else {
      F.mx = max(F.mx, mx);
    }
    rem &= ~__ballot_sync(kFullMask, same);
  }
}

#ifdef KVM_STREAM_PROF
__device__ unsigned long long g_stream_prof[16];  // thread 0 of every CTA: cycles per phase, CTA count
__device__ __forceinline__ long long stream_clock() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
#define STREAM_T(var) const long long var = stream_clock()
#define STREAM_SET(var) var = stream_clock()
#define STREAM_COUNT(i, n) do { if ((n) != 0) atomicAdd(&g_stream_prof[i], (unsigned long long)(n)); } while (0)
#else
#define STREAM_COUNT(i, n)
#define STREAM_T(var)
#define STREAM_SET(var)
#endif

// Warp-private shared-memory queue of the windows the hot loop found inside the outer variance band: (window, ex, ex2).
// Drained 32 at a time with one window per lane (stream_candidate).
struct WarpQueue {
  int* w;
  double* ex;
  double* ex2;
};

// One queued candidate per lane: the exact outer-band tests, then the table tier of the lower bound — ED: the 32
// largest-|zQ| terms of the distance; DTW: LB_KimFL (K/utils/DtwUtils.java:149-189) and 32 evenly spread terms of
// LB_Keogh on the query envelope.  Returns 1 when the window certainly passes the gate and certainly is no answer (it
// is only counted), 2 when it is ambiguous or survives the table: the caller flags it, its chain is re-walked exactly and
// the exact stage (cnsm_ed_exact_kernel / cnsm_dtw_lb_list_kernel, a warp resp. a thread per window over the whole
// GPU) finishes the bound.  (A second in-stream tier over all m terms was measured: clusters of near matches made
// single CTAs run for 0.3 ms.)  `wv` = the window's first sample in shared memory.
template <int kMode>
__device__ __forceinline__ unsigned stream_candidate(const StreamParams& P, const StreamTile& tile, const double* __restrict__ wv,
                                                     int w, double ex, double ex2, const StreamGate& G, const double thr) {
  // (centred band tests, see StreamGate; a NaN sum fails them like it fails the interval tests)
  const double de = fabs(ex - G.e_mid);
  if (!le_nonneg(de, G.e_rad_out)) return 0u;
  const double t2 = __fma_rn(P.dm, ex2, -__dmul_rn(ex, ex));
  const double dv = fabs(t2 - G.v_mid);
  if (!le_nonneg(dv, G.v_rad_out)) return 0u;
  // rad_in < 0 (bit pattern negative as an integer) makes both tests false
  const bool sure = __double_as_longlong(G.e_rad_in) >= __double_as_longlong(de) && __double_as_longlong(G.v_rad_in) >= __double_as_longlong(dv);
  bool survive = true;
  const int m = P.m;
  if (__double_as_longlong(t2) > 0LL && !P.force_all) {  // t2 > 0 (finite here: it passed the band test)
    const double mean = ex * P.inv_m;
    const double rstd = rsqrt(t2 * P.inv_m2);
    if (kMode == 1 && m >= 6) {
      const double* __restrict__ q = P.q_full;
      const double x0 = (wv[0] - mean) * rstd, x1 = (wv[1] - mean) * rstd, x2 = (wv[2] - mean) * rstd;
      const double y0 = (wv[m - 1] - mean) * rstd, y1 = (wv[m - 2] - mean) * rstd, y2 = (wv[m - 3] - mean) * rstd;
      const double q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
      const double p0 = __ldg(q + m - 1), p1 = __ldg(q + m - 2), p2 = __ldg(q + m - 3);
      auto sq = [](double a, double b) { const double d = a - b; return d * d; };
      auto mn = [](double a, double b) { return min_nonneg(a, b); };
      double kim = sq(x0, q0) + sq(y0, p0);
      kim += mn(mn(sq(x1, q0), sq(x0, q1)), sq(x1, q1));
      kim += mn(mn(sq(y1, p0), sq(y0, p1)), sq(y1, p1));
      kim += mn(mn(mn(sq(x0, q2), sq(x1, q2)), sq(x2, q2)), mn(sq(x2, q1), sq(x2, q0)));
      kim += mn(mn(mn(sq(y0, p2), sq(y1, p2)), sq(y2, p2)), mn(sq(y2, p1), sq(y2, p0)));
      survive = le_nonneg(kim, thr);
    }
    const int n_scr = P.n_screen;
    double dist = 0.0;
    auto term = [&](int k, double wk) {
      const double x = (wk - mean) * rstd;
      double d;
      if (kMode == 0) {
        d = x - P.scr_a[k];
      } else {  // scr_a = centre, scr_b = half width of the envelope at this position: excess = max(|x - c| - h, 0)
        const double e = fabs(x - P.scr_a[k]) - P.scr_b[k];
        d = (__double2hiint(e) < 0) ? 0.0 : e;
      }
      return d;
    };
    if (n_scr == kScreenTerms) {
      // The usual case (m >= 32), fully unrolled: every table entry is a constant-bank operand with an immediate
      // offset, the eight window samples of a round are loaded up front and summed in two independent chains; the
      // partial bound is tested once per round.  (With run-time indices each term was a serial LDC -> LDS -> 4 x FP64
      // chain of its own, ~130 cycles: a tile full of candidates then ran for ~100 us and set the launch's tail.)
      // rounds of 4, 4, 8, 8, 8 terms: with eps^2 ~ 25 and |zQ| ~ 2.5 .. 3 on the leading terms a window that is no match
      // is over the threshold after three or four of them, so the first test comes early
      auto round = [&](auto k0c, auto cntc) {
        constexpr int k0 = decltype(k0c)::value, cnt = decltype(cntc)::value;
        double wk[cnt];
#pragma unroll
        for (int u = 0; u < cnt; u++) wk[u] = wv[P.scr_idx[k0 + u]];
        double acc0 = 0.0, acc1 = 0.0;
#pragma unroll
        for (int u = 0; u < cnt; u += 2) {
          const double da = term(k0 + u, wk[u]), db = term(k0 + u + 1, wk[u + 1]);
          acc0 = __fma_rn(da, da, acc0);
          acc1 = __fma_rn(db, db, acc1);
        }
        dist += acc0 + acc1;
        survive = le_nonneg(dist, thr);
      };
      using std::integral_constant;
      if (survive) round(integral_constant<int, 0>{}, integral_constant<int, 4>{});
      if (survive) round(integral_constant<int, 4>{}, integral_constant<int, 4>{});
      if (survive) round(integral_constant<int, 8>{}, integral_constant<int, 8>{});
      if (survive) round(integral_constant<int, 16>{}, integral_constant<int, 8>{});
      if (survive) round(integral_constant<int, 24>{}, integral_constant<int, 8>{});
    } else {
      for (int kk = 0; kk < n_scr && survive; kk++) {
        const double d = term(kk, wv[P.scr_idx[kk]]);
        dist = __fma_rn(d, d, dist);
        survive = le_nonneg(dist, thr);
      }
    }
  }
  return (!survive && sure) ? 1u : 2u;
}

// Drain up to 32 entries from the front of the queue (one per lane); the rest moves to the front.  n = entries
// (warp-uniform); returns the new count.
template <int kMode>
__device__ __forceinline__ int stream_drain(const StreamParams& P, const StreamTile& tile, const double* __restrict__ xs,
                                            const WarpQueue& Q, int n, int lane, const StreamGate& G, const double thr,
                                            unsigned& my_gate, FlagCache& F) {
  const int nb = min(n, 32), rest = n - nb;
  int mw = 0;
  double mex = 0.0, mex2 = 0.0;
  if (lane < rest) {
    mw = Q.w[nb + lane];
    mex = Q.ex[nb + lane];
    mex2 = Q.ex2[nb + lane];
  }
  int w = 0;
  unsigned verdict = 0u;
  if (lane < nb) {
    w = Q.w[lane];
    verdict = stream_candidate<kMode>(P, tile, xs + w, w, Q.ex[lane], Q.ex2[lane], G, thr);
  }
  my_gate += verdict & 1u;
  const unsigned fl = __ballot_sync(kFullMask, verdict == 2u);
  if (fl) flag_warp(P, tile, w, verdict == 2u, fl, F, lane);
  __syncwarp();
  if (lane < rest) {
    Q.w[lane] = mw;
    Q.ex[lane] = mex;
    Q.ex2[lane] = mex2;
  }
  __syncwarp();
  return rest;
}

__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// kMode 0: cNSM-ED, 1: cNSM-DTW (they differ in the table tier of the in-stream lower bound only)
//
// One CTA per tile of up to 33*NT adjacent window starts; thread t owns the chain of 33 windows starting at 33t.
//   1. One elected lane issues the TMA bulk copies of the tile's nwin + m - 1 samples (one mbarrier); meanwhile the last
//      warp evaluates the tile's guard bands from the block-maximum table.
//   2. Thread t sums the 33 samples (and their squares) of group t, t + NT, ...   — barrier —
//   3. Warp-local: the window sums of the warp's first chain are a direct sum of group sums, the other 31 chains follow
//      from a shuffle scan of gs[g + m/33] - gs[g] (+ m % 33 samples each).
//   4. Chain-level skip from the group sums (see below); only warps with an unskippable chain slide, branch-free,
//      recording the windows inside the outer variance band; those are replayed into the warp's queue and finished
//      one per lane.
template <int NT, int kMode>
__global__ void __launch_bounds__(NT, NT <= 192 ? 3 : (NT <= 256 ? 2 : 1)) cnsm_stream_kernel(StreamParams P) {
  extern __shared__ __align__(16) unsigned char stream_smem[];
  __shared__ StreamGate s_gate;
  __shared__ double s_thr;
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(kFullMask, tid >> 5, 0);  // (tells the compiler it is warp-uniform)
  STREAM_T(c0);
  StreamTile tile;
  if (P.uniform) {
    const int i0 = (int)blockIdx.x * (kGroup * NT);
    tile.s0 = P.s_base + i0;
    tile.nwin = min(kGroup * NT, P.total_win - i0);
    tile.chain = -1;
    tile.v0 = i0;
  } else {
    tile = P.tiles[blockIdx.x];
  }
  const int m = P.m;
  const int nwin = tile.nwin;
  const int ns = nwin + m - 1;                 // samples the tile's windows cover
  const int s0a = tile.s0 & ~1;                // 16-byte aligned copy start
  const int lead = tile.s0 - s0a;
  const int nload = (lead + ns + 1) & ~1;      // even number of doubles
  const int ng = (ns + kGroup - 1) / kGroup;   // 33-sample groups holding samples
  const size_t n_xs = stream_xs_doubles(NT, m);
  const int gs_cap = (int)stream_gs_doubles(NT, m);
  double* xs_raw = reinterpret_cast<double*>(stream_smem + 16);
  double* xs = xs_raw + lead;
  double* gs1 = xs_raw + n_xs;
  double* gs2 = gs1 + gs_cap;
  double* q_ex = gs2 + gs_cap;
  double* q_ex2 = q_ex + (NT / 32) * kQ1Cap;
  int* q_w = reinterpret_cast<int*>(q_ex2 + (NT / 32) * kQ1Cap);
  const uint32_t bar = smem_u32(stream_smem);

  if (warp == 0) {
    if (elect_one()) {
      mbar_init(bar, 1);
      fence_mbar_init();
      const uint32_t bytes = (uint32_t)nload * 8u;
      mbar_expect_tx(bar, bytes);
      const unsigned char* src = reinterpret_cast<const unsigned char*>(P.T + s0a);
      const uint32_t dst = smem_u32(xs_raw);
      constexpr uint32_t kChunk = 32768;
      for (uint32_t o = 0; o < bytes; o += kChunk) tma_bulk_g2s(dst + o, src + o, min(kChunk, bytes - o), bar);
    }
  }
  // Guard bands of this tile (last warp; finished after barrier (1), while the samples are in flight): A = max |sample|
  // over every sample a chain can have summed before or inside one of the tile's windows, i.e. from l_max samples
  // before the tile to its end (block-maximum table, kBmaxBlock granularity).  The table load is issued first.
  const int bm_lo = max(0, tile.s0 - P.l_max) / kBmaxBlock;
  const int bm_hi = min(P.n_bmax - 1, (tile.s0 + ns) / kBmaxBlock);
  double amax_first = 0.0;  // (not consumed before the barrier: the load stays in flight)
  if (warp == NT / 32 - 1 && bm_lo + lane <= bm_hi) amax_first = __ldg(P.bmax + bm_lo + lane);
  // zero what follows the copied samples (never part of a valid window; keeps the group sums finite)
  for (int i = nload + tid; i < lead + ns + kGroup + 2; i += NT) xs_raw[i] = 0.0;
  __syncthreads();  // (1) mbarrier initialised; zero tail staged
  STREAM_T(c1);
  if (warp == NT / 32 - 1) {
    unsigned long long mx = (unsigned long long)__double_as_longlong(amax_first);
    for (int b = bm_lo + 32 + lane; b <= bm_hi; b += 32) mx = max(mx, (unsigned long long)__double_as_longlong(__ldg(P.bmax + b)));
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
  mbar_wait(bar, 0);
  STREAM_T(c2);
#ifdef KVM_STREAM_PROF
  if (P.force_all == 99) return;  // experiment: the copy alone
  if (P.force_all == 98 && blockIdx.x > 1000000) return;
#endif

  // ---- sums of 33-sample groups (lane stride 33 doubles: conflict-free); three partial sums per group for ILP;
  // groups past the last sample are written as zeros up to the arrays' capacity
  for (int g = tid; g < gs_cap; g += NT) {
    double a1[3] = {0.0, 0.0, 0.0}, a2[3] = {0.0, 0.0, 0.0};
    if (g < ng) {
      const double* __restrict__ x = xs + g * kGroup;
#pragma unroll
      for (int j = 0; j < kGroup / 3; j++) {
#pragma unroll
        for (int u = 0; u < 3; u++) {
          const double v = x[3 * j + u];
          a1[u] += v;
          a2[u] = __fma_rn(v, v, a2[u]);
        }
      }
    }
    gs1[g] = (a1[0] + a1[1]) + a1[2];
    gs2[g] = (a2[0] + a2[1]) + a2[2];
  }
  __syncthreads();  // (2) group sums and guard bands ready
  STREAM_T(c3);

  // ---- warp-local from here on
  const int w0 = tid * kGroup;
  unsigned my_gate = 0;
  FlagCache F{-1, -1, 0u};
  if (warp * 32 * kGroup < nwin) {  // (warp-uniform) the warp holds at least one window
    const int qa = m / kGroup, qb = m - qa * kGroup;
    const int gw = 32 * warp;  // the warp's first chain = group index
    // window sums of the warp's first chain: W(gw) = sum_{k < qa} gs[gw + k]
    double W1 = 0.0, W2 = 0.0;
    for (int k = lane; k < qa; k += 32) {
      W1 += gs1[gw + k];
      W2 += gs2[gw + k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      W1 += __shfl_xor_sync(kFullMask, W1, o);
      W2 += __shfl_xor_sync(kFullMask, W2, o);
    }
    // chain g = tid: W(g) - W(gw) = sum_{gw <= k < g} (gs[k + qa] - gs[k])
    const int g = tid;
    const double so = gs1[g], go = gs2[g];               // sum / sum of squares of the chain's outgoing samples
    const double r1a = gs1[g + qa], r2a = gs2[g + qa];
    const double r1b = qb ? gs1[g + qa + 1] : 0.0, r2b = qb ? gs2[g + qa + 1] : 0.0;
    const double d1 = r1a - so, d2 = r2a - go;
    double i1 = d1, i2 = d2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u1 = __shfl_up_sync(kFullMask, i1, o), u2 = __shfl_up_sync(kFullMask, i2, o);
      if (lane >= o) {
        i1 += u1;
        i2 += u2;
      }
    }
    double ex = W1 + (i1 - d1), ex2 = W2 + (i2 - d2);
    {
      const double* __restrict__ x = xs + (g + qa) * kGroup;
      for (int j = 0; j < qb; j++) {
        const double v = x[j];
        ex += v;
        ex2 = __fma_rn(v, v, ex2);
      }
    }
    const int nw = max(0, min(kGroup, nwin - w0));
    const StreamGate& G = s_gate;
    // Chain-level skip.  Along the chain's 32 slide steps, with y = x - c centred on the first window's own mean
    // c = ex/m (t2 = m*ex2 - ex^2 does not depend on c, and ex_y starts at 0):
    //   ex2_y moves within [-Gout, +Gin]   (Gout / Gin = sum of y^2 over the outgoing / incoming samples, from the
    //                                        group sums: G - c*(2S - n*c), n = 33 outgoing, 33 or 66 incoming)
    //   |ex_y| <= sum|y_in| + sum|y_out| <= T, T^2 <= 2*(n_in*Gin + 33*Gout)         (Cauchy-Schwarz)
    // so t2 stays within [t2 - m*Gout - T^2, t2 + m*Gin] and ex within ex -+ T.  If those ranges miss the outer
    // variance band or the outer mean band, none of the chain's windows can be inside both: the chain is not slid.
    // (errG covers the cancellation in G - c*(2S - n*c): 2|c S| <= n c^2 + G.)
    bool need;
    {
      const double e_mid = 0.5 * (G.e_lo_out + G.e_hi_out), e_rad = 0.5 * (G.e_hi_out - G.e_lo_out);
      const double n_in = qb ? 66.0 : 33.0;
      const double si = r1a + r1b, gi = r2a + r2b;
      const double c = ex * P.inv_m;
      const double errG = 64.0 * 1.1102230246251565e-16 * (gi + go + 99.0 * c * c);
      const double Gin = fmax(gi - c * (2.0 * si - n_in * c), 0.0) + errG;
      const double Gout = fmax(go - c * (2.0 * so - 33.0 * c), 0.0) + errG;
      const double T2 = 2.0 * (n_in * Gin + 33.0 * Gout) * (1.0 + 1e-12);
      const double t2v = __fma_rn(P.dm, ex2, -__dmul_rn(ex, ex));
      const double dmean = fabs(ex - e_mid) - e_rad;
      const bool skip_mean = dmean > 0.0 && dmean * dmean > T2;
      const bool skip_var = (t2v + P.dm * Gin * (1.0 + 1e-12) < G.v_lo_out) || (t2v - (P.dm * Gout + T2) * (1.0 + 1e-12) > G.v_hi_out);
      need = !(skip_mean || skip_var) && nw > 0;
    }
    STREAM_T(c4);
    STREAM_COUNT(10, need ? 1 : 0);  // chains that could not be skipped
    if (__any_sync(kFullMask, need)) {
      STREAM_COUNT(9, lane == 0 ? 1 : 0);  // warps that slide
      WarpQueue Q;
      Q.ex = q_ex + warp * kQ1Cap;
      Q.ex2 = q_ex2 + warp * kQ1Cap;
      Q.w = q_w + warp * kQ1Cap;
      const double* __restrict__ xo = xs + w0;
      const double* __restrict__ xi = xs + w0 + m;
      // Hot-loop test: the outer variance band only, on the high word of t2 = m*ex2 - ex^2 as a signed integer range
      // (for t2 >= 0 the high words order like the values; a band that reaches down to 0 or below admits every
      // negative t2).  DSETP issues at a fraction of the DADD rate on sm_100, and the mean band is already enforced
      // per chain by the skip test; stream_candidate() repeats both tests exactly.
      int kv_lo = (G.v_lo_out > 0.0) ? __double2hiint(G.v_lo_out) : (int)0x80000000;
      const int kv_hi = (G.v_hi_out >= 0.0) ? __double2hiint(G.v_hi_out) : -1;
      if (!(G.v_hi_out >= 0.0)) kv_lo = 0x7fffffff;  // empty band
      const double dmv = P.dm;
      // slide: branch-free, one dependent DADD / DFMA per window; bit j = window j inside the band
      unsigned long long mk = 0ULL;
      {
        double e1 = ex, e2 = ex2;
#pragma unroll
        for (int j = 0; j < kGroup; j++) {
          const double a = xi[j], o = xo[j];
          const int h = __double2hiint(__fma_rn(dmv, e2, -__dmul_rn(e1, e1)));
          mk |= (h >= kv_lo && h <= kv_hi) ? (1ULL << j) : 0ULL;
          const double dl = a - o, sm = a + o;
          e1 += dl;
          e2 = __fma_rn(dl, sm, e2);
        }
      }
      if (!need) mk = 0ULL;
      if (nw < kGroup) mk &= (1ULL << nw) - 1ULL;
      STREAM_COUNT(11, __popcll(mk));  // candidates (hot-loop bits)
      const unsigned any_lo = __reduce_or_sync(kFullMask, (unsigned)mk), any_hi = __reduce_or_sync(kFullMask, (unsigned)(mk >> 32));
      const unsigned long long all = ((unsigned long long)any_hi << 32) | any_lo;
      const int dens = __reduce_add_sync(kFullMask, __popcll(mk));
      if (dens >= 16 * kGroup) {
        // A warp at least half full of candidates (a region whose statistics match the query's) finishes them in place:
        // the replay hands every lane its own window sums, so no queue, no compaction — what those buy when one lane
        // in ten holds a candidate costs a third of the instructions when nearly all of them do.
        double rex = ex, rex2 = ex2;
#pragma unroll 1
        for (int j = 0; (all >> j) != 0ULL; j++) {
          unsigned verdict = 0u;
          if ((mk >> j) & 1ULL) verdict = stream_candidate<kMode>(P, tile, xs + w0 + j, w0 + j, rex, rex2, G, s_thr);
          my_gate += verdict & 1u;
          const unsigned fl = __ballot_sync(kFullMask, verdict == 2u);
          if (fl) flag_warp(P, tile, w0 + j, verdict == 2u, fl, F, lane);
          const double av = xi[j], ov = xo[j];
          const double dl = av - ov, sm = av + ov;
          rex += dl;
          rex2 = __fma_rn(dl, sm, rex2);
        }
        flag_flush(P, F, lane);
      } else if (all != 0ULL) {
        // replay (same operations, same values) step by step, each step pushing its candidates into the queue
        int n1 = 0;
        double rex = ex, rex2 = ex2;
        for (int j = 0; (all >> j) != 0ULL; j++) {
          const bool hit = (mk >> j) & 1ULL;
          const unsigned bal = __ballot_sync(kFullMask, hit);
          if (hit) {
            const int slot = n1 + __popc(bal & ((1u << lane) - 1u));
            Q.w[slot] = w0 + j;
            Q.ex[slot] = rex;
            Q.ex2[slot] = rex2;
          }
          n1 += __popc(bal);
          const double av = xi[j], ov = xo[j];
          const double dl = av - ov, sm = av + ov;
          rex += dl;
          rex2 = __fma_rn(dl, sm, rex2);
          __syncwarp();
          if (n1 >= kQ1Drain) n1 = stream_drain<kMode>(P, tile, xs, Q, n1, lane, G, s_thr, my_gate, F);
        }
        while (n1 > 0) n1 = stream_drain<kMode>(P, tile, xs, Q, n1, lane, G, s_thr, my_gate, F);
        flag_flush(P, F, lane);
      }
    }
  }
  STREAM_T(c5);
  // ---- certain gate passes, flagged windows
  const unsigned tot = __reduce_add_sync(kFullMask, my_gate), tot_flag = __reduce_add_sync(kFullMask, F.n);
  if (lane == 0 && tot) atomicAdd(P.gate_pass, (unsigned long long)tot);
  if (lane == 0 && tot_flag) atomicAdd(P.n_need, (unsigned long long)tot_flag);
#ifdef KVM_STREAM_PROF
  if (tid == 0) {
    const long long c6 = stream_clock();
    const long long d[9] = {c1 - c0, c2 - c1, c3 - c2, 0, 0, c5 - c3, c6 - c5, 0, 1};
    for (int i = 0; i < 9; i++) atomicAdd(&g_stream_prof[i], (unsigned long long)d[i]);
    atomicMax(&g_stream_prof[12], (unsigned long long)(c6 - c0));        // the slowest CTA
    if (c6 - c0 > 60000) atomicAdd(&g_stream_prof[13], 1ULL);           // CTAs slower than 60k cycles
    atomicMax(&g_stream_prof[14], (unsigned long long)(c5 - c3));        // the slowest warp-local phase (thread 0's warp)
  }
#endif
}

// ---------------------------------------------------------------------------------------------------------------
struct RewalkParams {
  const double* __restrict__ T;
  ChainTable chains;
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

constexpr int kRewalkStages = 12;  // 32-position tiles in flight per warp (covers ~2.5k cycles of DRAM latency at 16 cycles per position)

// One warp per flagged chain.  All lanes run the same recurrence (the chain is sequential; what the warp buys is
// coalesced, deeply prefetched sample rows and nothing but LDS + DADD/DMUL on the dependent path):
//   K/NormQueryEngine.java:498-499  ex += d; ex2 += d*d        (every sample)
//   K/NormQueryEngine.java:523-524  ex -= T[j]; ex2 -= T[j]^2  (after each complete window)
// Out-of-range slots of a 32-position tile are zero-filled: adding / subtracting +0.0 is the identity here (the sums
// start at +0.0 and x + y is -0.0 only when both are).
template <bool kSq>
__device__ __forceinline__ void chain_rewalk_body(const RewalkParams& P, int block, int n_blocks);

__global__ void __launch_bounds__(32) chain_rewalk_kernel(RewalkParams P) { chain_rewalk_body<true>(P, blockIdx.x, gridDim.x); }

// Several independent chain sets in one launch (the window-mean pass: one per width): blockIdx.y picks the set.
struct RewalkBatch {
  RewalkParams set[5];
};
// (window means need the sum only: kSq = false leaves the squares out; the sets are interleaved over the block index so
// that the few busy warps of every set spread over different SMs)
__global__ void __launch_bounds__(32) chain_rewalk_batch_kernel(RewalkBatch B, int n_sets) {
  const int set = blockIdx.x % n_sets;
  chain_rewalk_body<false>(B.set[set], blockIdx.x / n_sets, gridDim.x / n_sets);
}

template <bool kSq>
__device__ __forceinline__ void chain_rewalk_body(const RewalkParams& P, int block, int n_blocks) {
  __shared__ __align__(16) double s_in[kRewalkStages][32];
  __shared__ __align__(16) double s_out[kRewalkStages][32];
  __shared__ double s_ex[32], s_ex2[32];
  const int lane = threadIdx.x;
  const int m = P.m;
  const unsigned long long n = *P.n_flagged;
  for (unsigned long long ci = block; ci < n; ci += n_blocks) {
    const int p = P.flagged[ci];
    const int cb = P.chains.begin(p);
    const int last = P.chain_last[p];
    const int vb = P.chains.ordinal(p);
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
          if (kSq) ex2 = xadd(ex2, xmul(a.x, a.x));
          ex = xadd(ex, a.y);
          if (kSq) ex2 = xadd(ex2, xmul(a.y, a.y));
        }
      } else if (mask == 0) {
#pragma unroll
        for (int h = 0; h < 16; h++) {
          const double2 a = pin[h], o = pout[h];
          ex = xadd(ex, a.x);
          if (kSq) ex2 = xadd(ex2, xmul(a.x, a.x));
          ex = xsub(ex, o.x);
          if (kSq) ex2 = xsub(ex2, xmul(o.x, o.x));
          ex = xadd(ex, a.y);
          if (kSq) ex2 = xadd(ex2, xmul(a.y, a.y));
          ex = xsub(ex, o.y);
          if (kSq) ex2 = xsub(ex2, xmul(o.y, o.y));
        }
      } else {
        // a tile with flagged windows: the same walk, every post-add state parked in shared memory; afterwards lane pp
        // emits position pp if its window is flagged (one counter update per warp, coalesced stores)
#pragma unroll
        for (int h = 0; h < 16; h++) {
          const double2 a = pin[h], o = pout[h];
          ex = xadd(ex, a.x);
          if (kSq) ex2 = xadd(ex2, xmul(a.x, a.x));
          s_ex[2 * h] = ex;
          s_ex2[2 * h] = ex2;
          ex = xsub(ex, o.x);
          if (kSq) ex2 = xsub(ex2, xmul(o.x, o.x));
          ex = xadd(ex, a.y);
          if (kSq) ex2 = xadd(ex2, xmul(a.y, a.y));
          s_ex[2 * h + 1] = ex;
          s_ex2[2 * h + 1] = ex2;
          ex = xsub(ex, o.y);
          if (kSq) ex2 = xsub(ex2, xmul(o.y, o.y));
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

// ---- The same tail over peer memory (NVLink / NVSwitch P2P), one kernel: every rank writes its packed block straight
// into every other rank's exchange buffer, publishes a sequence number behind a system-scope fence, and waits until
// the sequence numbers of all ranks have arrived in its own buffer.  No library call on the data path; the blocks are
// 4 KB, so latency is everything (an ncclAllGather of them costs ~65 us, this kernel a few us plus the wait for the
// slowest rank).  Two alternating areas: a rank can run at most one exchange ahead of the slowest one (it needs that
// rank's sequence number to finish), and the slowest rank reads area k before it starts exchange k + 1 (stream order).
constexpr int kXchgSlot = 16 + 2 * 256;  // = kGatherHead + 2 * kGatherCap doubles per rank
struct XchgParams {
  double* peer[8];              // base of every rank's exchange buffer as mapped here
  const double* send;           // this rank's block (pinned host memory, read once)
  int rank, world, parity;
  unsigned long long seq;
  int* err;
};
__device__ __forceinline__ size_t xchg_seq_base(int world) { return (size_t)2 * world * kXchgSlot; }
__global__ void __launch_bounds__(256) xchg_kernel(XchgParams P) {
  const int tid = threadIdx.x;
  const size_t area = (size_t)P.parity * P.world * kXchgSlot;
  for (int p = 0; p < P.world; p++) {
    double* dst = P.peer[p] + area + (size_t)P.rank * kXchgSlot;
    for (int i = tid; i < kXchgSlot; i += blockDim.x) dst[i] = P.send[i];
  }
  __threadfence_system();
  __syncthreads();
  if (tid < P.world) {
    unsigned long long* flag = reinterpret_cast<unsigned long long*>(P.peer[tid] + xchg_seq_base(P.world)) + (size_t)P.parity * P.world + P.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(P.seq) : "memory");
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(P.peer[P.rank] + xchg_seq_base(P.world)) + (size_t)P.parity * P.world + tid;
    const long long t0 = clock64();
    unsigned long long v;
    do {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(mine) : "memory");
      if (v >= P.seq) break;
      if (clock64() - t0 > 6000000000LL) {  // ~3 s: a rank never arrived
        *P.err = 1;
        break;
      }
      __nanosleep(64);
    } while (true);
  }
  __syncthreads();
}

}  // namespace kvm
