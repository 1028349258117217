// cnsm_kernels.cuh — constrained z-normalised matching (cNSM), shared by the ED and DTW engines.
// Replaces K/NormQueryEngine.java:432-528 and the statistics half of K/NormQueryEngineDtw.java:457-603.
//
// Why a "chain walker": the reference's window mean/std come from a sequential add-then-subtract
// chain (ex += d; ...; ex -= T[j]) that restarts at every merged interval.  Its rounding errors
// accumulate along the chain, so a prefix-sum or fresh-sum kernel cannot reproduce it to better than
// ~1e-7 relative on long chains — not enough for bit-exact alpha/beta gates or 1e-9 distances.  The
// chain is therefore emulated exactly: ONE THREAD PER CHAIN, the 32 lanes of a warp walking 32
// different chains in lock-step.  Samples reach the lanes through a multi-stage ring of shared-memory
// tiles filled asynchronously with 16-byte cp.async copies (two full 256-byte rows per instruction), preceded
// by L2 line prefetches, so HBM sees only long sequential streams and the warp's issue slots go to the
// two dependent FP64 chains.
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

constexpr int kWalkTile = 32;      // samples (columns) per shared-memory tile row
constexpr int kWalkPitch = 34;     // row pitch in doubles: 16-byte aligned rows, conflict-free lane-private LDS.128
constexpr int kFifoDepth = 16;     // per-lane staging of work-list entries between flushes
constexpr int kFifoPitch = kFifoDepth + 1;
constexpr int kFrontPad = 64;      // zero samples the ctx keeps in front of / behind the series so that
constexpr int kTailPad = 192;      // whole-row bulk copies never leave the allocation
constexpr int kPrefetchTiles = 16; // L2 prefetch distance of the incoming stream, in tiles
constexpr int kEvalTile = 128;     // work-list entries per evaluator tile (= evaluator CTA size)

constexpr int walk_tile_doubles(int stages) { return stages * 2 * 32 * kWalkPitch; }
constexpr size_t walk_smem_bytes(int stages) {
  return sizeof(double) * (size_t)(walk_tile_doubles(stages) + 2 * 32 * kFifoPitch) + sizeof(int32_t) * 32 * kFifoPitch +
         16;
}

// ---- asynchronous global->shared copies (LDGSTS) ----------------------------------------------------
// Per-lane 256-byte TMA row copies (cp.async.bulk) were measured first: the TMA unit serialises such small
// requests (~60 cycles each, 64 per tile), which made the walker ~10x slower than 16-byte cp.async.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_hint(uint32_t dst, const void* src, unsigned long long policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(policy)
               : "memory");
}
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void l2_prefetch_line(const void* src) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(src));
}

// Monotone integer key of a double's high word: key(a) <= key(b) whenever a <= b.  Range tests on it
// cost integer-pipe instructions only; the bounds are widened by one key unit on each side by the host.
__device__ __forceinline__ int hi_key(double x) {
  const int h = __double2hiint(x);
  return h ^ ((h >> 31) & 0x7fffffff);
}

struct WalkParams {
  const double* __restrict__ T;         // sample 0 of this shard; kFrontPad/kTailPad zero samples surround it
  const int32_t* __restrict__ cbegin;   // per chain: local 0-based index of its first sample
  const int32_t* __restrict__ cnsamp;   // per chain: samples in the chain (0 = nothing to do)
  const long long* __restrict__ region_base;  // per walker warp: first work-list slot of its region
  int K;
  int m;
  int32_t first_global;
  int idx_hi;  // last even local index a 16-byte copy may start at (inside the tail pad)
  int prefetch_tiles;  // L2 prefetch distance of the incoming stream in tiles (0 = off)
  int l2_hints;        // 1: incoming rows evict_last, outgoing rows evict_first
  // conservative pre-gate on the chain sums (superset of the exact gate; see DESIGN.md "cNSM pre-gate"):
  //   key(ex) in [mean_klo, mean_klo + mean_kspan]   <=>  |ex/m - meanQ| <= beta (+slack)
  //   key(m*ex2 - ex^2) in [var_klo, var_klo + var_kspan]  <=>  (std/stdQ) in [1/alpha, alpha] (+slack)
  int mean_klo, var_klo;
  unsigned mean_kspan, var_kspan;
  double dm;
  int32_t* e_off;
  double* e_ex;
  double* e_ex2;
  int32_t* region_count;
};

// One CTA = one warp = 32 chains = one work-list region.  STAGES-deep ring of (incoming, outgoing) tiles,
// each row = 32 consecutive samples of one chain, filled by 16-byte cp.async copies (half a warp per row,
// so every copy instruction moves two full 256-byte rows) and read back with lane-private LDS.128.
// Rows are 16-byte aligned in global memory: the incoming row starts at pos & ~1, and kDelta = 1 when m is
// even (the outgoing row is then aligned one sample later, so its columns lag the incoming ones by one).
template <int STAGES, int kDelta>
__global__ void __launch_bounds__(32) cnsm_walk_kernel(WalkParams P) {
  extern __shared__ __align__(16) unsigned char walk_smem_raw[];
  double* tiles = reinterpret_cast<double*>(walk_smem_raw);   // [STAGES][2][32][pitch]
  double* f_ex = tiles + walk_tile_doubles(STAGES);           // [32][kFifoPitch]
  double* f_ex2 = f_ex + 32 * kFifoPitch;
  int32_t* f_off = reinterpret_cast<int32_t*>(f_ex2 + 32 * kFifoPitch);

  const int lane = threadIdx.x;
  const int region = blockIdx.x;
  const int c = region * 32 + lane;
  int pos = 0, len = 0;
  if (c < P.K) {
    pos = P.cbegin[c];
    len = P.cnsamp[c];
  }
  const int m = P.m;
  const int sha = pos & 1;                 // column of sample 0 in the incoming tiles
  const int ab = pos - sha;                // 16-byte aligned base of the incoming rows
  const int ob = pos - (m - 1) - sha + kDelta;  // aligned base of the outgoing rows (parity of m-1 fixed by kDelta)
  const int ntl = (len > 0) ? (len + sha + kWalkTile - 1) / kWalkTile : 0;
  const int ntiles = warp_max_i32(ntl);
  if (ntiles == 0) {
    if (lane == 0) P.region_count[region] = 0;
    return;
  }
  const double* __restrict__ T = P.T;
  const int idx_lo = -kFrontPad, idx_hi = P.idx_hi;  // even bounds of the padded allocation

  // copy i of a tile: lanes 0-15 fill row 2i, lanes 16-31 row 2i+1, 16 bytes each
  int a_idx[16], o_idx[16];
  {
    const int half = lane >> 4, piece = (lane & 15) * 2;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      a_idx[i] = __shfl_sync(kFullMask, ab, 2 * i + half) + piece;
      o_idx[i] = __shfl_sync(kFullMask, ob, 2 * i + half) + piece;
    }
  }
  const bool hints = P.l2_hints != 0;
  const int pf_tiles = P.prefetch_tiles;
  const unsigned long long pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
  const uint32_t dst0 = smem_u32(tiles) + (uint32_t)(((lane >> 4) * kWalkPitch + (lane & 15) * 2) * 8);
  constexpr uint32_t kStageBytes = 2 * 32 * kWalkPitch * 8, kStreamBytes = 32 * kWalkPitch * 8;
  constexpr uint32_t kPairBytes = 2 * kWalkPitch * 8;

  auto issue = [&](int k) {
    const uint32_t dst = dst0 + (uint32_t)(k % STAGES) * kStageBytes;
    const int koff = k * kWalkTile;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int ia = max(min(a_idx[i] + koff, idx_hi), idx_lo);
      const int io = max(min(o_idx[i] + koff, idx_hi), idx_lo);
      if (hints) {
        cp_async16_hint(dst + i * kPairBytes, T + ia, pol_keep);
        cp_async16_hint(dst + kStreamBytes + i * kPairBytes, T + io, pol_drop);
      } else {
        cp_async16(dst + i * kPairBytes, T + ia);
        cp_async16(dst + kStreamBytes + i * kPairBytes, T + io);
      }
    }
    if (pf_tiles > 0) {
      const int pf = min(ab + (k + pf_tiles) * kWalkTile, idx_hi);
      l2_prefetch_line(T + pf);
      l2_prefetch_line(T + pf + 16);
    }
  };

  const long long base = P.region_base[region];
  int rcount = 0, cnt = 0;
  double ex = 0.0, ex2 = 0.0, carry = 0.0;
  const int mean_klo = P.mean_klo, var_klo = P.var_klo;
  const unsigned mean_kspan = P.mean_kspan, var_kspan = P.var_kspan;
  const double dm = P.dm;
  const int32_t off0 = P.first_global + pos - (m - 1);  // window start (1-based, global) of the window ending at sample 0

  // Warp-cooperative flush: entry e of the staged block goes to lane e%32, so the global stores are
  // coalesced and each chain's entries stay contiguous (the evaluator's lanes then read neighbouring windows).
  auto flush = [&]() {
    __syncwarp();
    const int incl = warp_incl_scan_i32(cnt, lane);
    const int total = __shfl_sync(kFullMask, incl, 31);
    const int excl = incl - cnt;
    for (int e0 = 0; e0 < total; e0 += 32) {
      const int e = e0 + lane;
      int owner = 0;
#pragma unroll
      for (int step = 16; step > 0; step >>= 1) {
        const int t = __shfl_sync(kFullMask, excl, owner + step);
        if (t <= e) owner += step;
      }
      const int j = e - __shfl_sync(kFullMask, excl, owner);
      if (e < total) {
        const long long g = base + rcount + e;
        P.e_off[g] = f_off[owner * kFifoPitch + j];
        P.e_ex[g] = f_ex[owner * kFifoPitch + j];
        P.e_ex2[g] = f_ex2[owner * kFifoPitch + j];
      }
    }
    rcount += total;
    cnt = 0;
    __syncwarp();
  };

  double* const my_ex = f_ex + lane * kFifoPitch;
  double* const my_ex2 = f_ex2 + lane * kFifoPitch;
  int32_t* const my_off = f_off + lane * kFifoPitch;
  auto step = [&](int s, double a, double o) {
    const bool act = (unsigned)s < (unsigned)len;
    const bool win = act & (s >= m - 1);
    a = act ? a : 0.0;
    o = win ? o : 0.0;
    ex = xadd(ex, a);                 // K/NormQueryEngine.java:498
    ex2 = xadd(ex2, xmul(a, a));      // :499
    const double v = __fma_rn(ex2, dm, -(ex * ex));
    const bool pass = win & ((unsigned)(hi_key(ex) - mean_klo) <= mean_kspan) &
                      ((unsigned)(hi_key(v) - var_klo) <= var_kspan);
    if (pass) {
      my_ex[cnt] = ex;
      my_ex2[cnt] = ex2;
      my_off[cnt] = off0 + s;
      cnt++;
    }
    ex = xsub(ex, o);                 // :523
    ex2 = xsub(ex2, xmul(o, o));      // :524
  };

#pragma unroll
  for (int k = 0; k < STAGES - 1; k++) {
    if (k < ntiles) issue(k);
    cp_async_commit();
  }
  for (int k = 0; k < ntiles; k++) {
    if (k + STAGES - 1 < ntiles) issue(k + STAGES - 1);  // refills the stage that tile k-1 used
    cp_async_commit();
    cp_async_wait<STAGES - 1>();  // tile k has landed (groups complete in order)
    __syncwarp();
    const int stage = k % STAGES;
    const double* ra = tiles + (size_t)stage * (2 * 32 * kWalkPitch) + lane * kWalkPitch;
    const double* ro = ra + 32 * kWalkPitch;
    const int sbase = k * kWalkTile - sha;
#pragma unroll
    for (int blk = 0; blk < kWalkTile / 8; blk++) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const int col = blk * 8 + 2 * i;
        const double2 A = *reinterpret_cast<const double2*>(ra + col);
        const double2 O = *reinterpret_cast<const double2*>(ro + col);
        const double o0 = kDelta ? carry : O.x;
        const double o1 = kDelta ? O.x : O.y;
        carry = O.y;
        step(sbase + col, A.x, o0);
        step(sbase + col + 1, A.y, o1);
      }
      if (__any_sync(kFullMask, cnt > kFifoDepth - 8)) flush();
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

constexpr int kExactChunk = 4096;  // terms staged in shared memory per pass (32 KB per warp)

// K/NormQueryEngine.java:513-520 verbatim arithmetic: x = (T[order[k]+j]-mean)/std; dist += (x-zQ[k])^2.
// One warp per survivor: all lanes compute the per-term values (divisions in parallel, each term rounded
// exactly as the reference's), lane 0 then adds them in the reference's order — the sum is bit-identical,
// and the critical path is one dependent DADD per term instead of one scattered DRAM load per term.
__global__ void __launch_bounds__(128) cnsm_ed_exact_kernel(ExactEdParams P) {
  extern __shared__ double exact_terms[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  double* term = exact_terms + (size_t)warp * kExactChunk;
  unsigned long long n = *P.in.count;
  if ((long long)n > P.in.cap) n = (unsigned long long)P.in.cap;
  const int m = P.m;
  for (unsigned long long e = (unsigned long long)blockIdx.x * n_warps + warp; e < n;
       e += (unsigned long long)gridDim.x * n_warps) {
    const int32_t off = P.in.off[e];
    const double mean = P.in.mean[e], stdv = P.in.stdv[e];
    const double* __restrict__ w = P.T + (off - P.first_global);
    double dist = 0.0;
    bool alive = true;
    for (int k0 = 0; k0 < m && alive; k0 += kExactChunk) {
      const int kc = min(kExactChunk, m - k0);
      __syncwarp();
      for (int k = lane; k < kc; k += 32) {
        const double x = xdiv(xsub(w[__ldg(P.order + k0 + k)], mean), stdv);
        term[k] = xsqdist(x, __ldg(P.zq + k0 + k));
      }
      __syncwarp();
      if (lane == 0) {
        int k = 0;
        for (; k + 8 <= kc && alive; k += 8) {
#pragma unroll
          for (int u = 0; u < 8; u++) dist = xadd(dist, term[k + u]);
          alive = dist <= P.eps2;
        }
        if (alive) {
          for (; k < kc; k++) dist = xadd(dist, term[k]);
          alive = dist <= P.eps2;
        }
      }
      alive = __shfl_sync(kFullMask, alive, 0);
    }
    if (lane == 0 && dist <= P.eps2) P.sink.emit(off, xsqrt(dist));
  }
}

}  // namespace kvm
