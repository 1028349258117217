// cnsm_kernels.cuh — constrained z-normalised matching (cNSM), shared by the ED and DTW engines.
// Replaces K/NormQueryEngine.java:432-528 and the statistics half of K/NormQueryEngineDtw.java:457-603.
//
// Why a "chain walker": the reference's window mean/std come from a sequential add-then-subtract
// chain (ex += d; ...; ex -= T[j]) that restarts at every merged interval.  Its rounding errors
// accumulate along the chain, so a prefix-sum or fresh-sum kernel cannot reproduce it to better than
// ~1e-7 relative on long chains — not enough for bit-exact alpha/beta gates or 1e-9 distances.  The
// chain is therefore emulated exactly: one lane per chain, 32 chains per CTA, and the only sequential
// part — the (ex, ex2) recurrence — relayed between warps so that nothing else sits on its critical path.
// Samples reach the lanes through a multi-stage ring of shared-memory tiles filled asynchronously with
// 16-byte cp.async copies (two full 256-byte rows per instruction) carrying L2 eviction hints.
//
//   cnsm_relay_kernel  chain-exact ex/ex2 per window + a cheap conservative alpha/beta pre-gate;
//                      windows that may pass are appended (offset, ex, ex2) to the CTA's private
//                      region of the work list (no global atomics in the streaming loop);
//                      kMode 1: per-window key buckets for IndexBuilder's window-mean pass
//   plan_scan_cta      exclusive scan of per-region tile counts -> flat tile index for the evaluators, run by the
//                      last walker CTA to finish; kMode 2: one statistics pass gating a set of queries
//   cnsm_ed_eval_kernel  one thread per work-list entry: exact mean/std/gate (reference arithmetic),
//                      then a fast 32-term FMA screen in |zQ|-descending order against
//                      eps^2*(1+1e-9); survivors go to the exact list
//   cnsm_ed_exact_kernel one warp per survivor: warp-cooperative fast distance, then the reference's
//                      sequential, unfused sum -> the accepted distances are bit-identical to the Java loop's
#pragma once
#include <type_traits>

#include "common.cuh"

namespace kvm {

constexpr int kWalkTile = 32;      // samples (columns) per shared-memory tile row
constexpr int kFrontPad = 64;      // zero samples the ctx keeps in front of / behind the series so that
constexpr int kTailPad = 192;      // whole-row bulk copies never leave the allocation
constexpr int kFastTerms = 32;     // terms of the evaluator's fast screen
constexpr int kEvalTile = 128;     // work-list entries per evaluator tile (= evaluator CTA size)

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

// kMode 2 (query sets): the chain sums do not depend on the query, only the gate's key ranges do, so one statistics
// pass can gate a whole set of queries of one length over one interval list and feed one work list per query.
constexpr int kMaxBatch = 16;
struct BatchGate {
  int n_q;
  int mean_klo[kMaxBatch], var_klo[kMaxBatch];
  unsigned mean_kspan[kMaxBatch], var_kspan[kMaxBatch];
  int32_t* e_off[kMaxBatch];
  double* e_ex[kMaxBatch];
  double* e_ex2[kMaxBatch];
  int32_t* region_count[kMaxBatch];
  int32_t* tile_prefix[kMaxBatch];
  unsigned long long* totals[kMaxBatch];
};

struct WalkParams {
  const double* __restrict__ T;         // sample 0 of this shard; kFrontPad/kTailPad zero samples surround it
  const int32_t* __restrict__ cbegin;   // per chain: local 0-based index of its first sample
  const int32_t* __restrict__ cnsamp;   // per chain: samples in the chain (0 = nothing to do)
  const long long* __restrict__ region_base;  // per walker warp: first work-list slot of its region
  int K;
  int m;
  int32_t first_global;
  int idx_hi;  // last even local index a 16-byte copy may start at (inside the tail pad)
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
  // the last CTA to finish turns the region counts into the evaluators' tile index (null in kMode 1)
  int32_t* tile_prefix;
  unsigned long long* totals;
  unsigned int* done;  // zeroed before the launch
  // kMode == 1 (IndexBuilder window means): per-window bucket ids instead of a work list
  int32_t* bucket_out;  // bucket_out[local window start] = floor(2 * fl(fl(ex/w) * 10))
  double c20w;          // 20 / w
  int* overflow;        // set when a bucket does not fit int32
  const BatchGate* batch;  // kMode == 2 only (device memory)
};

// ---- named barriers (producer/consumer hand-off inside a CTA) --------------------------------------
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// ---- mbarriers (tile ring hand-off between the loader warp and the chain warp) ----------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// the executing thread's arrival is triggered when all its prior cp.async copies have landed
__device__ __forceinline__ void mbar_arrive_on_cp_async(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}

// ---------------------------------------------------------------------------------------------------
// Relay walker.  The only sequential thing in the statistics pass is the chain state (ex, ex2): loading
// the samples, squaring them, gating the window sums and appending to the work list are not.  So the
// CTA's 32 chains are walked by kRelayWarps warps taking turns: a turn is kRelayBlock window positions.
// A warp's turn has three phases,
//   prepare  wait for the tile, pull the turn's samples from shared memory into registers (incoming a[],
//            outgoing o[]), mask them if the turn touches a chain end (and, for 16-position turns, square them);
//   walk     receive (ex, ex2) from the previous turn's warp (shared-memory slot + named barrier), run the
//            2 x 2 dependent DADDs per position with no memory instruction in the stream, hand the state on;
//   gate     pre-gate the turn's post-add sums from registers and append the passing ones,
// and only `walk` is on the critical path: the other warps prepare and gate while one walks.  Per
// position the critical path is 2 dependent DADDs (16.1 cycles) plus 1/kRelayBlock of a hand-off
// (measured ~230 cycles: STS -> BAR.ARV -> BAR.SYNC -> LDS; polling a shared-memory slot instead was
// slower, ~540 cycles, and so was an mbarrier).
// Tiles: incoming rows as in the walker above; outgoing rows carry one extra leading 16-byte pair
// (34 samples, pitch 38) so that a turn never needs the previous tile when m is even (kDelta = 1: the
// outgoing sample of tile column j is row element j + 1; kDelta = 0: element j + 2).
#ifndef KVM_RELAY_WARPS
#define KVM_RELAY_WARPS 5
#endif
#ifndef KVM_RELAY_BLOCK
#define KVM_RELAY_BLOCK 16
#endif
constexpr int kRelayWarps = KVM_RELAY_WARPS;
constexpr int kRelayThreads = 32 * (kRelayWarps + 1);  // warps 0..kRelayWarps-1 take turns, the last warp loads tiles
constexpr int kRelayBlock = KVM_RELAY_BLOCK;           // window positions per turn: 16 or 32
constexpr int kRelayTurnsPerTile = kWalkTile / kRelayBlock;
constexpr bool kRelaySquareInWalk = kRelayBlock > 16;  // 32-position turns: no registers left for precomputed squares
static_assert(kRelayBlock == 16 || kRelayBlock == 32, "a turn is half a tile or a tile");
constexpr int kInPitch = 34;   // doubles; 16-byte aligned rows, conflict-free lane-private LDS.128
constexpr int kOutPitch = 38;  // 34 samples + pad: lane stride 12 banks (mod 32) -> conflict-free LDS.128
constexpr int kRelayStageDoubles = 32 * (kInPitch + kOutPitch);
constexpr int kGatePitch = kRelayBlock + 1;  // double2 per lane row of a relay warp's flush staging (dense turns only)
constexpr size_t relay_smem_bytes(int stages) {
  return sizeof(double) * (size_t)stages * kRelayStageDoubles + sizeof(double2) * kRelayWarps * 32 +
         sizeof(double2) * kRelayWarps * 32 * kGatePitch + sizeof(unsigned long long) * 2 * stages + 16;
}

// Exclusive scan over the walker regions by one CTA of NT threads: tile_prefix[r] = sum_{r'<r} ceil(count[r']/kEvalTile),
// totals[0] = #tiles, totals[1] = #entries.  Run by the last walker CTA to finish (no separate launch); the counts
// come from other CTAs, hence the L2 (cache-global) loads.
template <int NT>
__device__ __forceinline__ void plan_scan_cta(const int32_t* region_count, int n_regions, int32_t* __restrict__ tile_prefix,
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
  for (int r0 = 0; r0 < n_regions; r0 += NT) {
    const int r = r0 + tid;
    const int cnt = (r < n_regions) ? __ldcg(region_count + r) : 0;
    my_entries += (unsigned long long)cnt;
    const int tiles = (cnt + kEvalTile - 1) / kEvalTile;
    const int incl = warp_incl_scan_i32(tiles, lane);
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = (lane < NT / 32) ? s_warp[lane] : 0;
      const int wi = warp_incl_scan_i32(w, lane);
      s_warp[lane] = wi - w;
    }
    __syncthreads();
    const int excl = s_carry + s_warp[warp] + incl - tiles;
    if (r < n_regions) tile_prefix[r] = excl;
    __syncthreads();
    if (tid == NT - 1) s_carry = excl + tiles;
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

#ifdef KVM_RELAY_PROF
__device__ unsigned long long g_relay_prof[8 * 8];  // per relay warp: cycles in tile wait, prepare, state wait, walk, gate; turns
__device__ __forceinline__ long long relay_clock() {
  long long t;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t)::"memory");
  return t;
}
#define RELAY_T(var) const long long var = relay_clock()
#define RELAY_ACC(acc, expr) acc += (expr)
#else
#define RELAY_T(var)
#define RELAY_ACC(acc, expr)
#endif

template <int STAGES, int kDelta, int kMode = 0>
__global__ void __launch_bounds__(kRelayThreads, kRelayThreads <= 224 ? 2 : 1) cnsm_relay_kernel(WalkParams P) {
  extern __shared__ __align__(16) unsigned char walk_smem_raw[];
  double* tiles = reinterpret_cast<double*>(walk_smem_raw);  // [STAGES][ in 32 x kInPitch | out 32 x kOutPitch ]
  double2* state = reinterpret_cast<double2*>(tiles + (size_t)STAGES * kRelayStageDoubles);  // [2][32]
  double2* gate_stage = state + kRelayWarps * 32;                                                      // [kRelayWarps][32][kGatePitch]
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(gate_stage + kRelayWarps * 32 * kGatePitch);
  int* s_rcount = reinterpret_cast<int*>(bars + 2 * STAGES);

  __shared__ int s_rcount_q[kMode == 2 ? kMaxBatch : 1];

  int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  asm volatile("mov.u32 %0, %0;" : "+r"(lane));  // keep out of SR_TID.X re-reads in the loops
  asm volatile("mov.u32 %0, %0;" : "+r"(warp));
  const int region = blockIdx.x;
  const int c = region * 32 + lane;
  int pos = 0, len = 0;
  if (c < P.K) {
    pos = P.cbegin[c];
    len = P.cnsamp[c];
  }
  const int m = P.m;
  const int sha = pos & 1;  // column of sample 0 in the incoming tiles
  const int ntl = (len > 0) ? (len + sha + kWalkTile - 1) / kWalkTile : 0;
  const int ntiles = warp_max_i32(ntl);
  const int k_w = (m - 1) / kWalkTile;  // first tile that can contain the end of a complete window
  const uint32_t bar_ready = smem_u32(bars), bar_free = bar_ready + 8 * STAGES;
  if (kMode == 2 && threadIdx.x < kMaxBatch) s_rcount_q[threadIdx.x] = 0;
  if (threadIdx.x == 0) {
    *s_rcount = 0;
#pragma unroll
    for (int s = 0; s < STAGES; s++) {
      mbar_init(bar_ready + 8 * s, 32);
      mbar_init(bar_free + 8 * s, 32 * kRelayTurnsPerTile);  // every lane of the tile's turns arrives for itself
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kRelayWarps) {
    // ------------------------------------------------------------------ tile loader
    const int ab = pos - sha;                         // 16-byte aligned base of the incoming rows
    const int ob = pos - (m - 1) - sha + kDelta - 2;  // aligned base of the outgoing rows (one pair early)
    const double* __restrict__ T = P.T;
    const int idx_lo = -kFrontPad, idx_hi = P.idx_hi;
    int a_idx[16], o_idx[16];
    {
      const int half = lane >> 4, piece = (lane & 15) * 2;
#pragma unroll
      for (int i = 0; i < 16; i++) {
        a_idx[i] = __shfl_sync(kFullMask, ab, 2 * i + half) + piece;
        o_idx[i] = __shfl_sync(kFullMask, ob, 2 * i + half) + piece + 2;
      }
    }
    const unsigned long long pol_keep = l2_policy_evict_last(), pol_drop = l2_policy_evict_first();
    const uint32_t t0 = smem_u32(tiles);
    const uint32_t dst_in = t0 + (uint32_t)(((lane >> 4) * kInPitch + (lane & 15) * 2) * 8);
    const uint32_t dst_out = t0 + (uint32_t)((32 * kInPitch + (lane >> 4) * kOutPitch + (lane & 15) * 2 + 2) * 8);
    const uint32_t dst_lead = t0 + (uint32_t)((32 * kInPitch + lane * kOutPitch) * 8);
    constexpr uint32_t kStageBytes = kRelayStageDoubles * 8;
    for (int k = 0; k < ntiles; k++) {
      const int stage = k % STAGES;
      if (k >= STAGES) mbar_wait(bar_free + 8 * stage, (uint32_t)(((k / STAGES) - 1) & 1));
      const uint32_t so = (uint32_t)stage * kStageBytes;
      const int koff = k * kWalkTile;
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int ia = max(min(a_idx[i] + koff, idx_hi), idx_lo);
        cp_async16_hint(dst_in + so + i * (2 * kInPitch * 8), T + ia, pol_keep);
      }
      if (k >= k_w) {  // warm-up tiles subtract nothing (their outgoing operands are masked): do not fetch them
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const int io = max(min(o_idx[i] + koff, idx_hi), idx_lo);
          cp_async16_hint(dst_out + so + i * (2 * kOutPitch * 8), T + io, pol_drop);
        }
        const int io = max(min(ob + koff, idx_hi), idx_lo);  // this lane's row, leading pair
        cp_async16_hint(dst_lead + so, T + io, pol_drop);
      }
      mbar_arrive_on_cp_async(bar_ready + 8 * stage);
    }
  } else {
    // ------------------------------------------------------------------ relay warps
    const int nb = kRelayTurnsPerTile * ntiles;
    const long long base = P.region_base[region];
    const int mean_klo = P.mean_klo, var_klo = P.var_klo;
    const unsigned mean_kspan = P.mean_kspan, var_kspan = P.var_kspan;
    const double dm = P.dm;
    const int32_t off0 = P.first_global + pos - (m - 1) - sha;  // + tile column = 1-based global window start
#ifdef KVM_RELAY_PROF
    long long pf_tile = 0, pf_prep = 0, pf_state = 0, pf_walk = 0, pf_gate = 0, pf_turns = 0;
#endif
    auto turn = [&](int b, auto steady_tag) {
      constexpr bool kSteady = decltype(steady_tag)::value;
      RELAY_T(t_a);
      const int k = b / kRelayTurnsPerTile, h = b % kRelayTurnsPerTile;
      const int col0 = b * kRelayBlock;  // tile column of the turn's first position
      const int stage = k % STAGES;
      // ---- prepare
      const double* ra = tiles + (size_t)stage * kRelayStageDoubles + lane * kInPitch + h * kRelayBlock;
      const double* ro = tiles + (size_t)stage * kRelayStageDoubles + 32 * kInPitch + lane * kOutPitch + h * kRelayBlock;
      double a[kRelayBlock], o[kRelayBlock];
#pragma unroll
      for (int i = 0; i < kRelayBlock / 2; i++) {
        const double2 v = *reinterpret_cast<const double2*>(ra + 2 * i);
        a[2 * i] = v.x;
        a[2 * i + 1] = v.y;
      }
      if (kDelta) {
        double2 v = *reinterpret_cast<const double2*>(ro);
        o[0] = v.y;
#pragma unroll
        for (int i = 1; i < kRelayBlock / 2; i++) {
          v = *reinterpret_cast<const double2*>(ro + 2 * i);
          o[2 * i - 1] = v.x;
          o[2 * i] = v.y;
        }
        v = *reinterpret_cast<const double2*>(ro + kRelayBlock);
        o[kRelayBlock - 1] = v.x;
      } else {
#pragma unroll
        for (int i = 0; i < kRelayBlock / 2; i++) {
          const double2 v = *reinterpret_cast<const double2*>(ro + 2 + 2 * i);
          o[2 * i] = v.x;
          o[2 * i + 1] = v.y;
        }
      }
      const int s0 = col0 - sha;  // chain-relative index of the turn's first sample
      if (!kSteady) {
#pragma unroll
        for (int j = 0; j < kRelayBlock; j++) {
          const int s = s0 + j;
          const bool act = (unsigned)s < (unsigned)len;
          a[j] = act ? a[j] : 0.0;
          o[j] = (act & (s >= m - 1)) ? o[j] : 0.0;
        }
      }
      constexpr int kSq = kRelaySquareInWalk ? 1 : kRelayBlock;
      double as[kSq], os[kSq];
      if (!kRelaySquareInWalk) {
#pragma unroll
        for (int j = 0; j < kSq; j++) {
          as[j] = xmul(a[j], a[j]);
          os[j] = xmul(o[j], o[j]);
        }
        // pin the prepared operands in front of the hand-off (an empty volatile asm is not moved across bar.sync)
#pragma unroll
        for (int j = 0; j < kSq; j++) asm volatile("" : "+d"(as[j]), "+d"(os[j]));
      } else {
#pragma unroll
        for (int j = 0; j < kRelayBlock; j++) asm volatile("" : "+d"(a[j]), "+d"(o[j]));
      }
      mbar_arrive(bar_free + 8 * stage);  // this lane's reads of the tile are done (its loaded values have been consumed)
      RELAY_T(t_b);
      // ---- walk
      double ex = 0.0, ex2 = 0.0;
      if (b > 0) {
        bar_sync(1 + ((b - 1) % kRelayWarps), 64);
        const double2 st = state[((b - 1) & 1) * 32 + lane];
        ex = st.x;
        ex2 = st.y;
      }
#ifdef KVM_RELAY_PROF
      asm volatile("" : "+d"(ex), "+d"(ex2));  // the state has arrived (BAR.SYNC blocks at its first consumer)
#endif
      RELAY_T(t_c);
      double qx[kRelayBlock], q2[kRelayBlock];
#pragma unroll
      for (int j = 0; j < kRelayBlock; j++) {
        const double aj2 = kRelaySquareInWalk ? xmul(a[j], a[j]) : as[j % kSq];
        const double oj2 = kRelaySquareInWalk ? xmul(o[j], o[j]) : os[j % kSq];
        ex = xadd(ex, a[j]);   // K/NormQueryEngine.java:498
        ex2 = xadd(ex2, aj2);  // :499
        qx[j] = ex;
        q2[j] = ex2;
        ex = xsub(ex, o[j]);   // :523
        ex2 = xsub(ex2, oj2);  // :524
      }
      if (b + 1 < nb) {
        state[(b & 1) * 32 + lane] = make_double2(ex, ex2);
        bar_arrive(1 + (b % kRelayWarps), 64);
      }
      RELAY_T(t_d);
      RELAY_ACC(pf_prep, t_b - t_a);
      RELAY_ACC(pf_state, t_c - t_b);
      RELAY_ACC(pf_walk, t_d - t_c);
      RELAY_ACC(pf_turns, 1);
      // ---- gate
      if (k >= k_w) {  // (warm-up tiles: no complete window ends there)
        if (kMode == 1) {
          // IndexBuilder step 1 (K/IndexBuilder.java:251-265): key bucket of every window mean; b from one multiply
          // when ex*(20/w) is clear of an integer by more than the rounding slack, else the reference's exact
          // divide / multiply (MeanIntervalUtils.toRound, K/utils/MeanIntervalUtils.java:51-61).
#pragma unroll
          for (int j = 0; j < kRelayBlock; j++) {
            const int s = s0 + j;
            const bool win = kSteady || (((unsigned)s < (unsigned)len) & (s >= m - 1));
            const double e = qx[j];
            const double v2 = e * P.c20w;
            double fl = floor(v2);
            const double frac = v2 - fl;
            const double gd = fabs(v2) * 4e-15 + 1e-290;
            if (win && !(frac >= gd && frac <= 1.0 - gd)) {
              const double v = xmul(xdiv(e, dm), 10.0);
              fl = floor(xadd(v, v));  // 2v is exact
            }
            if (win) {
              if (!(fabs(fl) < 2147483000.0)) *P.overflow = 1;
              P.bucket_out[pos + s - (m - 1)] = (int)fl;
            }
          }
        } else if (kMode == 2) {
          // query set: the window keys once, then one (mean, variance) range test per query and position
          const BatchGate* __restrict__ G = P.batch;
          const int n_q = __ldg(&G->n_q);
          int km[kRelayBlock], kv[kRelayBlock];
          unsigned wmask = 0;
#pragma unroll
          for (int j = 0; j < kRelayBlock; j++) {
            const int s = s0 + j;
            const bool win = kSteady || (((unsigned)s < (unsigned)len) & (s >= m - 1));
            wmask |= win ? (1u << j) : 0u;
            km[j] = hi_key(qx[j]);
            kv[j] = hi_key(__fma_rn(q2[j], dm, -(qx[j] * qx[j])));
          }
          // the turn's range of mean keys per lane: most queries are rejected by two compares and a vote
          int km_lo = INT32_MAX, km_hi = INT32_MIN;
#pragma unroll
          for (int j = 0; j < kRelayBlock; j++) {
            if ((wmask >> j) & 1u) {
              km_lo = min(km_lo, km[j]);
              km_hi = max(km_hi, km[j]);
            }
          }
          bool staged = false;  // this turn's sums are in the warp's staging rows (dense flushes)
          for (int q = 0; q < n_q; q++) {
            const int mklo = __ldg(&G->mean_klo[q]), vklo = __ldg(&G->var_klo[q]);
            const unsigned mspan = __ldg(&G->mean_kspan[q]), vspan = __ldg(&G->var_kspan[q]);
            // [km_lo, km_hi] misses [mklo, mklo + mspan] for every lane -> nothing of this turn can pass
            if (!__any_sync(kFullMask, (long long)km_hi >= (long long)mklo && (long long)km_lo <= (long long)mklo + (long long)mspan))
              continue;
            unsigned mask = 0;
#pragma unroll
            for (int j = 0; j < kRelayBlock; j++)
              mask |= (((unsigned)(km[j] - mklo) <= mspan) & ((unsigned)(kv[j] - vklo) <= vspan)) ? (1u << j) : 0u;
            mask &= wmask;
            if (!__any_sync(kFullMask, mask != 0)) continue;
            int32_t* __restrict__ qe_off = G->e_off[q];
            double* __restrict__ qe_ex = G->e_ex[q];
            double* __restrict__ qe_ex2 = G->e_ex2[q];
            const int cnt = __popc(mask);
            const int incl = warp_incl_scan_i32(cnt, lane);
            const int total = __shfl_sync(kFullMask, incl, 31);
            int rbase = 0;
            if (lane == 0) rbase = atomicAdd(&s_rcount_q[q], total);
            rbase = __shfl_sync(kFullMask, rbase, 0);
            const int excl = incl - cnt;
            if (kRelayBlock == 16 && total > 96) {
              double2* gs = gate_stage + (size_t)warp * 32 * kGatePitch;
              if (!staged) {
#pragma unroll
                for (int j = 0; j < kRelayBlock; j++) gs[lane * kGatePitch + j] = make_double2(qx[j], q2[j]);
                __syncwarp();
                staged = true;
              }
              const unsigned nz = __ballot_sync(kFullMask, cnt > 0);
              const int hl = lane & 15, hw = lane >> 4;
              for (int pr = 0; pr < 16; pr++) {
                if (((nz >> (2 * pr)) & 3u) == 0) continue;
                const int owner = 2 * pr + hw;
                const unsigned omask = __shfl_sync(kFullMask, mask, owner);
                const int oexcl = __shfl_sync(kFullMask, excl, owner);
                const int32_t ooff = __shfl_sync(kFullMask, off0, owner);
                if ((omask >> hl) & 1u) {
                  const double2 v2 = gs[owner * kGatePitch + hl];
                  const long long gidx = base + rbase + oexcl + __popc(omask & ((1u << hl) - 1u));
                  qe_off[gidx] = ooff + col0 + hl;
                  qe_ex[gidx] = v2.x;
                  qe_ex2[gidx] = v2.y;
                }
              }
            } else {
              long long gidx = base + rbase + excl;
#pragma unroll
              for (int j = 0; j < kRelayBlock; j++) {
                if ((mask >> j) & 1u) {
                  qe_off[gidx] = off0 + col0 + j;
                  qe_ex[gidx] = qx[j];
                  qe_ex2[gidx] = q2[j];
                  gidx++;
                }
              }
            }
          }
          if (staged) __syncwarp();
        } else {
          // mean gate first (integer pipe only); the variance keys are computed only if some window of the turn passed it
          unsigned mask = 0;
#pragma unroll
          for (int j = 0; j < kRelayBlock; j++) {
            const int s = s0 + j;
            const bool win = kSteady || (((unsigned)s < (unsigned)len) & (s >= m - 1));
            const bool pass = win & ((unsigned)(hi_key(qx[j]) - mean_klo) <= mean_kspan);
            mask |= pass ? (1u << j) : 0u;
          }
          if (__any_sync(kFullMask, mask != 0)) {
#pragma unroll
            for (int j = 0; j < kRelayBlock; j++) {
              const double v = __fma_rn(q2[j], dm, -(qx[j] * qx[j]));
              const bool pass = (unsigned)(hi_key(v) - var_klo) <= var_kspan;
              mask &= pass ? 0xffffffffu : ~(1u << j);
            }
            if (__any_sync(kFullMask, mask != 0)) {
              const int cnt = __popc(mask);
              const int incl = warp_incl_scan_i32(cnt, lane);
              const int total = __shfl_sync(kFullMask, incl, 31);
              int rbase = 0;
              if (lane == 0) rbase = atomicAdd(s_rcount, total);
              rbase = __shfl_sync(kFullMask, rbase, 0);
              const int excl = incl - cnt;
              if (kRelayBlock == 16 && total > 96) {
                // dense turn (a matching region): transpose through this warp's staging rows so that every store
                // instruction writes two runs of consecutive work-list slots — two chains per round, half a warp
                // each, lane c of a half copies column c of its chain if it passed
                double2* gs = gate_stage + (size_t)warp * 32 * kGatePitch;
#pragma unroll
                for (int j = 0; j < kRelayBlock; j++) gs[lane * kGatePitch + j] = make_double2(qx[j], q2[j]);
                __syncwarp();
                const unsigned nz = __ballot_sync(kFullMask, cnt > 0);
                const int hl = lane & 15, hw = lane >> 4;
                for (int pr = 0; pr < 16; pr++) {
                  if (((nz >> (2 * pr)) & 3u) == 0) continue;
                  const int owner = 2 * pr + hw;
                  const unsigned omask = __shfl_sync(kFullMask, mask, owner);
                  const int oexcl = __shfl_sync(kFullMask, excl, owner);
                  const int32_t ooff = __shfl_sync(kFullMask, off0, owner);
                  if ((omask >> hl) & 1u) {
                    const double2 v2 = gs[owner * kGatePitch + hl];
                    const long long gidx = base + rbase + oexcl + __popc(omask & ((1u << hl) - 1u));
                    P.e_off[gidx] = ooff + col0 + hl;
                    P.e_ex[gidx] = v2.x;
                    P.e_ex2[gidx] = v2.y;
                  }
                }
                __syncwarp();
              } else {
                long long gidx = base + rbase + excl;  // a chain's entries of the turn stay contiguous
#pragma unroll
                for (int j = 0; j < kRelayBlock; j++) {
                  if ((mask >> j) & 1u) {
                    P.e_off[gidx] = off0 + col0 + j;
                    P.e_ex[gidx] = qx[j];
                    P.e_ex2[gidx] = q2[j];
                    gidx++;
                  }
                }
              }
            }
          }
        }
      }
      RELAY_T(t_e);
      RELAY_ACC(pf_gate, t_e - t_d);
    };
    for (int b = warp; b < nb; b += kRelayWarps) {
      const int k = b / kRelayTurnsPerTile;
      RELAY_T(t_w0);
      mbar_wait(bar_ready + 8 * (k % STAGES), (uint32_t)((k / STAGES) & 1));  // tile k has landed
      RELAY_T(t_w1);
      RELAY_ACC(pf_tile, t_w1 - t_w0);
      const int sbase = b * kRelayBlock - sha;
      const bool steady = __all_sync(kFullMask, (sbase >= m - 1) && (sbase + kRelayBlock <= len));
      if (steady) turn(b, std::true_type{});
      else turn(b, std::false_type{});
    }
#ifdef KVM_RELAY_PROF
    if (lane == 0) {
      atomicAdd(&g_relay_prof[warp * 8 + 0], (unsigned long long)pf_tile);
      atomicAdd(&g_relay_prof[warp * 8 + 1], (unsigned long long)pf_prep);
      atomicAdd(&g_relay_prof[warp * 8 + 2], (unsigned long long)pf_state);
      atomicAdd(&g_relay_prof[warp * 8 + 3], (unsigned long long)pf_walk);
      atomicAdd(&g_relay_prof[warp * 8 + 4], (unsigned long long)pf_gate);
      atomicAdd(&g_relay_prof[warp * 8 + 5], (unsigned long long)pf_turns);
    }
#endif
  }
  __syncthreads();
  if (kMode == 0) {
    __shared__ bool s_last;
    if (threadIdx.x == 0) {
      P.region_count[region] = *s_rcount;
      __threadfence();
      s_last = atomicAdd(P.done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
      __threadfence();
      plan_scan_cta<kRelayThreads>(P.region_count, (int)gridDim.x, P.tile_prefix, P.totals);
    }
  } else if (kMode == 2) {
    __shared__ bool s_last2;
    const BatchGate* __restrict__ G = P.batch;
    const int n_q = G->n_q;
    if ((int)threadIdx.x < n_q) {
      G->region_count[threadIdx.x][region] = s_rcount_q[threadIdx.x];
      __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last2 = atomicAdd(P.done, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last2) {
      __threadfence();
      for (int q = 0; q < n_q; q++) {
        plan_scan_cta<kRelayThreads>(G->region_count[q], (int)gridDim.x, G->tile_prefix[q], G->totals[q]);
        __syncthreads();
      }
    }
  } else if (threadIdx.x == 0) {
    P.region_count[region] = *s_rcount;
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

__global__ void __launch_bounds__(kEvalTile, 16) cnsm_ed_eval_kernel(EvalParams P) {
  __shared__ int s_r;
  __shared__ unsigned int s_gate;
  if (threadIdx.x == 0) s_gate = 0;
  const int n_tiles = (int)P.totals[0];
  const int m = P.m;
  int r_prev = 0;  // thread 0: region of this CTA's previous tile (tiles are visited in ascending order)
  for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    __syncthreads();
    if (threadIdx.x == 0) {
      // tile -> region: gallop forward from the previous tile's region, then bisect.  In a matching region the next
      // tile of this CTA lies a region or two ahead: 2-3 dependent loads instead of log2(n_regions).
      int lo = r_prev, hi = r_prev + 1, step = 1;
      while (hi <= P.n_regions && __ldg(P.tile_prefix + hi) <= t) {
        lo = hi;
        step *= 2;
        hi = min(P.n_regions + 1, hi + step);
      }
      while (hi - lo > 1) {  // tile_prefix[lo] <= t < tile_prefix[hi] (tile_prefix[n_regions + 1] = +inf)
        const int mid = (lo + hi) >> 1;
        if (__ldg(P.tile_prefix + mid) <= t) lo = mid; else hi = mid;
      }
      r_prev = lo;
      s_r = lo;
    }
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
    const double rstd = 1.0 / stdv;  // x = (w - mean) * rstd: the error does not grow with |mean| / std
    const double* __restrict__ w = P.T + (off - P.first_global);
    // Fast screen: the kFastTerms largest-|zQ| terms (a lower bound of the full distance).  Windows still under
    // eps^2 after them are rare (near matches); they go to the warp-cooperative exact stage instead of having
    // one thread chase up to m scattered loads on its own.
    double dist = 0.0;
    const int kmax = min(m, kFastTerms);
    int k = 0;
    bool alive = true;
    for (; k + 4 <= kmax && alive; k += 4) {
      double wv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) wv[u] = w[__ldg(P.order + k + u)];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const double df = (wv[u] - mean) * rstd - __ldg(P.zq + k + u);
        dist = __fma_rn(df, df, dist);
      }
      alive = dist <= P.eps2_hi;
    }
    if (alive) {
      for (; k < kmax; k++) {
        const double df = (w[__ldg(P.order + k)] - mean) * rstd - __ldg(P.zq + k);
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

struct XList {  // windows whose exact chain sums were recomputed by chain_rewalk_kernel (stream_kernels.cuh)
  int32_t* off;
  double* ex;
  double* ex2;
  unsigned long long* count;
  long long cap;
};

struct ExactEdParams {
  const double* __restrict__ T;
  int32_t first_global;
  int m;
  const double* __restrict__ zq;
  const int32_t* __restrict__ order;
  double eps2, eps2_hi;
  CandList in;
  AnswerSink sink;
  unsigned long long* n_exact;  // windows that reached the reference-order summation
  // kFromSums: entries are (offset, ex, ex2); the exact gate runs here
  XList xin;
  double meanQ, stdQ, alpha, inv_alpha, beta;
  unsigned long long* gate_pass;
  int win_cap;  // doubles of per-warp window staging behind the term buffer (0 = none)
  int q_cap;    // doubles (and ints) of the per-CTA staging of zQ / order (0 = none)
};

// Screen in front of the exact stage (streaming path): one warp per re-walked window — exact statistics and gate from
// the chain sums (K/NormQueryEngine.java:508-511, counted), then the 128 largest-|zQ| terms of the distance in fast
// arithmetic, straight from global memory.  No shared memory, so the SM holds 64 warps of it and the scattered window
// gathers (the windows were streamed long ago: DRAM) overlap; survivors (rare: near matches) go to
// cnsm_ed_exact_kernel<false> as (offset, mean, std).  Inside the exact kernel the same screen ran at 8-12 warps per
// SM (its staging buffers): 105 us for 65 k windows at n = 1e9.
__global__ void __launch_bounds__(256) cnsm_ed_screen_kernel(ExactEdParams P) {
  const int lane = threadIdx.x & 31;
  const unsigned long long gwarp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  unsigned long long n = *P.xin.count;
  if ((long long)n > P.xin.cap) n = (unsigned long long)P.xin.cap;
  const int m = P.m;
  unsigned my_gate = 0;
  for (unsigned long long e = gwarp; e < n; e += n_warps) {
    double mean, stdv;
    const int32_t off = P.xin.off[e];
    if (!cnsm_exact_gate(P.xin.ex[e], P.xin.ex2[e], m, P.meanQ, P.stdQ, P.alpha, P.inv_alpha, P.beta, mean, stdv)) continue;
    my_gate++;
    const double* __restrict__ wg = P.T + (off - P.first_global);
    const double rstd = 1.0 / stdv;
    double part = 0.0;
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int k = u * 32 + lane;
      if (k < m) {
        const double df = (wg[__ldg(P.order + k)] - mean) * rstd - __ldg(P.zq + k);
        part = __fma_rn(df, df, part);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(kFullMask, part, o);
    if (lane == 0 && part <= P.eps2_hi) {
      const unsigned long long slot = atomicAdd(P.in.count, 1ULL);
      if ((long long)slot < P.in.cap) {
        P.in.off[slot] = off;
        P.in.mean[slot] = mean;
        P.in.stdv[slot] = stdv;
      }
    }
  }
  if (lane == 0 && my_gate) atomicAdd(P.gate_pass, (unsigned long long)my_gate);
}

constexpr int kExactChunk = 1024;  // terms staged in shared memory per pass (8 KB per warp)

// K/NormQueryEngine.java:513-520 verbatim arithmetic: x = (T[order[k]+j]-mean)/std; dist += (x-zQ[k])^2.
// One warp per survivor: all lanes compute the per-term values (divisions in parallel, each term rounded
// exactly as the reference's), lane 0 then adds them in the reference's order — the sum is bit-identical,
// and the critical path is one dependent DADD per term instead of one scattered DRAM load per term.
template <bool kFromSums>
__global__ void __launch_bounds__(128) cnsm_ed_exact_kernel(ExactEdParams P) {
  extern __shared__ double exact_terms[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  unsigned long long n = kFromSums ? *P.xin.count : *P.in.count;
  const long long cap = kFromSums ? P.xin.cap : P.in.cap;
  if ((long long)n > cap) n = (unsigned long long)cap;
  if ((unsigned long long)blockIdx.x * n_warps >= n) return;  // (CTA-uniform) nothing for this CTA
  const int m = P.m;
  // Shared memory: [q_cap doubles zQ | q_cap ints order] once per CTA (q_cap >= m: the |zQ|-ordered query, staged so
  // that the dependent order -> sample gathers never leave the SM), then per warp kExactChunk per-term values and
  // (win_cap >= m) the window itself, staged with coalesced loads.
  double* zq_s = exact_terms;
  int32_t* ord_s = reinterpret_cast<int32_t*>(exact_terms + P.q_cap);
  double* warp_base = exact_terms + P.q_cap + (P.q_cap + 1) / 2;
  double* term = warp_base + (size_t)warp * (kExactChunk + P.win_cap);
  double* win = term + kExactChunk;
  const bool staged = P.win_cap >= m;
  const bool q_staged = P.q_cap >= m;
  if (q_staged) {
    for (int k = threadIdx.x; k < m; k += blockDim.x) {
      zq_s[k] = __ldg(P.zq + k);
      ord_s[k] = __ldg(P.order + k);
    }
    __syncthreads();
  }
  const double* __restrict__ zq = q_staged ? zq_s : P.zq;
  const int32_t* __restrict__ order = q_staged ? ord_s : P.order;
  for (unsigned long long e = (unsigned long long)blockIdx.x * n_warps + warp; e < n;
       e += (unsigned long long)gridDim.x * n_warps) {
    int32_t off;
    double mean, stdv;
    if (kFromSums) {  // exact statistics and gate from the re-walked chain sums (K/NormQueryEngine.java:508-511)
      off = P.xin.off[e];
      const bool pass = cnsm_exact_gate(P.xin.ex[e], P.xin.ex2[e], m, P.meanQ, P.stdQ, P.alpha, P.inv_alpha, P.beta, mean, stdv);
      if (!pass) continue;
      if (lane == 0) atomicAdd(P.gate_pass, 1ULL);
    } else {
      off = P.in.off[e];
      mean = P.in.mean[e];
      stdv = P.in.stdv[e];
    }
    const double* __restrict__ wg = P.T + (off - P.first_global);
    const double rstd = 1.0 / stdv;  // x = (w - mean) * rstd: the error does not grow with |mean| / std
    // Tier 2a: the 128 largest-|zQ| terms straight from global memory (fast FMA arithmetic): most entries end here,
    // before their window is staged.
    {
      double part = 0.0;
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int k = u * 32 + lane;
        if (k < m) {
          const double df = (wg[order[k]] - mean) * rstd - zq[k];
          part = __fma_rn(df, df, part);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(kFullMask, part, o);
      if (!(part <= P.eps2_hi)) continue;
    }
    if (staged) {
      __syncwarp();
      for (int k0 = 0; k0 < m; k0 += 256) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) v[u] = (k0 + u * 32 + lane < m) ? wg[k0 + u * 32 + lane] : 0.0;
#pragma unroll
        for (int u = 0; u < 8; u++)
          if (k0 + u * 32 + lane < m) win[k0 + u * 32 + lane] = v[u];
      }
      __syncwarp();
    }
    const double* w = staged ? win : wg;
    // Tier 2b: warp-cooperative fast distance over all m terms, 256 terms per round in |zQ| order, abandoned as soon
    // as the partial sum exceeds eps^2*(1+1e-9).  Only windows that survive every round reach the reference-order
    // summation below.
    {
      double part = 0.0;
      bool over = false;
      for (int k0 = 0; k0 < m && !over; k0 += 256) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int k = k0 + u * 32 + lane;
          if (k < m) {
            const double df = (w[order[k]] - mean) * rstd - zq[k];
            part = __fma_rn(df, df, part);
          }
        }
        double tot = part;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(kFullMask, tot, o);
        over = !(tot <= P.eps2_hi);
      }
      if (over) continue;
    }
    if (lane == 0) atomicAdd(P.n_exact, 1ULL);
    // Tier 3: K/NormQueryEngine.java:513-520 verbatim arithmetic, x = (T[order[k]+j]-mean)/std; dist += (x-zQ[k])^2:
    // all lanes compute the per-term values (each rounded exactly as the reference's), lane 0 adds them in the
    // reference's order.
    double dist = 0.0;
    bool alive = true;
    for (int k0 = 0; k0 < m && alive; k0 += kExactChunk) {
      const int kc = min(kExactChunk, m - k0);
      __syncwarp();
      for (int kb = 0; kb < kc; kb += 256) {
        double wv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int k = kb + u * 32 + lane;
          wv[u] = (k < kc) ? w[order[k0 + k]] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int k = kb + u * 32 + lane;
          if (k < kc) term[k] = xsqdist(xdiv(xsub(wv[u], mean), stdv), zq[k0 + k]);
        }
      }
      __syncwarp();
      if (lane == 0) {
        // The reference leaves its loop at the first partial sum above eps^2 (:516); the terms are non-negative, so
        // testing once per chunk accepts the same windows with the same sums, and the additions form one dependent
        // chain with every load hoisted out of it.
        int k = 0;
        for (; k + 32 <= kc; k += 32) {
#pragma unroll
          for (int u = 0; u < 32; u++) dist = xadd(dist, term[k + u]);
        }
        for (; k < kc; k++) dist = xadd(dist, term[k]);
        alive = dist <= P.eps2;
      }
      alive = __shfl_sync(kFullMask, alive, 0);
    }
    if (lane == 0 && dist <= P.eps2) P.sink.emit(off, xsqrt(dist));
  }
}

}  // namespace kvm
