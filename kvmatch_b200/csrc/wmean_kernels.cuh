// wmean_kernels.cuh — IndexBuilder step 1 for all windows of Sigma in ONE streaming pass.
// Replaces the five sequential passes of K/IndexBuilder.java:98-120 ("TODO: naive") over the epoch loop of
// K/IndexBuilder.java:194-301 and K/utils/MeanIntervalUtils.java:51-61.
//
// Per window width w the reference walks EPOCH = 100000-sample buffers with a running add/subtract chain and turns
// every window mean into the key toRound(mean) = (floor(10 mean) [+ 0.5]) * 0.1, i.e. the bucket
//     b = floor(2 * fl(fl(ex / w) * 10)).
// Like the cNSM statistics (stream_kernels.cuh) the chain's exact value is only needed where it can change the result:
// when v2 = ex * 20/w lies within the chain's provable rounding drift g of an integer.  So:
//
//   wmean_stream_kernel  one CTA per tile of 33*NT window positions.  The tile's samples arrive once (TMA bulk copy),
//                        33-sample group sums are formed once, and for each width a warp-local windowed sum / shuffle
//                        scan gives every thread the sum of its first window.  The thread then slides over its 33
//                        positions for all widths together (the outgoing sample is shared).  Per width and position
//                        two FMAs with round-down against 1.5*2^52 give floor(v2 - g) and floor(v2 + g): equal -> the
//                        bucket, different -> the window is AMBIGUOUS (flagged for the exact re-walk, emitted as a
//                        run of its own with bucket kAmbiguous).  Pass 1 records the run starts (a thread always
//                        starts a run at its first position: the host merges equal neighbours anyway), a warp scan
//                        and ONE atomic per warp and width reserve the warp's slice of the run list, pass 2 replays
//                        the slide (2 FP64 per step) and writes (position, bucket) at the starts.
//   chain_rewalk_kernel  (stream_kernels.cuh) walks the (width, epoch) chains that hold an ambiguous window exactly and
//                        returns their sums; the host computes those few buckets with the reference's arithmetic,
//                        stitches the warps' slices in position order, merges equal neighbours and splits at 255.
#pragma once
#include "stream_kernels.cuh"

namespace kvm {

constexpr int kMaxWidths = 5;
constexpr int kAmbiguous = INT32_MIN;  // bucket of a run whose key must come from the exact re-walk

struct WmeanWidth {
  int w;
  int n_win;            // window positions 0 .. n_win-1 (the reference's loc = position + 1)
  double c20w;          // 20 / w
  double cd1;           // |ex_stream - ex_true| <= cd1 * A   (A = max |sample| the chain has seen)
  double cd_chain;      // |ex_chain - ex_true| <= cd_chain * (samples the chain has consumed) * A
  int2* runs;           // (first position, bucket), position order within a warp's slice
  long long run_cap;
  unsigned long long* n_runs;    // cursor over `runs`
  int* seg_off;         // per (tile, warp): first slot of the slice
  int* seg_cnt;
  unsigned* need_bits;  // exact re-walk (regular grid of epochs: chunk = EPOCH - w + 1)
  int32_t* chain_last;
  int32_t* flagged;
  unsigned long long* n_flagged;
};

struct WmeanParams {
  const double* __restrict__ T;
  int n_widths;
  int w_max;
  int total_pos;        // max n_win over the widths
  int epoch;            // 100000
  const double* __restrict__ bmax;
  int n_bmax;
  int* overflow;        // set when v2 does not fit the 31-bit bucket range or the guard exceeds 1/4
  WmeanWidth W[kMaxWidths];
};

constexpr size_t wmean_smem_bytes(int nt, int w_max) {
  return 16 + sizeof(double) * (stream_xs_doubles(nt, w_max) + stream_gs_doubles(nt, w_max)) + 16;
}

template <int NT>
__global__ void __launch_bounds__(NT, 3) wmean_stream_kernel(WmeanParams P) {
  constexpr int NW = kMaxWidths;
  extern __shared__ __align__(16) unsigned char wm_smem[];
  __shared__ double s_amax;
  __shared__ int s_wtot[kMaxWidths][NT / 32];
  __shared__ int s_cta_tot[kMaxWidths];
  __shared__ long long s_cta_base[kMaxWidths];
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(kFullMask, tid >> 5, 0);
  constexpr int kW = kGroup * NT;
  const int i0 = (int)blockIdx.x * kW;  // first window position of the tile = local index of its first sample
  const int npos = min(kW, P.total_pos - i0);
  const int w_max = P.w_max;
  const int ns = npos + w_max - 1;
  const int s0a = i0 & ~1;
  const int lead = i0 - s0a;
  const int nload = (lead + ns + 1) & ~1;
  const int ng = (ns + kGroup - 1) / kGroup;
  const size_t n_xs = stream_xs_doubles(NT, w_max);
  const int gs_cap = (int)stream_gs_doubles(NT, w_max);
  double* xs_raw = reinterpret_cast<double*>(wm_smem + 16);
  double* xs = xs_raw + lead;  // xs[k] = sample i0 + k
  double* gs1 = xs_raw + n_xs;
  const uint32_t bar = smem_u32(wm_smem);

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
  // A = max |sample| over everything an epoch chain can have summed when it reaches this tile
  if (warp == NT / 32 - 1) {
    const int bm_lo = max(0, i0 - P.epoch) / kBmaxBlock;
    const int bm_hi = min(P.n_bmax - 1, (i0 + ns) / kBmaxBlock);
    unsigned long long mx = 0ULL;
    for (int b = bm_lo + lane; b <= bm_hi; b += 32) mx = max(mx, (unsigned long long)__double_as_longlong(__ldg(P.bmax + b)));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(kFullMask, mx, o));
    if (lane == 0) s_amax = __longlong_as_double((long long)mx);
  }
  for (int i = nload + tid; i < lead + ns + kGroup + 2; i += NT) xs_raw[i] = 0.0;
  __syncthreads();  // (1) mbarrier initialised; zero tail and the tile's amplitude staged
  mbar_wait(bar, 0);
  // ---- sums of 33-sample groups (shared by all widths)
  for (int g = tid; g < gs_cap; g += NT) {
    double a[3] = {0.0, 0.0, 0.0};
    if (g < ng) {
      const double* __restrict__ x = xs + g * kGroup;
#pragma unroll
      for (int j = 0; j < kGroup / 3; j++) {
#pragma unroll
        for (int u = 0; u < 3; u++) a[u] += x[3 * j + u];
      }
    }
    gs1[g] = (a[0] + a[1]) + a[2];
  }
  __syncthreads();  // (2) group sums ready
  // (a warp without positions — the last tile — keeps running with zero valid windows: it takes part in the barriers
  // of the slice reservation below)

  const double A = s_amax;
  const int p0 = tid * kGroup;  // this thread's first position within the tile
  const double magic = 6755399441055744.0;  // 1.5 * 2^52: RD(x*c + magic) carries floor(x*c) in its low word
  // ---- per width: the sum of this thread's first window and the guard, in ex units
  double ex0[NW], gsh[NW], c20[NW];
  int wq[NW];
#pragma unroll
  for (int q = 0; q < NW; q++) {
    ex0[q] = gsh[q] = c20[q] = 0.0;
    wq[q] = 1;
    if (q < P.n_widths) {
      const WmeanWidth& Wq = P.W[q];
      const int w = Wq.w;
      wq[q] = w;
      c20[q] = Wq.c20w;
      // guard on v2 = ex*20/w: the chain's drift and the stream's own rounding (cd1*A), the reference's two roundings
      // (divide by w, multiply by 10) and ours (20/w, ex -+ shift), relative to |v2| <= 20 A
      // (cd_chain: per chain position, the epoch chain has done 2 roundings of at most u*w*A each; a tile that crosses
      // an epoch boundary takes the full epoch)
      const int chunk = P.epoch - w + 1;
      const int c_lo = i0 % chunk;
      const double pos_hi = (c_lo + npos <= chunk) ? (double)(c_lo + npos + w) : (double)P.epoch;
      const double g = (Wq.cd_chain * pos_hi + Wq.cd1) * A * Wq.c20w * (1.0 + 1e-12) + 16.0 * 1.1102230246251565e-16 * (20.0 * A) + 1e-300;
      if ((!(g < 0.25) || !(20.0 * A < 2.0e9)) && tid == 0) *P.overflow = 1;  // the host falls back to the exact walker
      gsh[q] = g / Wq.c20w * (1.0 + 1e-12);
      const int qa = w / kGroup, qb = w - qa * kGroup;
      const int gw = 32 * warp;
      double W1 = 0.0;
      for (int k = lane; k < qa; k += 32) W1 += gs1[gw + k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) W1 += __shfl_xor_sync(kFullMask, W1, o);
      const double d1 = gs1[tid + qa] - gs1[tid];
      double i1 = d1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double u1 = __shfl_up_sync(kFullMask, i1, o);
        if (lane >= o) i1 += u1;
      }
      double e = W1 + (i1 - d1);
      const double* __restrict__ x = xs + (tid + qa) * kGroup;
      for (int j = 0; j < qb; j++) e += x[j];
      ex0[q] = e;
    }
  }
  const double* __restrict__ xo = xs + p0;
  const int np = max(0, min(kGroup, npos - p0));

  // ---- pass 1: run starts and ambiguous windows
  unsigned long long starts[NW], amb[NW];
  {
    double e[NW];
    int prevb[NW];
#pragma unroll
    for (int q = 0; q < NW; q++) {
      e[q] = ex0[q];
      prevb[q] = 0;
      starts[q] = 1ULL;  // a thread always starts a run at its first position
      amb[q] = 0ULL;
    }
#pragma unroll 3
    for (int j = 0; j < kGroup; j++) {
      const double o = xo[j];
      const unsigned long long bit = 1ULL << j;
#pragma unroll
      for (int q = 0; q < NW; q++) {
        const int bl = __double2loint(__fma_rd(e[q] - gsh[q], c20[q], magic));
        const int bh = __double2loint(__fma_rd(e[q] + gsh[q], c20[q], magic));
        const bool a = bl != bh;
        const int cur = a ? kAmbiguous : bl;
        if (a | (cur != prevb[q])) starts[q] |= bit;
        if (a) amb[q] |= bit;
        prevb[q] = cur;
        e[q] += xo[j + wq[q]] - o;
      }
    }
  }
  // ---- per width: validity and run counts
  int slot[NW], incl_q[NW], cnt_q[NW];
#pragma unroll
  for (int q = 0; q < NW; q++) {
    slot[q] = 0;
    if (q >= P.n_widths) {
      starts[q] = 0ULL;
      continue;
    }
    const WmeanWidth& Wq = P.W[q];
    const int valid = max(0, min(np, Wq.n_win - (i0 + p0)));  // this thread's positions that are windows of this width
    const unsigned long long vm = (1ULL << valid) - 1ULL;    // (valid <= 33)
    starts[q] &= vm;
    amb[q] &= vm;
    const int cnt = __popcll(starts[q]);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFullMask, incl, o);
      if (lane >= o) incl += t;
    }
    incl_q[q] = incl;
    cnt_q[q] = cnt;
    if (lane == 31) s_wtot[q][warp] = incl;  // the warp's runs of this width
  }
  // ---- slice reservation: ONE atomic per CTA and width (thread q), the warps' slices follow each other in warp order.
  // (One returning atomic per warp and width put 95 k serialised updates on each of five addresses at n = 1e8 and
  // every warp waited for its turn.)
  __syncthreads();
  if (tid < P.n_widths) {
    int tot = 0;
#pragma unroll
    for (int w8 = 0; w8 < NT / 32; w8++) {
      const int c = s_wtot[tid][w8];
      s_wtot[tid][w8] = tot;  // exclusive prefix over the warps
      tot += c;
    }
    s_cta_tot[tid] = tot;
    s_cta_base[tid] = tot > 0 ? (long long)atomicAdd(P.W[tid].n_runs, (unsigned long long)tot) : 0LL;
  }
  __syncthreads();
#pragma unroll
  for (int q = 0; q < NW; q++) {
    if (q >= P.n_widths) continue;
    const WmeanWidth& Wq = P.W[q];
    {
      const long long base = s_cta_base[q] + s_wtot[q][warp];
      const int total = __shfl_sync(kFullMask, incl_q[q], 31);
      const int seg = (int)blockIdx.x * (NT / 32) + warp;
      if (lane == 0) {
        Wq.seg_off[seg] = (int)base;
        Wq.seg_cnt[seg] = total;
      }
      if (s_cta_base[q] + s_cta_tot[q] > Wq.run_cap) starts[q] = 0ULL;  // overflow: counted, not stored (the host grows the list and re-runs)
      slot[q] = (int)base + (incl_q[q] - cnt_q[q]);
    }
    unsigned long long m2 = amb[q];
    if (m2) {  // flag the ambiguous windows for the exact re-walk
      const int chunk = P.epoch - Wq.w + 1;
      while (m2) {
        const int j = __ffsll((long long)m2) - 1;
        m2 &= m2 - 1;
        const unsigned v = (unsigned)(i0 + p0 + j);
        const int p = (int)(v / (unsigned)chunk);
        atomicOr(Wq.need_bits + (v >> 5), 1u << (v & 31));
        const int old = atomicMax(Wq.chain_last + p, (int)v - p * chunk);
        if (old < 0) {
          const unsigned long long s = atomicAdd(Wq.n_flagged, 1ULL);
          Wq.flagged[s] = p;
        }
      }
    }
  }
  // ---- pass 2: replay the slide (same operations, same sums) and write (position, bucket) at the run starts
  {
    double e[NW];
#pragma unroll
    for (int q = 0; q < NW; q++) e[q] = ex0[q];
    unsigned long long any = 0ULL;
#pragma unroll
    for (int q = 0; q < NW; q++) any |= starts[q];
#pragma unroll 3
    for (int j = 0; j < kGroup; j++) {
      const double o = xo[j];
      if ((any >> j) & 1ULL) {
#pragma unroll
        for (int q = 0; q < NW; q++) {
          if ((starts[q] >> j) & 1ULL) {
            const int b = ((amb[q] >> j) & 1ULL) ? kAmbiguous : __double2loint(__fma_rd(e[q] - gsh[q], c20[q], magic));
            P.W[q].runs[slot[q]++] = make_int2(i0 + p0 + j, b);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < NW; q++) e[q] += xo[j + wq[q]] - o;
    }
  }
}

}  // namespace kvm
