// index_kernels.cuh — IndexBuilder step 1: sliding-window mean -> toRound key -> equal-key runs.
// Replaces the epoch loop of K/IndexBuilder.java:194-301 and K/utils/MeanIntervalUtils.java:51-61.
//
// The reference restarts its running sum every EPOCH = 100000 buffered samples (epochs overlap by
// w-1), so the mean of every window comes from a sequential chain of at most 1e5 add/subtract steps.
// As in cnsm_kernels.cuh the chain is emulated exactly, one thread per (w, epoch) chain, samples
// delivered through coalesced shared-memory tiles.  The key toRound(mean) depends only on the bucket
// b = floor(2 * fl(fl(ex/w) * 10)); b is taken from one multiply when ex*(20/w) is clear of an
// integer by more than the rounding slack, and from the reference's exact divide/multiply otherwise.
// Each chain emits its maximal equal-bucket segments (b, first, last) into its own slot range of a
// position-indexed buffer; mean_gather_kernel compacts them in loc order and converts b to the key
// with the reference's arithmetic.  Stitching of segments across epoch borders and the split at
// 255 positions (IndexNode.MAXIMUM_DIFF-1, K/IndexBuilder.java:268) are done by the host on the
// compacted list.
#pragma once
#include "cnsm_kernels.cuh"

namespace kvm {

struct MeanChain {
  int32_t begin;   // local 0-based index of the chain's first sample
  int32_t nsamp;   // samples in the chain
  int32_t nwin;    // windows to emit (loc <= n)
  int32_t w;
  long long slot;  // first slot of this chain in the segment buffer
};

struct MeanWalkParams {
  const double* __restrict__ T;
  const MeanChain* __restrict__ chains;
  int n_chains;
  int32_t* seg_b;
  int32_t* seg_first;  // 1-based loc
  int32_t* seg_last;
  int32_t* chain_count;
  int* overflow;  // set when a bucket does not fit int32
};

constexpr int kMeanWarps = 4;
constexpr int kMeanSmemDoublesPerWarp = 2 * 32 * kWalkPitch + 3 * (kFifoDepth * 32) / 2;

__global__ void __launch_bounds__(kMeanWarps * 32) mean_walk_kernel(MeanWalkParams P) {
  extern __shared__ double walk_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* tA = walk_smem + (size_t)warp * kMeanSmemDoublesPerWarp;
  double* tS = tA + 32 * kWalkPitch;
  int32_t* f_b = reinterpret_cast<int32_t*>(tS + 32 * kWalkPitch);  // [depth][32]
  int32_t* f_first = f_b + kFifoDepth * 32;
  int32_t* f_last = f_first + kFifoDepth * 32;

  const int c = (blockIdx.x * kMeanWarps + warp) * 32 + lane;
  MeanChain ch{0, 0, 0, 2, 0};
  if (c < P.n_chains) ch = P.chains[c];
  const int pos = ch.begin, len = ch.nsamp, w = ch.w;
  const int last_s = w - 1 + ch.nwin;  // samples needed: windows end at s = w-1 .. w-2+nwin
  const int need = min(len, last_s);
  const int maxlen = warp_max_i32(need);
  if (maxlen == 0) {
    if (c < P.n_chains) P.chain_count[c] = 0;
    return;
  }
  const double* __restrict__ T = P.T;
  const double dw = (double)w;
  const double c20w = 20.0 / dw;
  int fcnt = 0, emitted = 0;
  double ex = 0.0;
  bool have = false;
  int curb = 0, first = 0;

  auto drain = [&]() {  // lane-private: each chain owns its slot range, no cross-lane ordering needed
    for (int i = 0; i < fcnt; i++) {
      const long long e = ch.slot + emitted + i;
      P.seg_b[e] = f_b[i * 32 + lane];
      P.seg_first[e] = f_first[i * 32 + lane];
      P.seg_last[e] = f_last[i * 32 + lane];
    }
    emitted += fcnt;
    fcnt = 0;
  };

  for (int s0 = 0; s0 < maxlen; s0 += kWalkTile) {
#pragma unroll 8
    for (int r = 0; r < 32; r++) {
      const int p = __shfl_sync(kFullMask, pos, r);
      const int l = __shfl_sync(kFullMask, need, r);
      const int ww = __shfl_sync(kFullMask, w, r);
      const int idx = s0 + lane;
      double a = 0.0, o = 0.0;
      if (idx < l) {
        a = T[p + idx];
        if (idx >= ww - 1) o = T[p + idx - (ww - 1)];
      }
      tA[r * kWalkPitch + lane] = a;
      tS[r * kWalkPitch + lane] = o;
    }
    __syncwarp();
#pragma unroll 4
    for (int i = 0; i < kWalkTile; i++) {
      const int s = s0 + i;
      if (s < need) {
        ex = xadd(ex, tA[lane * kWalkPitch + i]);  // K/IndexBuilder.java:240
        if (s >= w - 1) {
          double v2 = ex * c20w;
          double fl = floor(v2);
          const double frac = v2 - fl;
          const double g = fabs(v2) * 4e-15 + 1e-290;
          if (!(frac >= g && frac <= 1.0 - g)) {
            const double mean = xdiv(ex, dw);                    // :251
            const double v = xmul(mean, 10.0);                   // MeanIntervalUtils.java:52
            fl = floor(xadd(v, v));                              // 2v is exact
          }
          if (!(fabs(fl) < 2147483000.0)) *P.overflow = 1;
          const int b = (int)fl;
          const int loc = ch.begin + s - (w - 1) + 1;  // whole series is resident: local index == global-1
          if (!have || b != curb) {
            if (have) {
              f_b[fcnt * 32 + lane] = curb;
              f_first[fcnt * 32 + lane] = first;
              f_last[fcnt * 32 + lane] = loc - 1;
              fcnt++;
            }
            curb = b;
            first = loc;
            have = true;
          }
          ex = xsub(ex, tS[lane * kWalkPitch + i]);  // :289
        }
      }
      if (fcnt == kFifoDepth) drain();
    }
    __syncwarp();
  }
  if (have) {
    f_b[fcnt * 32 + lane] = curb;
    f_first[fcnt * 32 + lane] = first;
    f_last[fcnt * 32 + lane] = ch.begin + (need - 1) - (w - 1) + 1;
    fcnt++;
  }
  drain();
  if (c < P.n_chains) P.chain_count[c] = emitted;
}

// Exclusive scan of chain_count (single CTA); prefix has n_chains+1 entries.
__global__ void __launch_bounds__(1024) mean_scan_kernel(const int32_t* __restrict__ chain_count, int n_chains,
                                                         long long* __restrict__ prefix) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int r0 = 0; r0 < n_chains; r0 += 1024) {
    const int r = r0 + tid;
    const long long v = (r < n_chains) ? chain_count[r] : 0;
    long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(kFullMask, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const long long wv = s_warp[lane];
      long long wi = wv;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(kFullMask, wi, o);
        if (lane >= o) wi += t;
      }
      s_warp[lane] = wi - wv;
    }
    __syncthreads();
    const long long excl = s_carry + s_warp[warp] + incl - v;
    if (r < n_chains) prefix[r] = excl;
    __syncthreads();
    if (tid == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) prefix[n_chains] = s_carry;
}

// One CTA per chain: copy its segments to the dense, loc-ordered output and turn buckets into keys
// with the reference's arithmetic: ret = floor(v) (+0.5 if v-floor(v) >= 0.5); ret *= 0.1.
__global__ void __launch_bounds__(256) mean_gather_kernel(const MeanChain* __restrict__ chains,
                                                          const int32_t* __restrict__ chain_count,
                                                          const long long* __restrict__ prefix,
                                                          const int32_t* __restrict__ seg_b,
                                                          const int32_t* __restrict__ seg_first,
                                                          const int32_t* __restrict__ seg_last,
                                                          double* __restrict__ out_key, int32_t* __restrict__ out_b,
                                                          int32_t* __restrict__ out_first,
                                                          int32_t* __restrict__ out_last) {
  const int c = blockIdx.x;
  const int cnt = chain_count[c];
  const long long src = chains[c].slot, dst = prefix[c];
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const int b = seg_b[src + i];
    const double int_value = floor((double)b * 0.5);
    const double ret = (b & 1) ? xadd(int_value, 0.5) : int_value;
    out_key[dst + i] = xmul(ret, 0.1);
    out_b[dst + i] = b;
    out_first[dst + i] = seg_first[src + i];
    out_last[dst + i] = seg_last[src + i];
  }
}

__global__ void bswap64_kernel(unsigned long long* __restrict__ p, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long v = p[i];
    const unsigned lo = (unsigned)v, hi = (unsigned)(v >> 32);
    p[i] = ((unsigned long long)__byte_perm(lo, 0, 0x0123) << 32) | (unsigned long long)__byte_perm(hi, 0, 0x0123);
  }
}

}  // namespace kvm
