// index_kernels.cuh — IndexBuilder step 1: sliding-window mean -> toRound key -> equal-key runs.
// Replaces the epoch loop of K/IndexBuilder.java:194-301 and K/utils/MeanIntervalUtils.java:51-61.
//
// The reference restarts its running sum every EPOCH = 100000 buffered samples (epochs overlap by
// w-1), so the mean of every window comes from a sequential chain of at most 1e5 add/subtract steps.
// The chain is emulated exactly by cnsm_walk_kernel<4, kDelta, 1> (cnsm_kernels.cuh), one thread per (w, epoch)
// chain; its gate warps turn every window's chain sum into the key bucket b = floor(2 * fl(fl(ex/w) * 10)) and write
// b[window].  The kernels below run-length encode that array fully in parallel (count / scan / emit); only the split
// at 255 positions (IndexNode.MAXIMUM_DIFF-1, K/IndexBuilder.java:268) is left to the host.
#pragma once
#include "cnsm_kernels.cuh"

namespace kvm {

constexpr int kRleTile = 4096;  // windows per CTA in the run-length pass

// Run boundaries of the per-window bucket array: window i starts a run iff i == 0 or b[i] != b[i-1].  The
// reference's run state persists across its epochs (K/IndexBuilder.java:190-192), so this global comparison is
// exactly its `lastMeanRound.equals(curMeanRound)` test; the split at 255 positions is applied afterwards.
__global__ void __launch_bounds__(256) rle_count_kernel(const int32_t* __restrict__ b, long long n_win,
                                                        int32_t* __restrict__ tile_count) {
  __shared__ int s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  const long long base = (long long)blockIdx.x * kRleTile;
  int c = 0;
  for (int i = threadIdx.x; i < kRleTile; i += 256) {
    const long long g = base + i;
    if (g < n_win && (g == 0 || b[g] != b[g - 1])) c++;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFullMask, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&s_cnt, c);
  __syncthreads();
  if (threadIdx.x == 0) tile_count[blockIdx.x] = s_cnt;
}

// Exclusive scan of tile_count (single CTA); prefix has n+1 entries.
__global__ void __launch_bounds__(1024) rle_scan_kernel(const int32_t* __restrict__ tile_count, int n,
                                                        long long* __restrict__ prefix) {
  __shared__ long long s_warp[32];
  __shared__ long long s_carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_carry = 0;
  __syncthreads();
  for (int r0 = 0; r0 < n; r0 += 1024) {
    const int r = r0 + tid;
    const long long v = (r < n) ? tile_count[r] : 0;
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
    if (r < n) prefix[r] = excl;
    __syncthreads();
    if (tid == 1023) s_carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) prefix[n] = s_carry;
}

// Writes the run starts of each tile, in order, at its scanned offset: start position (0-based window index)
// and the key rebuilt from the bucket with the reference's arithmetic: ret = floor(v) (+0.5 if the half bit is
// set); ret *= 0.1 (MeanIntervalUtils.java:53-59).
__global__ void __launch_bounds__(256) rle_emit_kernel(const int32_t* __restrict__ b, long long n_win,
                                                       const long long* __restrict__ prefix,
                                                       int32_t* __restrict__ run_start, double* __restrict__ run_key) {
  __shared__ int s_warp_base[8];
  const long long base = (long long)blockIdx.x * kRleTile;
  long long out = prefix[blockIdx.x];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i0 = 0; i0 < kRleTile; i0 += 256) {  // 256 consecutive windows per round keep the output ordered
    const long long g = base + i0 + threadIdx.x;
    const bool head = g < n_win && (g == 0 || b[g] != b[g - 1]);
    const unsigned bal = __ballot_sync(kFullMask, head);
    if (lane == 0) s_warp_base[warp] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; w++) {
      const int c = s_warp_base[w];
      before += (w < warp) ? c : 0;
      total += c;
    }
    if (head) {
      const long long slot = out + before + __popc(bal & ((1u << lane) - 1u));
      const int bv = b[g];
      const double int_value = floor((double)bv * 0.5);
      const double ret = (bv & 1) ? xadd(int_value, 0.5) : int_value;
      run_start[slot] = (int32_t)g;
      run_key[slot] = xmul(ret, 0.1);
    }
    out += total;
    __syncthreads();
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
