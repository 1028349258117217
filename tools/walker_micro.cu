// Micro-benchmark of the chain-walker inner loop variants (cycles per window position, one producer warp per SM).
// A: chain from registers only; B: + LDS.128 tile reads (ping-pong prefetch); C: + STS.128 staging;
// D: + named-barrier hand-off with a consumer warp that only syncs; E: D with consumer reading the block.
#include <cstdio>
#include <cuda_runtime.h>
#define PITCH 34
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile("{\n.reg .pred p;\nWL:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra WD;\nbra WL;\nWD:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void st_async_v2(unsigned addr, double x, double y, unsigned bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(addr), "d"(x), "d"(y), "r"(bar) : "memory");
}

// MODE 5: staging by st.async + transaction mbarriers (full), plain mbarriers (empty); consumer reads the block
__global__ void __cluster_dims__(1, 1, 1) __launch_bounds__(64) k5(double* out, long long* cyc, int nblocks) {
  __shared__ __align__(16) double tileA[32 * PITCH], tileO[32 * PITCH];
  __shared__ __align__(16) double2 stage[2][32 * 17];
  __shared__ __align__(8) unsigned long long bars[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * PITCH; i += 64) { tileA[i] = 1.0 + i * 1e-3; tileO[i] = 0.5 + i * 1e-3; }
  const unsigned full0 = smem_u32(bars), empty0 = full0 + 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; s++) { mbar_init(full0 + 8 * s, 32); mbar_init(empty0 + 8 * s, 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    double ex = 0, ex2 = 0;
    const double* ra = tileA + lane * PITCH;
    const double* ro = tileO + lane * PITCH;
    double2 A0[8], O0[8], A1[8], O1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { A0[i] = *(const double2*)(ra + 2 * i); O0[i] = *(const double2*)(ro + 2 * i); A1[i] = A0[i]; O1[i] = O0[i]; }
    long long t0 = clock64();
    for (int b = 0; b < nblocks; b += 2) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        double2(&Av)[8] = h ? A1 : A0; double2(&Ov)[8] = h ? O1 : O0;
        double2(&An)[8] = h ? A0 : A1; double2(&On)[8] = h ? O0 : O1;
        const int slot = (b + h) & 3;
        if (b + h >= 2) mbar_wait(empty0 + 8 * (slot & 1), (unsigned)((((b + h) >> 1) - 1) & 1));
        const unsigned st = smem_u32(&stage[slot & 1][lane * 17]);
        const unsigned fb = full0 + 8 * (slot & 1);
#pragma unroll
        for (int i = 0; i < 8; i++) {
          An[i] = *(const double2*)(ra + 16 * (1 - h) + 2 * i); On[i] = *(const double2*)(ro + 16 * (1 - h) + 2 * i);
          const double a0 = Av[i].x, a1 = Av[i].y, o0 = Ov[i].x, o1 = Ov[i].y;
          ex = __dadd_rn(ex, a0); ex2 = __dadd_rn(ex2, __dmul_rn(a0, a0));
          st_async_v2(st + 32 * i, ex, ex2, fb);
          ex = __dsub_rn(ex, o0); ex2 = __dsub_rn(ex2, __dmul_rn(o0, o0));
          ex = __dadd_rn(ex, a1); ex2 = __dadd_rn(ex2, __dmul_rn(a1, a1));
          st_async_v2(st + 32 * i + 16, ex, ex2, fb);
          ex = __dsub_rn(ex, o1); ex2 = __dsub_rn(ex2, __dmul_rn(o1, o1));
        }
      }
    }
    long long t1 = clock64();
    out[blockIdx.x * 32 + lane] = ex + ex2;
    if (lane == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  } else {
    double acc = 0;
    for (int b = 0; b < nblocks; b++) {
      const int slot = b & 1;
      mbar_expect_tx(full0 + 8 * slot, 16 * 16);   // this lane's row: 16 positions x 16 bytes (its own arrival)
      mbar_wait(full0 + 8 * slot, (unsigned)((b >> 1) & 1));
#pragma unroll
      for (int c = 0; c < 16; c++) acc += stage[slot][lane * 17 + c].x;
      mbar_arrive(empty0 + 8 * slot);
    }
    out[4096 + blockIdx.x * 32 + lane] = acc;
  }
}


__device__ __forceinline__ unsigned mbar_test(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok;
}
// MODE 6: plain STS staging; full/empty hand-off by mbarriers only; the producer polls the NEXT slot's empty barrier
// one block early (test_wait result consumed a block later), so no sync latency sits in front of the FP64 chain.
template <int NCONS>
__global__ void __launch_bounds__(32 + 32 * NCONS) k6(double* out, long long* cyc, int nblocks) {
  __shared__ __align__(16) double tileA[32 * PITCH], tileO[32 * PITCH];
  __shared__ __align__(16) double2 stage[4][32 * 17 / 2];   // (half-size rows: only 8 positions staged per slot to fit static smem)
  __shared__ __align__(8) unsigned long long bars[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * PITCH; i += blockDim.x) { tileA[i] = 1.0 + i * 1e-3; tileO[i] = 0.5 + i * 1e-3; }
  const unsigned full0 = smem_u32(bars), empty0 = full0 + 32;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; s++) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    double ex = 0, ex2 = 0;
    const double* ra = tileA + lane * PITCH;
    const double* ro = tileO + lane * PITCH;
    double2 A0[8], O0[8], A1[8], O1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { A0[i] = *(const double2*)(ra + 2 * i); O0[i] = *(const double2*)(ro + 2 * i); A1[i] = A0[i]; O1[i] = O0[i]; }
    unsigned next_ok = 1;
    long long t0 = clock64();
    for (int b = 0; b < nblocks; b += 2) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        double2(&Av)[8] = h ? A1 : A0; double2(&Ov)[8] = h ? O1 : O0;
        double2(&An)[8] = h ? A0 : A1; double2(&On)[8] = h ? O0 : O1;
        const int j = b + h, slot = j & 3;
        if (!next_ok) mbar_wait(empty0 + 8 * slot, (unsigned)(((j >> 2) - 1) & 1));   // rarely taken
        // poll the next block's slot now; the answer is needed only after this block's chain
        const int jn = j + 1;
        unsigned ok_n = 1;
        if (jn >= 4) ok_n = mbar_test(empty0 + 8 * (jn & 3), (unsigned)(((jn >> 2) - 1) & 1));
        double2* st = &stage[slot][lane * 8];
        double2 p0 = make_double2(0, 0), p1 = p0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          An[i] = *(const double2*)(ra + 16 * (1 - h) + 2 * i); On[i] = *(const double2*)(ro + 16 * (1 - h) + 2 * i);
          const double a0 = Av[i].x, a1 = Av[i].y, o0 = Ov[i].x, o1 = Ov[i].y;
          ex = __dadd_rn(ex, a0); ex2 = __dadd_rn(ex2, __dmul_rn(a0, a0));
          const double2 q0 = make_double2(ex, ex2);
          ex = __dsub_rn(ex, o0); ex2 = __dsub_rn(ex2, __dmul_rn(o0, o0));
          ex = __dadd_rn(ex, a1); ex2 = __dadd_rn(ex2, __dmul_rn(a1, a1));
          const double2 q1 = make_double2(ex, ex2);
          ex = __dsub_rn(ex, o1); ex2 = __dsub_rn(ex2, __dmul_rn(o1, o1));
          if (i > 0 && i < 5) { st[2 * i - 2] = p0; st[2 * i - 1] = p1; }
          p0 = q0; p1 = q1;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(full0 + 8 * slot);
        next_ok = ok_n;
      }
    }
    long long t1 = clock64();
    out[blockIdx.x * 32 + lane] = ex + ex2;
    if (lane == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  } else {
    double acc = 0;
    for (int b = warp - 1; b < nblocks; b += NCONS) {
      const int slot = b & 3;
      mbar_wait(full0 + 8 * slot, (unsigned)((b >> 2) & 1));
#pragma unroll
      for (int c = 0; c < 8; c++) acc += stage[slot][lane * 8 + c].x;
      __syncwarp();
      if (lane == 0) mbar_arrive(empty0 + 8 * slot);
    }
    out[4096 + blockIdx.x * 96 + threadIdx.x] = acc;
  }
}

template <int MODE>
__global__ void __launch_bounds__(64) k(double* out, long long* cyc, int nblocks) {
  __shared__ __align__(16) double tileA[32 * PITCH], tileO[32 * PITCH];
  __shared__ __align__(16) double2 stage[2][32 * 17];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 32 * PITCH; i += 64) { tileA[i] = 1.0 + i * 1e-3; tileO[i] = 0.5 + i * 1e-3; }
  __syncthreads();
  if (warp == 0) {
    double ex = 0, ex2 = 0;
    const double* ra = tileA + lane * PITCH;
    const double* ro = tileO + lane * PITCH;
    double2 A0[8], O0[8], A1[8], O1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { A0[i] = *(const double2*)(ra + 2 * i); O0[i] = *(const double2*)(ro + 2 * i); A1[i] = A0[i]; O1[i] = O0[i]; }
    long long t0 = clock64();
    for (int b = 0; b < nblocks; b += 2) {
#pragma unroll
      for (int h = 0; h < 2; h++) {
        double2(&Av)[8] = h ? A1 : A0; double2(&Ov)[8] = h ? O1 : O0;
        double2(&An)[8] = h ? A0 : A1; double2(&On)[8] = h ? O0 : O1;
        if (MODE >= 1) {
#pragma unroll
          for (int i = 0; i < 8; i++) { An[i] = *(const double2*)(ra + 16 * (1 - h) + 2 * i); On[i] = *(const double2*)(ro + 16 * (1 - h) + 2 * i); }
        }
        const int slot = (b + h) & 3;
        if (MODE >= 3 && b + h >= 4) bar_sync(5 + slot, 64);
        double2* st = &stage[slot & 1][lane * 17];
        double2 p0 = make_double2(0, 0), p1 = p0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const double a0 = Av[i].x, a1 = Av[i].y, o0 = Ov[i].x, o1 = Ov[i].y;
          ex = __dadd_rn(ex, a0); ex2 = __dadd_rn(ex2, __dmul_rn(a0, a0));
          const double2 q0 = make_double2(ex, ex2);
          ex = __dsub_rn(ex, o0); ex2 = __dsub_rn(ex2, __dmul_rn(o0, o0));
          ex = __dadd_rn(ex, a1); ex2 = __dadd_rn(ex2, __dmul_rn(a1, a1));
          const double2 q1 = make_double2(ex, ex2);
          ex = __dsub_rn(ex, o1); ex2 = __dsub_rn(ex2, __dmul_rn(o1, o1));
          if (MODE >= 2 && i > 0) { st[2 * i - 2] = p0; st[2 * i - 1] = p1; }
          p0 = q0; p1 = q1;
        }
        if (MODE >= 2) { st[14] = p0; st[15] = p1; }
        if (MODE >= 3) bar_arrive(1 + slot, 64);
      }
    }
    long long t1 = clock64();
    out[blockIdx.x * 32 + lane] = ex + ex2;
    if (lane == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  } else if (MODE >= 3) {
    double acc = 0;
    for (int b = 0; b < nblocks; b++) {
      const int slot = b & 3;
      bar_sync(1 + slot, 64);
      if (MODE >= 4) {
#pragma unroll
        for (int c = 0; c < 16; c++) acc += stage[slot & 1][lane * 17 + c].x;
      }
      bar_arrive(5 + slot, 64);
    }
    out[4096 + blockIdx.x * 32 + lane] = acc;
  }
}
int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 8 * 8192 * 2); cudaMalloc(&cyc, 8);
  const int nb = 20000;
#define RUN(M, G) k<M><<<G, 64>>>(out, cyc, nb); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
  printf("mode %d grid %3d: %.2f cycles/step  (%s)\n", M, G, (double)h / (nb * 16.0), cudaGetErrorString(cudaGetLastError()));
  RUN(0, 1) RUN(2, 1) RUN(3, 1) RUN(4, 1)
  k5<<<1, 64>>>(out, cyc, nb); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("mode 5 (st.async + mbarrier): %.2f cycles/step (%s) check %f\n", (double)h / (nb * 16.0), cudaGetErrorString(cudaGetLastError()), 0.0);
  k6<1><<<1, 64>>>(out, cyc, nb); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("mode 6 (mbarrier hand-off, 1 consumer): %.2f cycles/step (%s)\n", (double)h / (nb * 16.0), cudaGetErrorString(cudaGetLastError()));
  k6<2><<<1, 96>>>(out, cyc, nb); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("mode 6 (mbarrier hand-off, 2 consumers): %.2f cycles/step (%s)\n", (double)h / (nb * 16.0), cudaGetErrorString(cudaGetLastError()));
  k6<2><<<296, 96>>>(out, cyc, nb); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("mode 6 (2 consumers, grid 296): %.2f cycles/step (%s)\n", (double)h / (nb * 16.0), cudaGetErrorString(cudaGetLastError()));
  { double hv[64]; cudaMemcpy(hv, out + 4096, 8 * 32, cudaMemcpyDeviceToHost); printf("consumer acc lane0 %.6e\n", hv[0]); }
  return 0;
}
