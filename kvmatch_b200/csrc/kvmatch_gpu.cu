// kvmatch_gpu.cu — C ABI (include/kvmatch_gpu.h) and host orchestration of libkvmatch_gpu.so.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -shared ...
// (see Makefile).  -fmad=false plus the __d*_rn intrinsics in the exact paths keep every reported
// value free of FMA contraction; the fast paths request FMA explicitly with __fma_rn.
//
// There is NO CPU fallback in this file: every verification result is produced by the kernels in
// ed_kernels.cuh / cnsm_kernels.cuh / dtw_kernels.cuh / index_kernels.cuh.  The host side only
// clamps intervals, prepares the query (statistics, z-normalisation, |z| ordering, envelope — all
// O(m), as the reference does before its phase-2 loop), and sorts the sparse answers by offset.
#include "../../include/kvmatch_gpu.h"

#include <algorithm>
#include <dlfcn.h>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "dtw_kernels.cuh"
#include "stream_kernels.cuh"
#include "wmean_kernels.cuh"
#include "index_kernels.cuh"
#include "index_file.hpp"
#include "phase1.hpp"

using namespace kvm;

struct kvm_ctx;
int launch_dtw(kvm_ctx* ctx, const DtwParams& D);

namespace {

std::string g_create_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      e = cudaMalloc(&p, bytes);
      want = bytes;
    }
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 4 + 4096;
    cudaError_t e = cudaMallocHost(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

// Packs small host arrays into one pinned staging block that is uploaded with a single H2D copy.
struct Arena {
  std::vector<unsigned char> host;
  size_t add(const void* src, size_t bytes) {
    size_t off = (host.size() + 255) & ~size_t(255);
    host.resize(off + bytes);
    if (bytes) std::memcpy(host.data() + off, src, bytes);
    return off;
  }
};

enum Counter { kCntAnswers = 0, kCntCand = 1, kCntGate = 2, kCntTiles = 3, kCntEntries = 4, kCntFlag = 5, kCntDone = 6, kCntCand2 = 7, kCntCells = 8, kCntProbe = 9, kCntProbe2 = 10, kNumCounters = 12 };

struct Plan {
  std::vector<int32_t> cbegin, nsamp, ncand;
  int64_t cnt_candidate = 0, V = 0, S = 0;
};

// The cNSM engines' interval plan of the previous call, still resident in ctx->arena.  A scan that reuses its
// interval list (an index-free scan over a fixed chain grid, or the same candidate set under another query)
// skips re-planning and the upload.  Any other upload into the arena, and any series load, invalidates it.
struct NormPlanCache {
  bool valid = false;
  std::vector<int32_t> lr;
  int K = 0, shift = 0, m = 0, n_regions = 0;
  Plan plan;
  size_t o_cbegin = 0, o_nsamp = 0, o_rbase = 0;
};

// Host plan of the streaming statistics pass (stream_kernels.cuh): live chains, segments of adjacent chains cut into
// tiles.  Cached like NormPlanCache (same key) while it stays resident in ctx->sarena.
struct StreamPlan {
  bool valid = false;
  std::vector<int32_t> lr;
  int K = 0, shift = 0, m = 0, nt = 0;
  int n_chains = 0, n_tiles = 0;
  int n_segments = 0, s_base = 0;
  int regular = 0, chunk = 0;  // one segment cut into chains of `chunk` windows (the last one may be shorter)
  bool unchecked = false;      // regular by its first and last intervals only: the full pass over the list is still due
  int64_t V = 0, S = 0, cnt_candidate = 0, l_max = 0;
  size_t o_tiles = 0, o_cb = 0, o_nc = 0, o_vb = 0, bytes = 0;
};

// Per-query device buffers of a query set (kvm_verify_cnsm_ed_batch): swapped into the ctx's single-query slots while
// that query's evaluator / exact stages run, so those stages are the single-query code.
struct BatchSlot {
  DevBuf wl_off, wl_ex, wl_ex2, region_count, tile_prefix, counters, qarena, cand_off, cand_mean, cand_std, ans_off, ans_dist;
  PinBuf h_counters, h_off, h_dist;
  bool eager_valid = false;
  long long cand_cap = 0, ans_cap = 0;
  std::vector<int32_t> off;
  std::vector<double> dist;
  void release() {
    h_counters.release();
    h_off.release();
    h_dist.release();
    DevBuf* all[] = {&wl_off, &wl_ex, &wl_ex2, &region_count, &tile_prefix, &counters, &qarena, &cand_off, &cand_mean,
                     &cand_std, &ans_off, &ans_dist};
    for (DevBuf* b : all) b->release();
    cand_cap = ans_cap = 0;
  }
};
}  // namespace

struct kvm_ctx {
  int device = 0;
  int n_sms = 148;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t evs[4] = {nullptr, nullptr, nullptr, nullptr};  // stage boundaries inside a call ([2],[3]: query sets)
  std::string err;

  DevBuf series_buf;  // [kFrontPad zeros | samples | kTailPad zeros]
  double* series = nullptr;  // sample 0
  int64_t n = 0, first = 0, count = 0;  // global length; 1-based offset of series[0]; samples held

  DevBuf arena, qarena, counters, wl_off, wl_ex, wl_ex2, region_count, tile_prefix;
  DevBuf cand_off, cand_mean, cand_std, ans_off, ans_dist;
  DevBuf cand_lb;  // Keogh totals beside cand_* (between the two probe stages)
  int dtw_final_counter = 9;  // counter that holds the candidates handed to the band DTW (kCntProbe / kCntProbe2)
  DevBuf cand2_off, cand2_mean, cand2_std, cand2_lb;  // survivors of the lower bounds (same capacity as cand_*) + their Keogh totals
  // data envelope of the resident shard for one Sakoe-Chiba radius (lower / upper, laid out like series_buf): built at
  // the first DTW call with that radius, dropped when a series is loaded
  // multi-GPU tail (kvm_comm_*): one NCCL communicator per ctx = per rank, packed all-gather buffers
  void* comm = nullptr;      // ncclComm_t
  int comm_rank = 0, comm_world = 1;
  // peer-memory exchange (kvm_comm_ipc_*): every rank's exchange buffer mapped into every other rank
  DevBuf xbuf;                       // [2 parities][world][kXchgSlot doubles] then [2][world] sequence numbers
  void* xpeer[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // peers' xbuf (own: xbuf.p)
  bool xattached = false;
  unsigned long long xseq = 0;       // exchanges done
  DevBuf xerr;
  DevBuf g_send, g_recv;
  PinBuf g_hsend, g_hrecv;
  std::vector<int32_t> g_off;
  std::vector<double> g_dist;
  DevBuf env_lo, env_up;
  int64_t series_len = 0;  // samples in series_buf, pads included (the UCR scans change `count` for the duration of a call)
  int env_rho = -1;
  bool env_failed = false;  // the allocation did not fit: DTW calls use the per-candidate envelope kernel instead
  DevBuf seg_b, seg_first, seg_last, chain_count, chain_prefix, run_key, run_b, run_first, run_last;
  long long cand_cap = 0, ans_cap = 0;
  NormPlanCache norm_cache;
  // streaming cNSM path
  double absmax = 0.0;               // max |sample| of the loaded shard (guard band of the stream, stream_guard())
  StreamPlan splan;
  DevBuf sarena, need_bits, chain_last, flagged, x_off, x_ex, x_ex2, bmax;
  long long n_bmax = 0;
  long long x_cap = 0;
  size_t need_words = 0, chain_last_n = 0;
  bool stream_dirty = true;          // need_bits / chain_last may hold leftovers (first use, or a call that failed)
  int opt_relay = 0, opt_force_all = 0, opt_plan_cache = 1;  // kvm_set_option (defaults from the environment)
  PinBuf sstage;
  // fused window-mean pass: per width device buffers and the host-side result vectors
  struct WmeanSlot {
    DevBuf runs, seg_off, seg_cnt, need_bits, chain_last, flagged, x_off, x_ex, x_ex2;
    long long run_cap = 0, x_cap = 0;
    size_t need_words = 0, chains_n = 0, segs_n = 0;
    std::vector<double> keys;
    std::vector<int32_t> first, last;
  };
  WmeanSlot wm[kvm::kMaxWidths];
  DevBuf wm_counters;
  PinBuf wm_host;
  std::vector<BatchSlot> slots;      // query sets
  cudaStream_t aux[3] = {nullptr, nullptr, nullptr};  // query sets: per-query evaluator / exact stages run concurrently
  cudaEvent_t ev_set = nullptr;
  DevBuf batch_gate;
  Plan plan_scratch;                 // RSM engines: host planning buffers kept across calls (no fresh pages per call)
  std::vector<int32_t> tp_scratch;
  Arena arena_scratch;
  long long h2d_bytes = 0;  // host->device bytes of the current call
  bool eager_valid = false;  // h_off/h_dist hold the first kEagerAnswers answers of the last read_counters
  PinBuf stage, stage2, h_counters, h_off, h_dist, h_key, h_first, h_last, h_b;
  std::vector<int32_t> res_off, run_first_v, run_last_v;
  std::vector<double> res_dist, run_key_v;
};

extern "C" {
static void kvm_comm_release(kvm_ctx* ctx);
}

namespace {

int fail(kvm_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf; else g_create_error = buf;
  return code;
}

#define KVM_CUDA(ctx, expr)                                                                         \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      cudaGetLastError();                                                                           \
      return fail(ctx, _e == cudaErrorMemoryAllocation ? KVM_E_OOM : KVM_E_CUDA, "%s: %s", #expr,   \
                  cudaGetErrorString(_e));                                                          \
    }                                                                                               \
  } while (0)

// ---- Java-compatible helpers used for the O(m) query preparation --------------------------------
inline uint64_t java_bits(double v) {
  if (v != v) return 0x7ff8000000000000ULL;
  uint64_t b;
  std::memcpy(&b, &v, 8);
  return b;
}
// developer tuning knobs (tools/quick_bench.py); unset = built-in defaults
inline int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return (v && *v) ? std::atoi(v) : dflt;
}
// host twin of kvm::hi_key
inline int host_hi_key(double x) {
  uint64_t b;
  std::memcpy(&b, &x, 8);
  const int h = (int)(uint32_t)(b >> 32);
  return h ^ ((h >> 31) & 0x7fffffff);
}
inline int java_double_compare(double a, double b) {
  if (a < b) return -1;
  if (a > b) return 1;
  const int64_t x = (int64_t)java_bits(a), y = (int64_t)java_bits(b);
  return x == y ? 0 : (x < y ? -1 : 1);
}

// K/NormQueryEngine.java:192-198
void query_stats(const double* q, int m, double* meanQ, double* stdQ) {
  double ex = 0, ex2 = 0;
  for (int i = 0; i < m; i++) {
    ex += q[i];
    ex2 += q[i] * q[i];
  }
  *meanQ = ex / m;
  *stdQ = std::sqrt(ex2 / m - *meanQ * *meanQ);
}

// Clamped sliding min/max of radius r — what DtwUtils.lowerUpperLemire computes (K/utils/DtwUtils.java:50-91).
void envelope(const std::vector<double>& t, int r, std::vector<double>& lo, std::vector<double>& up) {
  const int n = (int)t.size();
  lo.resize(n);
  up.resize(n);
  std::vector<int> dq_max(n), dq_min(n);
  int hmax = 0, tmax = 0, hmin = 0, tmin = 0;
  int next = 0;
  for (int i = 0; i < n; i++) {
    const int hi = std::min(n - 1, i + r);
    for (; next <= hi; next++) {
      while (tmax > hmax && t[dq_max[tmax - 1]] <= t[next]) tmax--;
      dq_max[tmax++] = next;
      while (tmin > hmin && t[dq_min[tmin - 1]] >= t[next]) tmin--;
      dq_min[tmin++] = next;
    }
    const int lo_i = std::max(0, i - r);
    while (dq_max[hmax] < lo_i) hmax++;
    while (dq_min[hmin] < lo_i) hmin++;
    up[i] = t[dq_max[hmax]];
    lo[i] = t[dq_min[hmin]];
  }
}

// One interval after the reference's shift / clamp (K/QueryEngine.java:345-349).
// The O(K) pass of make_plan on raw arrays; branch-free so that it vectorises (AVX2 clone picked at load time).
__attribute__((target_clones("avx2", "default"))) int plan_pass(const int32_t* __restrict__ lr, int K, int64_t shift, int64_t m,
                                                                   int64_t n, int64_t lo, int64_t hi, int32_t* __restrict__ cb,
                                                                   int32_t* __restrict__ nsv, int32_t* __restrict__ ncv,
                                                                   int64_t* totals) {
  int64_t cnt = 0, S = 0, V = 0;
  int bad = 0;
  for (int p = 0; p < K; p++) {
    const int64_t left = lr[2 * p], right = lr[2 * p + 1];
    cnt += right - left + 1;
    int64_t begin = left - shift, end = right - shift + m - 1;
    begin = begin < 1 ? 1 : begin;
    end = end > n ? n : end;
    bad |= (end < begin) | (begin < lo) | (end > hi);
    const int64_t ns = end - begin + 1;
    const int64_t nc = ns >= m ? ns - m + 1 : 0;
    cb[p] = (int32_t)(begin - lo);
    nsv[p] = (int32_t)ns;
    ncv[p] = (int32_t)nc;
    S += ns;
    V += nc;
  }
  totals[0] = cnt;
  totals[1] = S;
  totals[2] = V;
  return bad;
}

int make_plan(kvm_ctx* ctx, const int32_t* lr, int K, int shift, int m, Plan* P) {
  P->cbegin.resize(K);
  P->nsamp.resize(K);
  P->ncand.resize(K);
  const int64_t lo = ctx->first, hi = ctx->first + ctx->count - 1, n = ctx->n;
  int64_t totals[3];
  const int bad = plan_pass(lr, K, shift, m, n, lo, hi, P->cbegin.data(), P->nsamp.data(), P->ncand.data(), totals);
  const int64_t cnt = totals[0], S = totals[1], V = totals[2];
  if (bad) {
    for (int p = 0; p < K; p++) {
      const int64_t left = lr[2 * p], right = lr[2 * p + 1];
      int64_t begin = left - shift, end = right - shift + (int64_t)m - 1;
      if (begin < 1) begin = 1;
      if (end > n) end = n;
      if (end < begin)
        return fail(ctx, KVM_E_RANGE, "interval %d [%lld,%lld] shift %d lies outside [1,%lld] (the reference throws)", p,
                    (long long)left, (long long)right, shift, (long long)n);
      if (begin < lo || end > hi)
        return fail(ctx, KVM_E_RANGE, "interval %d needs samples [%lld,%lld]; this ctx holds [%lld,%lld]", p,
                    (long long)begin, (long long)end, (long long)lo, (long long)hi);
    }
  }
  P->cnt_candidate = cnt;
  P->S = S;
  P->V = V;
  return KVM_OK;
}

// The raw-series engines keep no state across window starts, so intervals whose window starts are adjacent are one
// run of candidates for them: merge such neighbours (and drop empty intervals) in place.  An index-free scan handed over
// as 488 k chain-sized intervals becomes ONE run: no per-CTA search through the tile table (19 dependent loads in front
// of 2048 windows of work), 24 bytes of plan instead of 5.9 MB on the wire.  Returns the new interval count.
int coalesce_runs(int32_t* cb, int32_t* nc, int K) {
  int k2 = 0;
  for (int p = 0; p < K; p++) {
    if (nc[p] <= 0) continue;
    if (k2 > 0 && (int64_t)cb[k2 - 1] + nc[k2 - 1] == cb[p] && (int64_t)nc[k2 - 1] + nc[p] <= 0x7fff0000LL) {
      nc[k2 - 1] += nc[p];
    } else {
      cb[k2] = cb[p];
      nc[k2] = nc[p];
      k2++;
    }
  }
  return k2;
}

int check_common(kvm_ctx* ctx, const double* q, int m, double epsilon, const int32_t* lr, int K, kvm_result* out) {
  if (!ctx) return KVM_E_ARG;
  if (!out || !q || m < 1 || K < 0 || (K > 0 && !lr)) return fail(ctx, KVM_E_ARG, "null/invalid argument");
  if (!(epsilon == epsilon)) return fail(ctx, KVM_E_ARG, "epsilon is NaN");
  if (!ctx->series) return fail(ctx, KVM_E_STATE, "no series loaded");
  std::memset(out, 0, sizeof(*out));
  return KVM_OK;
}

int ensure_answers(kvm_ctx* ctx, long long cap) {
  if (cap <= ctx->ans_cap) return KVM_OK;
  ctx->ans_cap = 0;  // (a failure below must not leave a stale capacity over a freed buffer)
  KVM_CUDA(ctx, ctx->ans_off.ensure(sizeof(int32_t) * cap));
  KVM_CUDA(ctx, ctx->ans_dist.ensure(sizeof(double) * cap));
  ctx->ans_cap = cap;
  return KVM_OK;
}

int ensure_cands(kvm_ctx* ctx, long long cap) {
  if (cap <= ctx->cand_cap) return KVM_OK;
  ctx->cand_cap = 0;
  KVM_CUDA(ctx, ctx->cand_off.ensure(sizeof(int32_t) * cap));
  KVM_CUDA(ctx, ctx->cand_mean.ensure(sizeof(double) * cap));
  KVM_CUDA(ctx, ctx->cand_std.ensure(sizeof(double) * cap));
  KVM_CUDA(ctx, ctx->cand2_off.ensure(sizeof(int32_t) * cap));
  KVM_CUDA(ctx, ctx->cand2_mean.ensure(sizeof(double) * cap));
  KVM_CUDA(ctx, ctx->cand2_std.ensure(sizeof(double) * cap));
  KVM_CUDA(ctx, ctx->cand2_lb.ensure(sizeof(double) * cap));
  KVM_CUDA(ctx, ctx->cand_lb.ensure(sizeof(double) * cap));
  ctx->cand_cap = cap;
  return KVM_OK;
}

CandList cands2_of(kvm_ctx* ctx);
int launch_lb_data(kvm_ctx* ctx, const double* q, const double* uq, const double* lq, int m, int rho, double eps2_hi);
bool ensure_envelope(kvm_ctx* ctx, int rho);
int lb_warp_path_max();

AnswerSink sink_of(kvm_ctx* ctx) {
  return AnswerSink{ctx->ans_off.as<int32_t>(), ctx->ans_dist.as<double>(),
                    ctx->counters.as<unsigned long long>() + kCntAnswers, ctx->ans_cap};
}
CandList cands2_of(kvm_ctx* ctx) {
  return CandList{ctx->cand2_off.as<int32_t>(), ctx->cand2_mean.as<double>(), ctx->cand2_std.as<double>(),
                  ctx->counters.as<unsigned long long>() + kCntCand2, ctx->cand_cap, ctx->cand2_lb.as<double>()};
}
CandList cands_of(kvm_ctx* ctx) {
  return CandList{ctx->cand_off.as<int32_t>(), ctx->cand_mean.as<double>(), ctx->cand_std.as<double>(),
                  ctx->counters.as<unsigned long long>() + kCntCand, ctx->cand_cap};
}

int upload_arena(kvm_ctx* ctx, const Arena& A) {
  ctx->norm_cache.valid = false;
  ctx->h2d_bytes += (long long)A.host.size();
  KVM_CUDA(ctx, ctx->stage.ensure(A.host.size() + 256));
  KVM_CUDA(ctx, ctx->arena.ensure(A.host.size() + 256));
  std::memcpy(ctx->stage.p, A.host.data(), A.host.size());
  KVM_CUDA(ctx, cudaMemcpyAsync(ctx->arena.p, ctx->stage.p, A.host.size(), cudaMemcpyHostToDevice, ctx->stream));
  return KVM_OK;
}

constexpr long long kEagerAnswers = 1024;  // answers fetched together with the counters (one sync for typical queries)

// Enqueue the device->host copies of the counters and of the first kEagerAnswers answers (no synchronisation).
int enqueue_counter_read(kvm_ctx* ctx) {
  KVM_CUDA(ctx, ctx->h_counters.ensure(sizeof(unsigned long long) * kNumCounters));
  KVM_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters.p, ctx->counters.p, sizeof(unsigned long long) * kNumCounters,
                                cudaMemcpyDeviceToHost, ctx->stream));
  ctx->eager_valid = false;
  if (ctx->ans_cap >= kEagerAnswers) {
    KVM_CUDA(ctx, ctx->h_off.ensure(sizeof(int32_t) * kEagerAnswers));
    KVM_CUDA(ctx, ctx->h_dist.ensure(sizeof(double) * kEagerAnswers));
    KVM_CUDA(ctx, cudaMemcpyAsync(ctx->h_off.p, ctx->ans_off.p, sizeof(int32_t) * kEagerAnswers, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    KVM_CUDA(ctx, cudaMemcpyAsync(ctx->h_dist.p, ctx->ans_dist.p, sizeof(double) * kEagerAnswers, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    ctx->eager_valid = true;
  }
  return KVM_OK;
}

int read_counters(kvm_ctx* ctx, unsigned long long* out) {
  KVM_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters.p, ctx->counters.p, sizeof(unsigned long long) * kNumCounters,
                                cudaMemcpyDeviceToHost, ctx->stream));
  ctx->eager_valid = false;
  if (ctx->ans_cap >= kEagerAnswers) {
    KVM_CUDA(ctx, ctx->h_off.ensure(sizeof(int32_t) * kEagerAnswers));
    KVM_CUDA(ctx, ctx->h_dist.ensure(sizeof(double) * kEagerAnswers));
    KVM_CUDA(ctx, cudaMemcpyAsync(ctx->h_off.p, ctx->ans_off.p, sizeof(int32_t) * kEagerAnswers, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    KVM_CUDA(ctx, cudaMemcpyAsync(ctx->h_dist.p, ctx->ans_dist.p, sizeof(double) * kEagerAnswers, cudaMemcpyDeviceToHost,
                                  ctx->stream));
    ctx->eager_valid = true;
  }
  KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::memcpy(out, ctx->h_counters.p, sizeof(unsigned long long) * kNumCounters);
  return KVM_OK;
}

// Copy the sparse answers back and put them in ascending offset order (the reference's scan order).
int fetch_answers(kvm_ctx* ctx, long long count, kvm_result* out) {
  out->h2d_bytes = (int32_t)std::min<long long>(ctx->h2d_bytes, INT32_MAX);
  ctx->res_off.resize((size_t)count);
  ctx->res_dist.resize((size_t)count);
  if (count > 0) {
    if (!(ctx->eager_valid && count <= kEagerAnswers)) {  // otherwise read_counters already brought them
      KVM_CUDA(ctx, ctx->h_off.ensure(sizeof(int32_t) * count));
      KVM_CUDA(ctx, ctx->h_dist.ensure(sizeof(double) * count));
      KVM_CUDA(ctx, cudaMemcpyAsync(ctx->h_off.p, ctx->ans_off.p, sizeof(int32_t) * count, cudaMemcpyDeviceToHost,
                                    ctx->stream));
      KVM_CUDA(ctx, cudaMemcpyAsync(ctx->h_dist.p, ctx->ans_dist.p, sizeof(double) * count, cudaMemcpyDeviceToHost,
                                    ctx->stream));
      KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->eager_valid = false;
    const int32_t* ho = ctx->h_off.as<int32_t>();
    const double* hd = ctx->h_dist.as<double>();
    std::vector<int64_t> idx((size_t)count);
    for (int64_t i = 0; i < count; i++) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](int64_t a, int64_t b) { return ho[a] < ho[b]; });
    for (int64_t i = 0; i < count; i++) {
      ctx->res_off[i] = ho[idx[i]];
      ctx->res_dist[i] = hd[idx[i]];
    }
  }
  out->count = count;
  out->offsets = ctx->res_off.data();
  out->distances = ctx->res_dist.data();
  return KVM_OK;
}

int begin_call(kvm_ctx* ctx) {
  ctx->h2d_bytes = 0;
  KVM_CUDA(ctx, cudaSetDevice(ctx->device));
  KVM_CUDA(ctx, ctx->counters.ensure(sizeof(unsigned long long) * kNumCounters));
  KVM_CUDA(ctx, ctx->h_counters.ensure(sizeof(unsigned long long) * kNumCounters));
  return KVM_OK;
}

int zero_counters(kvm_ctx* ctx) {
  KVM_CUDA(ctx, cudaMemsetAsync(ctx->counters.p, 0, sizeof(unsigned long long) * kNumCounters, ctx->stream));
  return KVM_OK;
}

// Tile prefix for kernels that enumerate candidates straight from the intervals.
void tile_prefix_of(const Plan& P, int tile, int64_t* n_tiles, std::vector<int32_t>& tp) {
  tp.resize(P.ncand.size() + 1);
  int64_t acc = 0;
  for (size_t p = 0; p < P.ncand.size(); p++) {
    tp[p] = (int32_t)acc;
    acc += (P.ncand[p] + tile - 1) / tile;
  }
  tp[P.ncand.size()] = (int32_t)acc;
  *n_tiles = acc;
}

// Everything the cNSM engines share: query statistics, pre-gate constants, walker launch.
struct NormSetup {
  double meanQ = 0, stdQ = 0, inv_alpha = 0;
  int n_regions = 0;
  bool degenerate = false;  // stdQ is 0/NaN: no window can pass the gate
};

// Conservative pre-gate key ranges of one query (see launch_walker).
void gate_keys(int m, const NormSetup& S, double alpha, double beta, int* mean_klo, unsigned* mean_kspan, int* var_klo,
               unsigned* var_kspan) {
  const double dm = (double)m;
  const double amq = std::fabs(S.meanQ) + std::fabs(beta);
  const double beta_hi = beta + 1e-14 * amq + 1e-290;
  const double hi2 = (alpha * S.stdQ) * (alpha * S.stdQ), lo2 = (S.stdQ * S.inv_alpha) * (S.stdQ * S.inv_alpha);
  const double d2 = 1e-13 * (hi2 + 2.0 * amq * amq) + 1e-290;
  const double var_hi = hi2 * (1.0 + 1e-12) + d2, var_lo = lo2 * (1.0 - 1e-12) - d2;
  double e_lo = dm * (S.meanQ - beta_hi), e_hi = dm * (S.meanQ + beta_hi);
  e_lo -= std::fabs(e_lo) * 1e-12;
  e_hi += std::fabs(e_hi) * 1e-12;
  double v_lo = dm * dm * var_lo, v_hi = dm * dm * var_hi;
  v_lo -= std::fabs(v_lo) * 1e-12;
  v_hi += std::fabs(v_hi) * 1e-12;
  const long long mk_lo = (long long)host_hi_key(e_lo) - 1, mk_hi = (long long)host_hi_key(e_hi) + 1;
  const long long vk_lo = (long long)host_hi_key(v_lo) - 1, vk_hi = (long long)host_hi_key(v_hi) + 1;
  *mean_klo = (int)std::max<long long>(mk_lo, INT32_MIN);
  *var_klo = (int)std::max<long long>(vk_lo, INT32_MIN);
  *mean_kspan = (unsigned)(std::min<long long>(mk_hi, INT32_MAX) - *mean_klo);
  *var_kspan = (unsigned)(std::min<long long>(vk_hi, INT32_MAX) - *var_klo);
}

#ifndef KVM_RELAY_STAGES
#define KVM_RELAY_STAGES 3
#endif
constexpr int kRelayStages = KVM_RELAY_STAGES;

int launch_walker(kvm_ctx* ctx, const Plan& P, int K, int m, double alpha, double beta, const NormSetup& S,
                  size_t off_cbegin, size_t off_nsamp, size_t off_region_base, int* launches) {
  const unsigned char* base = ctx->arena.as<unsigned char>();
  WalkParams W;
  W.T = ctx->series;
  W.cbegin = reinterpret_cast<const int32_t*>(base + off_cbegin);
  W.cnsamp = reinterpret_cast<const int32_t*>(base + off_nsamp);
  W.region_base = reinterpret_cast<const long long*>(base + off_region_base);
  W.K = K;
  W.m = m;
  W.first_global = (int32_t)ctx->first;
  W.dm = (double)m;
  W.idx_hi = (int)((ctx->count + kTailPad - 2) & ~int64_t(1));
  // Conservative pre-gate: a superset of the exact gate (rounding of the chain sums' products is far below
  // these slacks, and the integer keys widen each bound by one high-word unit); the exact gate is
  // re-evaluated with the reference's arithmetic by the evaluators.
  gate_keys(m, S, alpha, beta, &W.mean_klo, &W.mean_kspan, &W.var_klo, &W.var_kspan);
  W.e_off = ctx->wl_off.as<int32_t>();
  W.e_ex = ctx->wl_ex.as<double>();
  W.e_ex2 = ctx->wl_ex2.as<double>();
  W.region_count = ctx->region_count.as<int32_t>();
  W.tile_prefix = ctx->tile_prefix.as<int32_t>();
  W.totals = ctx->counters.as<unsigned long long>() + kCntTiles;
  W.done = reinterpret_cast<unsigned int*>(ctx->counters.as<unsigned long long>() + kCntDone);
  W.batch = nullptr;
  // Relay walker, 5 relay warps + loader, 3-stage tile ring: ~100 KB of shared memory per CTA, 2 CTAs (= 64 chains)
  // per SM.  More resident chains would not help: ~8k chains x (m-1) x 8 B of lag windows is what stays
  // L2-resident.  (A 2-stage instantiation is deliberately not built: ptxas 12.9 emits its hinted
  // LDGSTS with an uninitialised uniform descriptor register -> "illegal instruction"; the Makefile greps for it.)
#ifdef KVM_RELAY_PROF
  unsigned long long z8[64] = {0};
  cudaMemcpyToSymbolAsync(g_relay_prof, z8, sizeof(z8), 0, cudaMemcpyHostToDevice, ctx->stream);
#endif
  if ((m % 2) == 0) cnsm_relay_kernel<kRelayStages, 1><<<S.n_regions, kRelayThreads, relay_smem_bytes(kRelayStages), ctx->stream>>>(W);
  else cnsm_relay_kernel<kRelayStages, 0><<<S.n_regions, kRelayThreads, relay_smem_bytes(kRelayStages), ctx->stream>>>(W);
#ifdef KVM_RELAY_PROF
  cudaStreamSynchronize(ctx->stream);
  cudaMemcpyFromSymbol(z8, g_relay_prof, sizeof(z8));
  for (int w = 0; w < kRelayWarps; w++) {
    const unsigned long long* z = z8 + 8 * w;
    const double t = (double)std::max<unsigned long long>(z[5], 1);
    std::fprintf(stderr, "[relay prof] warp %d cycles per turn: tile-wait %.0f prepare %.0f state-wait %.0f walk %.0f gate %.0f (turns %llu)\n",
                 w, z[0] / t, z[1] / t, z[2] / t, z[3] / t, z[4] / t, z[5]);
  }
#endif
  KVM_CUDA(ctx, cudaEventRecord(ctx->evs[0], ctx->stream));
  *launches += 1;
  KVM_CUDA(ctx, cudaGetLastError());
  return KVM_OK;
}

// ev0 -> evs[0] -> evs[1] -> ev1
void add_stage_ms(kvm_ctx* ctx, kvm_result* out) {
  float a = 0.f, b = 0.f, c = 0.f;
  cudaEventElapsedTime(&a, ctx->ev0, ctx->evs[0]);
  cudaEventElapsedTime(&b, ctx->evs[0], ctx->evs[1]);
  cudaEventElapsedTime(&c, ctx->evs[1], ctx->ev1);
  out->stage_ms[0] += a;
  out->stage_ms[1] += b;
  out->stage_ms[2] += c;
}

double elapsed_ms(kvm_ctx* ctx) {
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
  return (double)ms;
}

enum class Mode { kEd, kDtw };
int verify_norm_stream(kvm_ctx* ctx, Mode mode, const double* q, int m, double epsilon, int rho, double alpha, double beta,
                       const int32_t* lr, int K, int shift, int nt, kvm_result* out, bool may_defer = true);
int stream_nt_for(int m);

// cNSM-ED and cNSM-DTW share everything up to the evaluator.
int verify_norm(kvm_ctx* ctx, Mode mode, const double* q, int m, double epsilon, int rho, double alpha, double beta,
                const int32_t* lr, int K, int shift, kvm_result* out) {
  const auto t_begin = std::chrono::steady_clock::now();
  auto since = [&](std::chrono::steady_clock::time_point t0) {
    return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
  };
  int rc = check_common(ctx, q, m, epsilon, lr, K, out);
  if (rc) return rc;
  if (mode == Mode::kDtw && (rho < 0 || m < 3)) return fail(ctx, KVM_E_ARG, "DTW needs rho >= 0 and m >= 3");
  if ((rc = begin_call(ctx))) return rc;
  {
    // default: the streaming statistics pass; KVM_CNSM_PATH=relay (or a query too long for its shared-memory tile, or a
    // series with non-finite samples) selects the round-1 relay walker, which walks every chain exactly
    const int nt = stream_nt_for(m);
    if (!ctx->opt_relay && nt > 0 && std::isfinite(ctx->absmax)) return verify_norm_stream(ctx, mode, q, m, epsilon, rho, alpha, beta, lr, K, shift, nt, out);
  }
  const double t_dbg0 = since(t_begin);
  NormPlanCache& C = ctx->norm_cache;
  const bool cache_on = ctx->opt_plan_cache != 0;  // off: plan and upload on every call (bench.py's e2e leg)
  const bool reuse = cache_on && C.valid && C.K == K && C.shift == shift && C.m == m &&
                     std::memcmp(C.lr.data(), lr, sizeof(int32_t) * 2 * (size_t)K) == 0;
  if (!reuse) {
    C.valid = false;  // (the plan's vectors keep their capacity from call to call)
    if ((rc = make_plan(ctx, lr, K, shift, m, &C.plan))) return rc;
  }
  const double t_dbg1 = since(t_begin);
  const Plan& P = C.plan;
  out->cnt_candidate = P.cnt_candidate;
  out->n_verified = P.V;
  out->s_total = P.S;

  NormSetup S;
  query_stats(q, m, &S.meanQ, &S.stdQ);
  const double t_dbg2 = since(t_begin);
  S.inv_alpha = 1.0 / alpha;
  S.degenerate = !(S.stdQ > 0.0) || !(S.stdQ < INFINITY);
  if (P.V == 0 || S.degenerate) return fetch_answers(ctx, 0, out);

  // chains -> walker regions (one region per walker CTA = 32 chains)
  const int n_regions = (K + 31) / 32;
  S.n_regions = n_regions;
  double t_prep = 0, t_upload = 0;
  if (!reuse) {
    // built in place in the pinned staging block: [cbegin i32 x K | walker nsamp i32 x K | region_base i64 x (R+1)]
    auto up256 = [](size_t x) { return (x + 255) & ~size_t(255); };
    C.o_cbegin = 0;
    C.o_nsamp = up256(sizeof(int32_t) * (size_t)K);
    C.o_rbase = up256(C.o_nsamp + sizeof(int32_t) * (size_t)K);
    const size_t bytes = C.o_rbase + sizeof(long long) * (size_t)(n_regions + 1);
    KVM_CUDA(ctx, ctx->stage.ensure(bytes + 256));
    KVM_CUDA(ctx, ctx->arena.ensure(bytes + 256));
    unsigned char* st = static_cast<unsigned char*>(ctx->stage.p);
    std::memcpy(st + C.o_cbegin, P.cbegin.data(), sizeof(int32_t) * (size_t)K);
    int32_t* walk_nsamp = reinterpret_cast<int32_t*>(st + C.o_nsamp);
    long long* region_base = reinterpret_cast<long long*>(st + C.o_rbase);
    const int32_t* __restrict__ ncv = P.ncand.data();
    const int32_t* __restrict__ nsv = P.nsamp.data();
    for (int c = 0; c < K; c++) walk_nsamp[c] = ncv[c] > 0 ? nsv[c] : 0;  // (vectorisable passes)
    long long acc = 0;
    for (int r = 0; r < n_regions; r++) {
      region_base[r] = acc;
      const int c_hi = std::min(K, r * 32 + 32);
      long long sum = 0;
      for (int c = r * 32; c < c_hi; c++) sum += ncv[c];
      acc += sum;
    }
    region_base[n_regions] = acc;
    t_prep = since(t_begin);
    KVM_CUDA(ctx, cudaMemcpyAsync(ctx->arena.p, ctx->stage.p, bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += (long long)bytes;
    if (cache_on) {
      C.lr.assign(lr, lr + 2 * (size_t)K);
      C.K = K;
      C.shift = shift;
      C.m = m;
      C.n_regions = n_regions;
      C.valid = true;  // (a failure below leaves the arena intact)
    }
  } else {
    t_prep = since(t_begin);
  }
  t_upload = since(t_begin);
  const size_t o_cbegin = C.o_cbegin, o_nsamp = C.o_nsamp, o_rbase = C.o_rbase;

  KVM_CUDA(ctx, ctx->wl_off.ensure(sizeof(int32_t) * (size_t)P.V));
  KVM_CUDA(ctx, ctx->wl_ex.ensure(sizeof(double) * (size_t)P.V));
  KVM_CUDA(ctx, ctx->wl_ex2.ensure(sizeof(double) * (size_t)P.V));
  KVM_CUDA(ctx, ctx->region_count.ensure(sizeof(int32_t) * (n_regions + 1)));
  KVM_CUDA(ctx, ctx->tile_prefix.ensure(sizeof(int32_t) * (n_regions + 2)));
  if ((rc = ensure_answers(ctx, std::max<long long>(ctx->ans_cap, 1 << 16)))) return rc;
  if ((rc = ensure_cands(ctx, std::max<long long>(ctx->cand_cap, 1 << 18)))) return rc;

  const unsigned char* base = ctx->arena.as<unsigned char>();
  const double eps2 = epsilon * epsilon;
  unsigned long long cnt[kNumCounters];
  double total_ms = 0;
  size_t o_zq = 0, o_order = 0, o_uq = 0, o_lq = 0;
  for (int attempt = 0; attempt < 8; attempt++) {
    int launches = 0;
    if ((rc = zero_counters(ctx))) return rc;
    KVM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    if ((rc = launch_walker(ctx, P, K, m, alpha, beta, S, o_cbegin, o_nsamp, o_rbase, &launches))) return rc;
    if (attempt == 0) {
      // The walker needs only the intervals and the query's mean/std; the O(m log m) query preparation below
      // (z-normalisation; ED: stable sort by |z| descending; DTW: envelope) runs on the host while it walks.
      std::vector<double> z(m), zq(m), uq, lq;
      std::vector<int32_t> order(m);
      for (int i = 0; i < m; i++) z[i] = (q[i] - S.meanQ) / S.stdQ;  // K/NormQueryEngine.java:438-441
      for (int i = 0; i < m; i++) order[i] = i;
      if (mode == Mode::kEd) {
        std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
          return java_double_compare(std::fabs(z[b]), std::fabs(z[a])) < 0;  // :448
        });
        for (int i = 0; i < m; i++) zq[i] = z[order[i]];
      } else {
        zq = z;
        envelope(z, rho, lq, uq);  // K/NormQueryEngineDtw.java:469
      }
      Arena Q;
      o_zq = Q.add(zq.data(), sizeof(double) * m);
      o_order = Q.add(order.data(), sizeof(int32_t) * m);
      o_uq = Q.add(uq.data(), sizeof(double) * uq.size());
      o_lq = Q.add(lq.data(), sizeof(double) * lq.size());
      KVM_CUDA(ctx, ctx->stage2.ensure(Q.host.size() + 256));
      KVM_CUDA(ctx, ctx->qarena.ensure(Q.host.size() + 256));
      std::memcpy(ctx->stage2.p, Q.host.data(), Q.host.size());
      KVM_CUDA(ctx, cudaMemcpyAsync(ctx->qarena.p, ctx->stage2.p, Q.host.size(), cudaMemcpyHostToDevice, ctx->stream));
      ctx->h2d_bytes += (long long)Q.host.size();
    }
    const unsigned char* qbase = ctx->qarena.as<unsigned char>();

    EvalParams E;
    E.T = ctx->series;
    E.first_global = (int32_t)ctx->first;
    E.m = m;
    E.e_off = ctx->wl_off.as<int32_t>();
    E.e_ex = ctx->wl_ex.as<double>();
    E.e_ex2 = ctx->wl_ex2.as<double>();
    E.region_base = reinterpret_cast<const long long*>(base + o_rbase);
    E.region_count = ctx->region_count.as<int32_t>();
    E.tile_prefix = ctx->tile_prefix.as<int32_t>();
    E.totals = ctx->counters.as<unsigned long long>() + kCntTiles;
    E.n_regions = n_regions;
    E.zq = reinterpret_cast<const double*>(qbase + o_zq);
    E.order = reinterpret_cast<const int32_t*>(qbase + o_order);
    E.meanQ = S.meanQ;
    E.stdQ = S.stdQ;
    E.alpha = alpha;
    E.inv_alpha = S.inv_alpha;
    E.beta = beta;
    E.eps2 = eps2;
    E.eps2_hi = eps2 * (1.0 + 1e-9) + 1e-18;
    E.out = cands_of(ctx);
    E.gate_pass = ctx->counters.as<unsigned long long>() + kCntGate;
    const int eval_grid = ctx->n_sms * 16;
    if (mode == Mode::kEd) {
      cnsm_ed_eval_kernel<<<eval_grid, kEvalTile, 0, ctx->stream>>>(E);
      KVM_CUDA(ctx, cudaEventRecord(ctx->evs[1], ctx->stream));
      ExactEdParams X;
      X.T = E.T;
      X.first_global = E.first_global;
      X.m = m;
      X.zq = E.zq;
      X.order = E.order;
      X.eps2 = eps2;
      X.eps2_hi = E.eps2_hi;
      X.n_exact = ctx->counters.as<unsigned long long>() + kCntFlag;
      X.in = E.out;
      X.sink = sink_of(ctx);
      X.win_cap = 0;
      X.win_cap = 0;
    X.q_cap = 0;
    cnsm_ed_exact_kernel<false><<<ctx->n_sms * 6, 128, sizeof(double) * kExactChunk * 4, ctx->stream>>>(X);
      launches += 2;
    } else {
      LbNormParams L;
      L.E = E;
      L.Q = LbQuery{E.zq, reinterpret_cast<const double*>(qbase + o_uq), reinterpret_cast<const double*>(qbase + o_lq),
                    m, E.eps2_hi};
      cnsm_dtw_lb_kernel<<<eval_grid, kEvalTile, 0, ctx->stream>>>(L);
      KVM_CUDA(ctx, cudaEventRecord(ctx->evs[1], ctx->stream));
      launches += 1;
      DtwParams D;
      D.T = E.T;
      D.first_global = E.first_global;
      D.m = m;
      D.rho = rho;
      D.q = E.zq;
      D.uq = L.Q.uq;
      D.lq = L.Q.lq;
      D.eps2 = eps2;
      D.eps2_hi = E.eps2_hi;
      D.n_abandoned = ctx->counters.as<unsigned long long>() + kCntFlag;
      D.n_cells = ctx->counters.as<unsigned long long>() + kCntCells;
      D.in = E.out;
      D.sink = sink_of(ctx);
      if ((rc = launch_dtw(ctx, D))) return rc;
      launches += 1;
    }
    KVM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    KVM_CUDA(ctx, cudaGetLastError());
    const double t_launch = since(t_begin);
    if ((rc = read_counters(ctx, cnt))) return rc;
    if (env_int("KVM_TIMING", 0))
      std::fprintf(stderr, "[kvm] begin %.0f plan %.0f qstats %.0f prep %.0f us, upload %.0f, launched %.0f, synced %.0f\n", t_dbg0, t_dbg1, t_dbg2, t_prep, t_upload, t_launch, since(t_begin));
    total_ms += elapsed_ms(ctx);
    add_stage_ms(ctx, out);
    out->n_launches += launches;
    const bool cand_over = (long long)cnt[kCntCand] > ctx->cand_cap, ans_over = (long long)cnt[kCntAnswers] > ctx->ans_cap;
    if (!cand_over && !ans_over) break;
    if (attempt == 7) return fail(ctx, KVM_E_OOM, "result buffers kept overflowing");
    if (cand_over && (rc = ensure_cands(ctx, (long long)cnt[kCntCand] + 1024))) return rc;
    if (ans_over && (rc = ensure_answers(ctx, (long long)cnt[kCntAnswers] + 1024))) return rc;
  }
  out->kernel_ms = total_ms;
  out->n_gate_pass = (int64_t)cnt[kCntGate];
  if (mode == Mode::kEd) out->n_exact = (int64_t)cnt[kCntFlag]; else out->n_lb_pass = (int64_t)cnt[kCntCand];
  return fetch_answers(ctx, (long long)cnt[kCntAnswers], out);
}


// ---------------------------------------------------------------------------------------------------------------
// Streaming cNSM path (stream_kernels.cuh): stream + guard band, exact re-walk of the flagged chains, exact stages.

constexpr double kUlpHalf = 1.1102230246251565e-16;  // 2^-53

// Guard bands of the stream's gate and of its in-stream lower bound.  Every term is a worst-case bound, in units of
// A = max |sample| over everything a chain has summed when it reaches a window (evaluated per tile on the device from
// the block-maximum table, kvm::stream_guard_eval):
//  * the reference chain (K/NormQueryEngine.java:498-499,523-524) has performed at most 2*l_max roundings when it
//    reaches a window, each at most u*|partial sum| <= u*m*A (ex) or u*m*A^2 (ex2); fl(d*d) itself is off by at most
//    u*d^2 and a window holds m of them;
//  * the stream's own sums go through fewer than 320 roundings (11-term group sums, block scan, remainder, 11 slides)
//    on values bounded by (samples of a tile)*A resp. *A^2;
//  * the reference's divide / multiply / subtract / sqrt at :508-511 and our threshold arithmetic add a few u
//    relative to the quantities near the thresholds (inside stream_guard_eval).
// The sum is doubled.  A window whose stream sums lie outside the *_out band certainly fails the reference's gate,
// inside the *_in band it certainly passes; in between it is re-walked exactly.
// In-stream lower bound: an answer has passed the exact gate (std >= stdQ/alpha) and |x_k| <= max|zQ| + eps in every
// term, so |x_stream - x_ref| <= dx and sqrt(partial sum) moves by at most sqrt(n_terms)*dx.
kvm::GuardCoef stream_guard(int m, int nt, int64_t l_max, const NormSetup& S, double alpha, double beta, double eps,
                            double zq_absmax, int n_terms) {
  const double u = kUlpHalf;
  const double dm = (double)m;
  const double nt_s = (double)kvm::kGroup * nt + dm;
  const double E1 = 2.0 * (double)l_max * u * dm * 1.01;
  const double E2 = (2.0 * (double)l_max * 1.01 + dm) * u * dm;
  const double e1 = 320.0 * u * nt_s;
  const double e2 = (320.0 * nt_s + 33.0 * 8.0) * u;
  kvm::GuardCoef C;
  C.cd1 = 2.0 * (E1 + e1) * 1.000001;
  C.cd2 = 2.0 * (E2 + e2) * 1.000002;
  C.dm = dm;
  C.inv_dm = 1.0 / dm;
  C.abs_mean_beta = std::fabs(S.meanQ) + std::fabs(beta);
  C.c_lo = dm * (S.meanQ - beta);
  C.c_hi = dm * (S.meanQ + beta);
  C.hi = (alpha * S.stdQ) * (alpha * S.stdQ);
  C.lo = (S.stdQ * S.inv_alpha) * (S.stdQ * S.inv_alpha);
  C.std_lo = S.stdQ * S.inv_alpha * (1.0 - 1e-9);
  C.inv_std_lo = C.std_lo > 0.0 ? 1.02 / C.std_lo : 0.0;
  C.xm = zq_absmax + std::fabs(eps) + 1.0;
  C.sqrt_terms = std::sqrt((double)std::max(n_terms, 8));
  C.eps_abs = std::fabs(eps);
  return C;
}

// threads per CTA of the stream kernel for this query length: the variant that keeps the most threads resident per
// SM (shared memory is the limit; ties go to the larger tile = fewer halo re-reads).  0 = the tile does not fit: the
// relay walker takes the call.
int stream_nt_for(int m) {
  static const int forced = env_int("KVM_STREAM_NT", 0);  // developer knob
  constexpr size_t kSmemSm = 227 * 1024, kSmemCta = 227 * 1024 - 512;
  const int cand[] = {256, 224, 192, 160, 128};
  int best = 0, best_threads = 0;
  for (int nt : cand) {
    const size_t sm = kvm::stream_smem_bytes(nt, m);
    if (sm > kSmemCta) continue;
    if (forced == nt) return nt;
    int ctas = (int)(kSmemSm / (sm + 1024));
    ctas = std::min(ctas, nt <= 192 ? 3 : (nt <= 256 ? 2 : 1));  // the kernels' __launch_bounds__
    if (nt * ctas > best_threads) {
      best_threads = nt * ctas;
      best = nt;
    }
  }
  return best;
}

template <int kMode>
cudaError_t launch_stream(int nt, int grid, size_t smem, cudaStream_t st, const kvm::StreamParams& P) {
  switch (nt) {
    case 256: kvm::cnsm_stream_kernel<256, kMode><<<grid, 256, smem, st>>>(P); break;
    case 224: kvm::cnsm_stream_kernel<224, kMode><<<grid, 224, smem, st>>>(P); break;
    case 192: kvm::cnsm_stream_kernel<192, kMode><<<grid, 192, smem, st>>>(P); break;
    case 160: kvm::cnsm_stream_kernel<160, kMode><<<grid, 160, smem, st>>>(P); break;
    default: kvm::cnsm_stream_kernel<128, kMode><<<grid, 128, smem, st>>>(P); break;
  }
  return cudaGetLastError();
}

int set_stream_attrs() {
  const int big = 227 * 1024 - 512;  // the opt-in maximum minus the kernel's static shared memory
  cudaError_t e = cudaSuccess;
#define KVM_ATTR(NT, MODE) \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(kvm::cnsm_stream_kernel<NT, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
  KVM_ATTR(256, 0) KVM_ATTR(224, 0) KVM_ATTR(192, 0) KVM_ATTR(160, 0) KVM_ATTR(128, 0)
  KVM_ATTR(256, 1) KVM_ATTR(224, 1) KVM_ATTR(192, 1) KVM_ATTR(160, 1) KVM_ATTR(128, 1)
#undef KVM_ATTR
  return e == cudaSuccess ? 0 : 1;
}

// 0 when intervals 1..K-1 each start right after their predecessor and hold c0 window starts (the last one: 1..c0).
// Written on the (left, right) pairs as 64-bit words so that the loop vectorises.
__attribute__((target_clones("avx2", "default"))) int regular_grid_pass(const int32_t* __restrict__ lr, int K, int32_t c0) {
  unsigned bad = 0;
  for (int p = 1; p < K - 1; p++) {
    const uint32_t left = (uint32_t)lr[2 * p], right = (uint32_t)lr[2 * p + 1], prev = (uint32_t)lr[2 * p - 1];
    bad |= (left - prev - 1u) | (right - left - (uint32_t)(c0 - 1));
  }
  if (K > 1) {
    const int64_t left = lr[2 * K - 2], right = lr[2 * K - 1], prev = lr[2 * K - 3];
    const int64_t c = right - left + 1;
    bad |= (unsigned)((left != prev + 1) | (c < 1) | (c > c0));
  }
  return bad != 0;
}

// Build (or reuse) the stream plan for this interval list: [tiles | cbegin | ncand | vbase] in ctx->sarena.
int stream_plan(kvm_ctx* ctx, const int32_t* lr, int K, int shift, int m, int nt, bool cache_on, bool may_defer) {
  StreamPlan& SP = ctx->splan;
  if (cache_on && SP.valid && SP.K == K && SP.shift == shift && SP.m == m && SP.nt == nt &&
      std::memcmp(SP.lr.data(), lr, sizeof(int32_t) * 2 * (size_t)K) == 0)
    return KVM_OK;
  SP.valid = false;
  const int W = kvm::kGroup * nt;
  // Regular grid (an index-free scan: adjacent intervals of one length, the last one possibly shorter, nothing clamped):
  // recognised in one vectorisable pass over the caller's list; needs no per-chain tables on either side.
  if (K >= 1) {
    const int64_t c0 = (int64_t)lr[1] - lr[0] + 1;
    int bad = (c0 < 1) | (c0 > INT32_MAX / 2);
    // A long list that looks regular from its two ends (K chains of c0 starts cover exactly [first, last]) is taken as
    // regular NOW and checked in full while the kernels run (verify_norm_stream): the pass over 488 k intervals costs
    // 0.13 ms, 6 % of an n = 1e9 query, and the kernels do not need it.  A list that fails the check is re-planned.
    bool defer = false;
    if (!bad && may_defer && !cache_on && K >= 4096) {
      const int64_t V0 = (int64_t)lr[2 * K - 1] - lr[0] + 1;
      defer = V0 > (int64_t)(K - 1) * c0 && V0 <= (int64_t)K * c0 && (int64_t)lr[2 * (K - 1)] == (int64_t)lr[0] + (int64_t)(K - 1) * c0;
    }
    if (!bad && !defer) bad = regular_grid_pass(lr, K, (int32_t)c0);
    const int64_t first_begin = (int64_t)lr[0] - shift, last_end = (int64_t)lr[2 * K - 1] - shift + m - 1;
    const int64_t lo = ctx->first, hi = ctx->first + ctx->count - 1;
    if (!bad && first_begin >= 1 && first_begin >= lo && last_end <= ctx->n && last_end <= hi) {
      const int64_t V = (int64_t)lr[2 * K - 1] - lr[0] + 1;
      SP.cnt_candidate = V;
      SP.V = V;
      SP.S = V + (int64_t)K * (m - 1);
      SP.regular = 1;
      SP.unchecked = defer;
      SP.chunk = (int)c0;
      SP.s_base = (int)(first_begin - lo);
      SP.n_chains = K;
      SP.n_segments = 1;
      SP.n_tiles = (int)((V + W - 1) / W);
      SP.l_max = c0 + m - 1;
      SP.bytes = 0;
      SP.K = K;
      SP.shift = shift;
      SP.m = m;
      SP.nt = nt;
      if (cache_on) {
        SP.lr.assign(lr, lr + 2 * (size_t)K);
        SP.valid = true;
      }
      return KVM_OK;
    }
  }
  Plan& P = ctx->plan_scratch;
  int rc = make_plan(ctx, lr, K, shift, m, &P);
  if (rc) return rc;
  SP.cnt_candidate = P.cnt_candidate;
  SP.V = P.V;
  SP.S = P.S;
  SP.regular = 0;
  SP.unchecked = false;
  SP.chunk = 0;
  // upper bounds for the staging layout: every live chain opens at most one extra tile
  const size_t max_tiles = (size_t)(P.V / W) + (size_t)K + 2;
  auto up256 = [](size_t x) { return (x + 255) & ~size_t(255); };
  SP.o_tiles = 0;
  SP.o_cb = up256(sizeof(kvm::StreamTile) * max_tiles);
  SP.o_nc = up256(SP.o_cb + sizeof(int32_t) * ((size_t)K + 1));
  SP.o_vb = up256(SP.o_nc + sizeof(int32_t) * ((size_t)K + 1));
  const size_t cap_bytes = up256(SP.o_vb + sizeof(int32_t) * ((size_t)K + 1));
  KVM_CUDA(ctx, ctx->sstage.ensure(cap_bytes + 256));
  unsigned char* st = static_cast<unsigned char*>(ctx->sstage.p);
  kvm::StreamTile* tiles = reinterpret_cast<kvm::StreamTile*>(st + SP.o_tiles);
  int32_t* cb = reinterpret_cast<int32_t*>(st + SP.o_cb);
  int32_t* nc = reinterpret_cast<int32_t*>(st + SP.o_nc);
  int32_t* vb = reinterpret_cast<int32_t*>(st + SP.o_vb);
  int n_live = 0, n_tiles = 0, n_segments = 0;
  int64_t v = 0, l_max = 0;
  int open = -1;  // index of the tile still being filled
  int64_t prev_end = INT64_MIN;
  for (int p = 0; p < K; p++) {
    const int32_t c = P.ncand[p];
    if (c <= 0) continue;
    const int32_t b = P.cbegin[p];
    cb[n_live] = b;
    nc[n_live] = c;
    vb[n_live] = (int32_t)v;
    l_max = std::max<int64_t>(l_max, P.nsamp[p]);
    if ((int64_t)b != prev_end) {  // not adjacent to the previous chain: a new segment
      open = -1;
      if (n_segments++ == 0) SP.s_base = b;
    }
    int32_t pos = b, rem = c;
    while (rem > 0) {
      if (open < 0) {
        open = n_tiles++;
        tiles[open] = kvm::StreamTile{pos, 0, n_live, (int32_t)(v + (pos - b))};
      }
      const int32_t take = std::min<int32_t>(rem, W - tiles[open].nwin);
      tiles[open].nwin += take;
      pos += take;
      rem -= take;
      if (tiles[open].nwin == W) open = -1;
    }
    prev_end = (int64_t)b + c;
    v += c;
    n_live++;
  }
  SP.n_chains = n_live;
  SP.n_segments = n_segments;
  SP.n_tiles = n_tiles;
  SP.l_max = l_max;
  SP.bytes = cap_bytes;
  if (n_tiles > 0) {
    KVM_CUDA(ctx, ctx->sarena.ensure(cap_bytes + 256));
    // (tiles, cbegin, ncand, vbase live in separate 256-byte aligned blocks of one staging buffer: one copy)
    KVM_CUDA(ctx, cudaMemcpyAsync(ctx->sarena.p, ctx->sstage.p, cap_bytes, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += (long long)(sizeof(kvm::StreamTile) * (size_t)n_tiles + 3 * sizeof(int32_t) * (size_t)n_live);
  }
  SP.K = K;
  SP.shift = shift;
  SP.m = m;
  SP.nt = nt;
  if (cache_on) {
    SP.lr.assign(lr, lr + 2 * (size_t)K);
    SP.valid = true;
  }
  return KVM_OK;
}

// Launch shape of cnsm_ed_exact_kernel: warps per CTA, the per-CTA query staging and the per-warp window staging that
// fit shared memory.
template <bool kFromSums>
void launch_exact(kvm_ctx* ctx, ExactEdParams& X) {
  // Shared memory per CTA: the |zQ|-ordered query (12 m bytes, once per CTA) and per warp the term buffer (8 KB) plus,
  // for short queries, the window itself.  What matters is how many warps the SM holds: survivors are worked on one
  // warp each and every one of them costs microseconds (m dependent additions in the reference's order), so a long
  // query must not buy its staging with parallelism — with everything staged an m = 8192 query ran ONE warp per SM and
  // its exact stage took 3-4 ms for a few thousand survivors.  Window staged up to m = 1024, query up to m = 4096.
  const int warps = 4;
  X.win_cap = X.m <= 1024 ? ((X.m + 1) & ~1) : 0;
  X.q_cap = X.m <= 4096 ? ((X.m + 1) & ~1) : 0;
  const size_t bytes = sizeof(double) * ((size_t)X.q_cap + (X.q_cap + 1) / 2 + (size_t)warps * (kExactChunk + X.win_cap));
  cnsm_ed_exact_kernel<kFromSums><<<ctx->n_sms * 6, warps * 32, bytes, ctx->stream>>>(X);
}

int ensure_xlist(kvm_ctx* ctx, long long cap) {
  if (cap <= ctx->x_cap) return KVM_OK;
  ctx->x_cap = 0;
  KVM_CUDA(ctx, ctx->x_off.ensure(sizeof(int32_t) * cap));
  KVM_CUDA(ctx, ctx->x_ex.ensure(sizeof(double) * cap));
  KVM_CUDA(ctx, ctx->x_ex2.ensure(sizeof(double) * cap));
  ctx->x_cap = cap;
  return KVM_OK;
}

int verify_norm_stream(kvm_ctx* ctx, Mode mode, const double* q, int m, double epsilon, int rho, double alpha, double beta,
                       const int32_t* lr, int K, int shift, int nt, kvm_result* out, bool may_defer) {
  const auto t_begin = std::chrono::steady_clock::now();
  auto since = [&]() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count(); };
  static const int timing = env_int("KVM_TIMING", 0);
  const bool cache_on = ctx->opt_plan_cache != 0;
  int rc = stream_plan(ctx, lr, K, shift, m, nt, cache_on, may_defer);
  if (rc) return rc;
  const double t_plan = since();
  const StreamPlan& SP = ctx->splan;
  out->cnt_candidate = SP.cnt_candidate;
  out->n_verified = SP.V;
  out->s_total = SP.S;
  NormSetup S;
  query_stats(q, m, &S.meanQ, &S.stdQ);
  S.inv_alpha = 1.0 / alpha;
  S.degenerate = !(S.stdQ > 0.0) || !(S.stdQ < INFINITY);
  if (SP.V == 0 || S.degenerate) return fetch_answers(ctx, 0, out);

  // ---- query: z-normalisation (K/NormQueryEngine.java:438-441); the stream needs only the screen table, the full
  // |z| ordering (:448) is prepared while it runs
  std::vector<double> z(m);
  double zmax = 0.0;
  for (int i = 0; i < m; i++) {
    z[i] = (q[i] - S.meanQ) / S.stdQ;
    zmax = std::max(zmax, std::fabs(z[i]));
  }
  const int n_screen = std::min(m, kvm::kScreenTerms);
  std::vector<int32_t> order(m);
  for (int i = 0; i < m; i++) order[i] = i;
  std::vector<double> uq, lq;
  int32_t scr_i[kvm::kScreenTerms] = {0};
  double scr_a[kvm::kScreenTerms] = {0}, scr_b[kvm::kScreenTerms] = {0};
  if (mode == Mode::kEd) {
    // the first n_screen entries of the stable sort by |z| descending (ties: lower index first); the full sort (:448),
    // which only the exact stage needs, runs on the host while the stream runs on the device
    std::vector<int32_t> top(order);
    std::partial_sort(top.begin(), top.begin() + n_screen, top.end(), [&](int32_t a, int32_t b) {
      const int c = java_double_compare(std::fabs(z[b]), std::fabs(z[a]));
      return c < 0 || (c == 0 && a < b);
    });
    for (int k = 0; k < n_screen; k++) {
      scr_i[k] = top[k];
      scr_a[k] = scr_b[k] = z[top[k]];
    }
  } else {
    envelope(z, rho, lq, uq);  // K/NormQueryEngineDtw.java:469
    for (int k = 0; k < n_screen; k++) {
      const int i = (int)(((int64_t)(2 * k + 1) * m) / (2 * n_screen));  // evenly spread positions
      scr_i[k] = i;
      // centre and half width of the envelope (the half width rounded up by more than the rounding of the centre,
      // so that max(|x - centre| - half, 0) never exceeds the true excess over [lq, uq])
      scr_a[k] = 0.5 * (uq[i] + lq[i]);
      scr_b[k] = 0.5 * (uq[i] - lq[i]) + 4.0 * kUlpHalf * (std::fabs(uq[i]) + std::fabs(lq[i]) + 1.0);
    }
  }
  const kvm::GuardCoef GC = stream_guard(m, nt, SP.l_max, S, alpha, beta, epsilon, zmax, m);

  // ---- buffers
  const size_t words = (size_t)(SP.V + 63) / 32 + 4;
  if (words > ctx->need_words) {
    KVM_CUDA(ctx, ctx->need_bits.ensure(sizeof(unsigned) * words));
    ctx->need_words = ctx->need_bits.cap / sizeof(unsigned);
    ctx->stream_dirty = true;
  }
  if ((size_t)SP.n_chains > ctx->chain_last_n) {
    KVM_CUDA(ctx, ctx->chain_last.ensure(sizeof(int32_t) * (size_t)SP.n_chains));
    KVM_CUDA(ctx, ctx->flagged.ensure(sizeof(int32_t) * (size_t)SP.n_chains));
    ctx->chain_last_n = std::min(ctx->chain_last.cap, ctx->flagged.cap) / sizeof(int32_t);
    ctx->stream_dirty = true;
  }
  if (ctx->stream_dirty) {
    KVM_CUDA(ctx, cudaMemsetAsync(ctx->need_bits.p, 0, ctx->need_bits.cap, ctx->stream));
    KVM_CUDA(ctx, cudaMemsetAsync(ctx->chain_last.p, 0xff, ctx->chain_last.cap, ctx->stream));
  }
  ctx->stream_dirty = true;  // until this call has completed
  if ((rc = ensure_xlist(ctx, std::max<long long>(ctx->x_cap, 1 << 18)))) return rc;
  if ((rc = ensure_answers(ctx, std::max<long long>(ctx->ans_cap, 1 << 16)))) return rc;
  if ((rc = ensure_cands(ctx, std::max<long long>(ctx->cand_cap, 1 << 18)))) return rc;

  // ---- query block: [scr_idx | scr_a | scr_b | zq | order | uq | lq]; the screen part goes up before the stream
  auto up256 = [](size_t x) { return (x + 255) & ~size_t(255); };
  const size_t o_si = 0, o_sa = 256, o_sb = 512, o_zq = 768;
  const size_t o_order = up256(o_zq + sizeof(double) * (size_t)m);
  const size_t o_uq = up256(o_order + sizeof(int32_t) * (size_t)m);
  const size_t o_lq = up256(o_uq + sizeof(double) * (size_t)m);
  const size_t q_bytes = up256(o_lq + sizeof(double) * (size_t)m);
  KVM_CUDA(ctx, ctx->stage2.ensure(q_bytes + 256));
  KVM_CUDA(ctx, ctx->qarena.ensure(q_bytes + 256));
  unsigned char* qs = static_cast<unsigned char*>(ctx->stage2.p);
  const bool dtw = mode == Mode::kDtw;
  if (dtw) {  // the stream's LB_KimFL reads the natural-order query: everything goes up before it
    std::memcpy(qs + o_zq, z.data(), sizeof(double) * (size_t)m);
    std::memcpy(qs + o_uq, uq.data(), sizeof(double) * (size_t)m);
    std::memcpy(qs + o_lq, lq.data(), sizeof(double) * (size_t)m);
    KVM_CUDA(ctx, cudaMemcpyAsync(ctx->qarena.as<unsigned char>() + o_zq, qs + o_zq, q_bytes - o_zq, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d_bytes += (long long)(q_bytes - o_zq);
  }
  bool sorted_up = dtw;
  const unsigned char* qbase = ctx->qarena.as<unsigned char>();
  const unsigned char* sbase = ctx->sarena.as<unsigned char>();
  unsigned long long* counters = ctx->counters.as<unsigned long long>();

  const double eps2 = epsilon * epsilon;
  const double eps2_hi = eps2 * (1.0 + 1e-9) + 1e-18;
  const int force_all = ctx->opt_force_all;
  unsigned long long cnt[kNumCounters];
  double total_ms = 0;
  for (int attempt = 0; attempt < 8; attempt++) {
    int launches = 0;
    if ((rc = zero_counters(ctx))) return rc;
    KVM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    kvm::StreamParams W{};
    W.T = ctx->series;
    W.tiles = reinterpret_cast<const kvm::StreamTile*>(sbase + SP.o_tiles);
    W.chains.cbegin = reinterpret_cast<const int32_t*>(sbase + SP.o_cb);
    W.chains.ncand = reinterpret_cast<const int32_t*>(sbase + SP.o_nc);
    W.chains.vbase = reinterpret_cast<const int32_t*>(sbase + SP.o_vb);
    W.chains.n_chains = SP.n_chains;
    W.chains.regular = SP.regular;
    W.chains.s_base = SP.s_base;
    W.chains.chunk = SP.chunk;
    W.chains.total_win = (int32_t)SP.V;
    W.m = m;
    W.dm = (double)m;
    W.inv_m = 1.0 / (double)m;
    W.inv_m2 = W.inv_m * W.inv_m;
    W.C = GC;
    W.bmax = ctx->bmax.as<double>();
    W.n_bmax = (int)ctx->n_bmax;
    W.l_max = (int)SP.l_max;
    W.uniform = SP.n_segments == 1 ? 1 : 0;  // one run of adjacent window starts: tiles follow from the CTA index
    W.s_base = SP.s_base;
    W.total_win = (int32_t)SP.V;
    std::memcpy(W.scr_idx, scr_i, sizeof(scr_i));
    std::memcpy(W.scr_a, scr_a, sizeof(scr_a));
    std::memcpy(W.scr_b, scr_b, sizeof(scr_b));
    W.q_full = reinterpret_cast<const double*>(qbase + o_zq);
    W.order_full = reinterpret_cast<const int32_t*>(qbase + o_order);
    W.uq_full = reinterpret_cast<const double*>(qbase + o_uq);
    W.lq_full = reinterpret_cast<const double*>(qbase + o_lq);
    W.n_screen = n_screen;
    W.force_all = force_all;
    W.need_bits = ctx->need_bits.as<unsigned>();
    W.chain_last = ctx->chain_last.as<int32_t>();
    W.flagged = ctx->flagged.as<int32_t>();
    W.n_flagged = counters + kCntTiles;
    W.gate_pass = counters + kCntGate;
    W.n_need = counters + kCntDone;
    const size_t smem = kvm::stream_smem_bytes(nt, m);
#ifdef KVM_STREAM_PROF
    unsigned long long z16[16] = {0};
    cudaMemcpyToSymbolAsync(kvm::g_stream_prof, z16, sizeof(z16), 0, cudaMemcpyHostToDevice, ctx->stream);
#endif
    KVM_CUDA(ctx, dtw ? launch_stream<1>(nt, SP.n_tiles, smem, ctx->stream, W) : launch_stream<0>(nt, SP.n_tiles, smem, ctx->stream, W));
#ifdef KVM_STREAM_PROF
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpyFromSymbol(z16, kvm::g_stream_prof, sizeof(z16));
    {
      const double c = (double)std::max<unsigned long long>(z16[8], 1);
      std::fprintf(stderr, "[stream prof] nt %d tiles %llu cycles/CTA (thread 0): setup %.0f guard+tma-wait %.0f group-sums %.0f warp-local %.0f "
                   "end %.0f | chains needed %llu, warps slid %llu of %llu, candidates %llu | slowest CTA %llu cycles, %llu CTAs > 60k, slowest warp-0 local phase %llu\n", nt, z16[8], z16[0] / c, z16[1] / c, z16[2] / c,
                   z16[5] / c, z16[6] / c, z16[10], z16[9], (unsigned long long)SP.n_tiles * (nt / 32), z16[11], z16[12], z16[13], z16[14]);
    }
#endif
    KVM_CUDA(ctx, cudaEventRecord(ctx->evs[0], ctx->stream));
    kvm::RewalkParams R{};
    R.T = ctx->series;
    R.chains = W.chains;
    R.m = m;
    R.first_global = (int32_t)ctx->first;
    R.need_bits = W.need_bits;
    R.chain_last = W.chain_last;
    R.flagged = W.flagged;
    R.n_flagged = W.n_flagged;
    R.out = kvm::XList{ctx->x_off.as<int32_t>(), ctx->x_ex.as<double>(), ctx->x_ex2.as<double>(), counters + kCntEntries, ctx->x_cap};
    kvm::chain_rewalk_kernel<<<std::min(SP.n_chains, ctx->n_sms * 16), 32, 0, ctx->stream>>>(R);
    KVM_CUDA(ctx, cudaGetLastError());
    KVM_CUDA(ctx, cudaEventRecord(ctx->evs[1], ctx->stream));
    launches += 2;
    if (!sorted_up) {  // cNSM-ED: the full |z| ordering for the exact stage, while the stream and the re-walk run
      std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) {
        return java_double_compare(std::fabs(z[b]), std::fabs(z[a])) < 0;  // K/NormQueryEngine.java:448
      });
      double* zq = reinterpret_cast<double*>(qs + o_zq);
      for (int i = 0; i < m; i++) zq[i] = z[order[i]];
      std::memcpy(qs + o_order, order.data(), sizeof(int32_t) * (size_t)m);
      KVM_CUDA(ctx, cudaMemcpyAsync(ctx->qarena.as<unsigned char>() + o_zq, qs + o_zq, o_uq - o_zq, cudaMemcpyHostToDevice, ctx->stream));
      ctx->h2d_bytes += (long long)(o_uq - o_zq);
      sorted_up = true;
    }
    if (!dtw) {
      ExactEdParams X{};
      X.T = ctx->series;
      X.first_global = (int32_t)ctx->first;
      X.m = m;
      X.zq = reinterpret_cast<const double*>(qbase + o_zq);
      X.order = reinterpret_cast<const int32_t*>(qbase + o_order);
      X.eps2 = eps2;
      X.eps2_hi = eps2_hi;
      X.n_exact = counters + kCntFlag;
      X.sink = sink_of(ctx);
      X.xin = R.out;
      X.meanQ = S.meanQ;
      X.stdQ = S.stdQ;
      X.alpha = alpha;
      X.inv_alpha = S.inv_alpha;
      X.beta = beta;
      X.gate_pass = counters + kCntGate;
      X.in = cands_of(ctx);
      static const int split = env_int("KVM_EXACT_SPLIT", 1);  // developer knob: 0 = gate and screen inside the exact kernel
      if (split) {
        cnsm_ed_screen_kernel<<<ctx->n_sms * 8, 256, 0, ctx->stream>>>(X);
        launch_exact<false>(ctx, X);
        launches += 2;
      } else {
        launch_exact<true>(ctx, X);
        launches += 1;
      }
    } else {
      LbListParams L{};
      L.T = ctx->series;
      L.first_global = (int32_t)ctx->first;
      L.m = m;
      L.in = R.out;
      L.meanQ = S.meanQ;
      L.stdQ = S.stdQ;
      L.alpha = alpha;
      L.inv_alpha = S.inv_alpha;
      L.beta = beta;
      L.Q = LbQuery{reinterpret_cast<const double*>(qbase + o_zq), reinterpret_cast<const double*>(qbase + o_uq),
                    reinterpret_cast<const double*>(qbase + o_lq), m, eps2_hi};
      L.out = cands_of(ctx);
      L.gate_pass = counters + kCntGate;
      if (ensure_envelope(ctx, rho)) {  // exact gate + LB_Kim + both LB_Keogh bounds in one pass: flagged windows -> cands2
        kvm::LbFusedParams F{};
        F.T = ctx->series;
        F.envL = ctx->env_lo.as<double>() + kFrontPad;
        F.envU = ctx->env_up.as<double>() + kFrontPad;
        F.first_global = (int32_t)ctx->first;
        F.m = m;
        F.xin = R.out;
        F.meanQ = S.meanQ;
        F.stdQ = S.stdQ;
        F.alpha = alpha;
        F.inv_alpha = S.inv_alpha;
        F.beta = beta;
        F.q = L.Q.q;
        F.uq = L.Q.uq;
        F.lq = L.Q.lq;
        F.eps2_hi = eps2_hi;
        F.out = cands2_of(ctx);
        F.gate_pass = counters + kCntGate;
        F.warp_path_max = lb_warp_path_max();
        kvm::dtw_lb_fused_kernel<true, true><<<ctx->n_sms * 6, kvm::kLbThreads, 0, ctx->stream>>>(F);
        KVM_CUDA(ctx, cudaGetLastError());
      } else {
        cnsm_dtw_lb_list_kernel<<<ctx->n_sms * 8, 128, 0, ctx->stream>>>(L);
        if ((rc = launch_lb_data(ctx, L.Q.q, L.Q.uq, L.Q.lq, m, rho, eps2_hi))) return rc;
      }
      KVM_CUDA(ctx, cudaEventRecord(ctx->evs[2], ctx->stream));
      DtwParams D{};
      D.T = ctx->series;
      D.first_global = (int32_t)ctx->first;
      D.m = m;
      D.rho = rho;
      D.q = L.Q.q;
      D.uq = L.Q.uq;
      D.lq = L.Q.lq;
      D.eps2 = eps2;
      D.eps2_hi = eps2_hi;
      D.n_abandoned = counters + kCntFlag;
      D.n_cells = counters + kCntCells;
      D.in = cands2_of(ctx);
      D.sink = sink_of(ctx);
      if ((rc = launch_dtw(ctx, D))) return rc;
      launches += 3;
    }
    KVM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    KVM_CUDA(ctx, cudaGetLastError());
    const double t_launched = since();
    if (ctx->splan.unchecked) {  // the deferred pass over the interval list, while the kernels run
      ctx->splan.unchecked = false;
      if (regular_grid_pass(lr, K, (int32_t)SP.chunk)) {  // not a regular grid after all: discard this attempt, plan in full
        KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->splan.valid = false;
        ctx->stream_dirty = true;
        const kvm_result zero{};
        *out = zero;
        return verify_norm_stream(ctx, mode, q, m, epsilon, rho, alpha, beta, lr, K, shift, nt, out, false);
      }
    }
    if ((rc = read_counters(ctx, cnt))) return rc;
    if (timing) std::fprintf(stderr, "[kvm stream] plan %.0f us, launched %.0f, synced %.0f (K %d, regular %d)\n", t_plan, t_launched, since(), K, SP.regular);
    total_ms += elapsed_ms(ctx);
    {
      float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
      cudaEventElapsedTime(&a, ctx->ev0, ctx->evs[0]);
      cudaEventElapsedTime(&b, ctx->evs[0], ctx->evs[1]);
      if (dtw) {
        cudaEventElapsedTime(&c, ctx->evs[1], ctx->evs[2]);
        cudaEventElapsedTime(&d, ctx->evs[2], ctx->ev1);
      } else {
        cudaEventElapsedTime(&c, ctx->evs[1], ctx->ev1);
      }
      out->stage_ms[0] += a;
      out->stage_ms[1] += b;
      out->stage_ms[2] += c;
      out->stage_ms[3] += d;
    }
    out->n_launches += launches;
    const bool x_over = (long long)cnt[kCntEntries] > ctx->x_cap;
    const long long cand_need = (long long)std::max(cnt[kCntCand], cnt[kCntCand2]);
    const bool cand_over = cand_need > ctx->cand_cap, ans_over = (long long)cnt[kCntAnswers] > ctx->ans_cap;
    if (!x_over && !cand_over && !ans_over) break;
    if (attempt == 7) return fail(ctx, KVM_E_OOM, "result buffers kept overflowing");
    if (x_over && (rc = ensure_xlist(ctx, (long long)cnt[kCntEntries] + 1024))) return rc;
    if (cand_over && (rc = ensure_cands(ctx, cand_need + 1024))) return rc;
    if (ans_over && (rc = ensure_answers(ctx, (long long)cnt[kCntAnswers] + 1024))) return rc;
  }
  ctx->stream_dirty = false;  // the re-walk consumed every flag it was given
  out->kernel_ms = total_ms;
  out->n_gate_pass = (int64_t)cnt[kCntGate];
  out->n_rewalked = (int64_t)cnt[kCntEntries];
  out->n_chains_rewalked = (int64_t)cnt[kCntTiles];
  if (!dtw) out->n_exact = (int64_t)cnt[kCntFlag];
  else {
    out->n_lb_pass = (int64_t)cnt[kCntCand2];
    out->n_exact = (int64_t)cnt[ctx->dtw_final_counter];  // DTW engines: candidates that reached the band DTW (after the corner probe)
  }
  out->n_dtw_cells = (int64_t)cnt[kCntCells];
  return fetch_answers(ctx, (long long)cnt[kCntAnswers], out);
}

}  // namespace

// Data envelope of the whole resident buffer (pads included: they hold the zeros some scans count as samples) for
// radius rho, cached per ctx.  Returns false (and remembers it) when the 16 bytes per sample do not fit.
namespace {
int lb_warp_path_max() {
  static const int v = env_int("KVM_LB_WARP_MAX", kvm::kLbWarpPathMax);  // developer knob (0: every list takes the cohort path)
  return v;
}

bool ensure_envelope(kvm_ctx* ctx, int rho) {
  static const int fused_on = env_int("KVM_LB_FUSED", 1);  // developer knob: 0 = the per-candidate kernels
  if (!fused_on || rho > 512) return false;
  if (ctx->env_rho == rho) return true;
  if (ctx->env_failed) return false;
  const size_t len = (size_t)ctx->series_len;
  ctx->env_rho = -1;
  if (ctx->env_lo.ensure(sizeof(double) * len) != cudaSuccess || ctx->env_up.ensure(sizeof(double) * len) != cudaSuccess) {
    cudaGetLastError();
    ctx->env_lo.release();
    ctx->env_up.release();
    ctx->env_failed = true;
    return false;
  }
  const size_t smem = sizeof(long long) * 2 * (size_t)(kvm::kEnvTile + 2 * rho);
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(kvm::envelope_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    attr = true;
  }
  kvm::envelope_kernel<<<(unsigned)((len + kvm::kEnvTile - 1) / kvm::kEnvTile), 256, smem, ctx->stream>>>(
      ctx->series_buf.as<double>(), (int)len, rho, ctx->env_lo.as<double>(), ctx->env_up.as<double>());
  if (cudaGetLastError() != cudaSuccess) return false;
  ctx->env_rho = rho;
  return true;
}

// Third bound of the cascade (LB_Keogh on the data envelope) over the survivors of the first LB stage: cands -> cands2.
int launch_lb_data(kvm_ctx* ctx, const double* q, const double* uq, const double* lq, int m, int rho, double eps2_hi) {
  if (ensure_envelope(ctx, rho)) {
    kvm::LbFusedParams F{};
    F.T = ctx->series;
    F.envL = ctx->env_lo.as<double>() + kFrontPad;
    F.envU = ctx->env_up.as<double>() + kFrontPad;
    F.first_global = (int32_t)ctx->first;
    F.m = m;
    F.cin = cands_of(ctx);
    F.q = q;
    F.uq = uq;
    F.lq = lq;
    F.eps2_hi = eps2_hi;
    F.out = cands2_of(ctx);
    F.warp_path_max = lb_warp_path_max();
    kvm::dtw_lb_fused_kernel<false, false><<<ctx->n_sms * 6, kvm::kLbThreads, 0, ctx->stream>>>(F);
    KVM_CUDA(ctx, cudaGetLastError());
    return KVM_OK;
  }
  Lb2Params L{};
  L.T = ctx->series;
  L.first_global = (int32_t)ctx->first;
  L.m = m;
  L.rho = rho;
  L.q = q;
  L.eps2_hi = eps2_hi;
  L.in = cands_of(ctx);
  L.out = cands2_of(ctx);
  const size_t per_warp = lb2_warp_bytes(m);
  const size_t budget = 200 * 1024;
  if (per_warp > budget) return fail(ctx, KVM_E_ARG, "DTW query length %d exceeds the shared-memory staging limit", m);
  const int warps = (int)std::max<size_t>(1, std::min<size_t>(8, budget / per_warp));
  const size_t smem = per_warp * warps;
  static bool attr = false;
  if (!attr) {
    KVM_CUDA(ctx, cudaFuncSetAttribute(dtw_lb_data_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
    attr = true;
  }
  const int grid = ctx->n_sms * std::max(1, (int)(budget / smem));
  dtw_lb_data_kernel<<<grid, warps * 32, smem, ctx->stream>>>(L);
  KVM_CUDA(ctx, cudaGetLastError());
  return KVM_OK;
}
}  // namespace

int launch_dtw(kvm_ctx* ctx, const DtwParams& D_in) {
  DtwParams D = D_in;
  // Corner probe (dtw_probe_kernel): survivors of the lower bounds whose K x K corner already exceeds eps^2 never reach
  // the wavefront kernels.  Needs the list that carries the Keogh totals (the fused LB stage's output); its survivors
  // go to the first candidate list, which no stage reads any more at this point.
  {
    static const int probe_on = env_int("KVM_DTW_PROBE", 1);
    static const int probe_k = env_int("KVM_DTW_PROBE_K", kvm::kProbeMaxK);  // developer knob
    const int K = std::min(std::min(D.rho + 1, std::min(probe_k, kvm::kProbeMaxK)), D.m);
    if (probe_on && D.in.lb != nullptr && K >= 24) {
      kvm::ProbeParams PP{};
      PP.T = D.T;
      PP.first_global = D.first_global;
      PP.m = D.m;
      PP.K = K;
      PP.q = D.q;
      PP.uq = D.uq;
      PP.lq = D.lq;
      PP.eps2_hi = D.eps2_hi;
      PP.n_cells = D.n_cells;
      static const int probe_min = env_int("KVM_DTW_PROBE_MIN", 4096);  // developer knob
      PP.min_count = probe_min;
      unsigned long long* counters = ctx->counters.as<unsigned long long>();
      const CandList first_list{ctx->cand_off.as<int32_t>(), ctx->cand_mean.as<double>(), ctx->cand_std.as<double>(),
                                counters + kCntProbe, ctx->cand_cap, ctx->cand_lb.as<double>()};
      static const int two_stage = env_int("KVM_DTW_PROBE64", 1);  // developer knob
      if (two_stage && K >= 96) {
        // wide bands: the 64 x 64 sub-square first, two candidates per warp; its survivors through the full square
        PP.in = D.in;
        PP.out = first_list;
        PP.handed_on = counters + 11;  // (a free counter slot, zeroed with the others)
        kvm::dtw_probe64_kernel<<<ctx->n_sms * 8, kvm::kProbeWarps * 32, 0, ctx->stream>>>(PP);
        KVM_CUDA(ctx, cudaGetLastError());
        PP.in = first_list;
        PP.out = CandList{ctx->cand2_off.as<int32_t>(), ctx->cand2_mean.as<double>(), ctx->cand2_std.as<double>(),
                          counters + kCntProbe2, ctx->cand_cap};
        PP.skip_sub = 1;
        PP.min_count = 0;  // (a short list was already handed on by the first stage; what arrives here is probed)
        kvm::dtw_probe_kernel<<<ctx->n_sms * 8, kvm::kProbeWarps * 32, 0, ctx->stream>>>(PP);
        KVM_CUDA(ctx, cudaGetLastError());
        ctx->dtw_final_counter = kCntProbe2;
      } else {
        PP.in = D.in;
        PP.out = first_list;
        kvm::dtw_probe_kernel<<<ctx->n_sms * 8, kvm::kProbeWarps * 32, 0, ctx->stream>>>(PP);
        KVM_CUDA(ctx, cudaGetLastError());
        ctx->dtw_final_counter = kCntProbe;
      }
      D.in = PP.out;
    }
  }
  const int need = (D.rho + 1 + 31) / 32;  // (even,odd) pairs per lane
  const size_t bytes_per_row = sizeof(double) * (size_t)D.m;
  const size_t warp_bytes = sizeof(double) * (size_t)(D.m + ((((D.m >> 3) + 2) + 1) & ~1));  // window + sampled cb
  const size_t smem_budget = 200 * 1024;
  if (bytes_per_row + warp_bytes > smem_budget)
    return fail(ctx, KVM_E_ARG, "DTW query length %d exceeds the shared-memory staging limit (%zu)", D.m,
                smem_budget / 17);
  // Wide bands (two or more cell pairs per lane of a warp) run on the CTA-cooperative kernel, one DTW per CTA at a
  // time: 0.33 ms per m = 2048 / rho = 102 DTW against 0.52 ms for one warp, and more DTWs per SM per ms as well
  // (five 4-warp CTAs per SM).  Narrow bands stay on the warp-per-candidate kernel.  KVM_DTW_COOP=0 disables.
  {
    const size_t coop_smem = warp_bytes;  // window + sampled cumulative bound (the query is read through L1)
    const int per_sm = std::max(1, (int)std::min<size_t>(smem_budget / coop_smem, 9));
    const int need_c = (D.rho + 1 + kCoopThreads - 1) / kCoopThreads;
    static const int coop_on = env_int("KVM_DTW_COOP", 1);
    D.coop_limit = 0;
    if (coop_on && need_c <= 4 && D.rho + 1 >= 64) {
      D.coop_limit = (long long)1 << 40;  // every candidate
      const int grid = ctx->n_sms * per_sm;
#define KVM_COOP_CASE(R)                                                                                                 \
  if (need_c <= R) {                                                                                                     \
    KVM_CUDA(ctx, cudaFuncSetAttribute(dtw_band_coop_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)coop_smem)); \
    dtw_band_coop_kernel<R><<<grid, kCoopThreads, coop_smem, ctx->stream>>>(D);                                         \
  } else
      KVM_COOP_CASE(1)
      KVM_COOP_CASE(2)
      KVM_COOP_CASE(4) {}
#undef KVM_COOP_CASE
      KVM_CUDA(ctx, cudaGetLastError());
      return KVM_OK;
    }
  }
  const int warps = (int)std::min<size_t>(8, (smem_budget - bytes_per_row) / warp_bytes);
  const size_t smem = bytes_per_row + warp_bytes * warps;
  const int grid = ctx->n_sms * std::max(1, (int)(smem_budget / smem));
#define KVM_DTW_CASE(R)                                                                                        \
  if (need <= R) {                                                                                             \
    KVM_CUDA(ctx, cudaFuncSetAttribute(dtw_band_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    dtw_band_kernel<R><<<grid, warps * 32, smem, ctx->stream>>>(D);                                            \
    return KVM_OK;                                                                                             \
  }
  KVM_DTW_CASE(1)
  KVM_DTW_CASE(2)
  KVM_DTW_CASE(3)
  KVM_DTW_CASE(4)
  KVM_DTW_CASE(6)
  KVM_DTW_CASE(8)
  KVM_DTW_CASE(10)
  KVM_DTW_CASE(13)
  KVM_DTW_CASE(16)
#undef KVM_DTW_CASE
  return fail(ctx, KVM_E_ARG, "Sakoe-Chiba radius %d exceeds the supported maximum 511", D.rho);
}

extern "C" {

int kvm_abi_version(void) { return KVM_ABI_VERSION; }

const char* kvm_last_error(const kvm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int kvm_create(kvm_ctx** out, int device_id) {
  if (!out) return fail(nullptr, KVM_E_ARG, "out is null");
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return fail(nullptr, KVM_E_NODEVICE, "no CUDA device (%s); this library has no CPU fallback",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device_id < 0 || device_id >= n_dev) return fail(nullptr, KVM_E_ARG, "device %d of %d", device_id, n_dev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device_id) != cudaSuccess || prop.major != 10)
    return fail(nullptr, KVM_E_NODEVICE, "device %d is not compute capability 10.x (B200); kernels are sm_100a only",
                device_id);
  kvm_ctx* ctx = new kvm_ctx();
  ctx->device = device_id;
  ctx->n_sms = prop.multiProcessorCount;
  {
    const char* e = std::getenv("KVM_CNSM_PATH");
    ctx->opt_relay = (e && std::strcmp(e, "relay") == 0) ? 1 : 0;
    ctx->opt_force_all = env_int("KVM_STREAM_FORCE_ALL", 0);
    ctx->opt_plan_cache = env_int("KVM_PLAN_CACHE", 1);
  }
  if (cudaSetDevice(device_id) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess ||
      cudaEventCreate(&ctx->evs[0]) != cudaSuccess || cudaEventCreate(&ctx->evs[1]) != cudaSuccess ||
      cudaEventCreate(&ctx->evs[2]) != cudaSuccess || cudaEventCreate(&ctx->evs[3]) != cudaSuccess) {
    const char* msg = cudaGetErrorString(cudaGetLastError());
    delete ctx;
    return fail(nullptr, KVM_E_CUDA, "stream/event creation failed: %s", msg);
  }
  const int r4 = (int)relay_smem_bytes(kRelayStages);
  if (cudaFuncSetAttribute(cnsm_relay_kernel<kRelayStages, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, r4) != cudaSuccess ||
      cudaFuncSetAttribute(cnsm_relay_kernel<kRelayStages, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, r4) != cudaSuccess ||
      cudaFuncSetAttribute(cnsm_relay_kernel<kRelayStages, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, r4) != cudaSuccess ||
      cudaFuncSetAttribute(cnsm_relay_kernel<kRelayStages, 0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, r4) != cudaSuccess ||
      cudaFuncSetAttribute(cnsm_relay_kernel<kRelayStages, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, r4) != cudaSuccess ||
      cudaFuncSetAttribute(cnsm_relay_kernel<kRelayStages, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, r4) != cudaSuccess) {
    const char* msg = cudaGetErrorString(cudaGetLastError());
    kvm_destroy(ctx);
    return fail(nullptr, KVM_E_CUDA, "cudaFuncSetAttribute(relay walker) failed: %s", msg);
  }
  if (cudaFuncSetAttribute(cnsm_ed_exact_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
      cudaFuncSetAttribute(cnsm_ed_exact_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess ||
      set_stream_attrs() != 0) {
    const char* msg = cudaGetErrorString(cudaGetLastError());
    kvm_destroy(ctx);
    return fail(nullptr, KVM_E_CUDA, "cudaFuncSetAttribute failed: %s", msg);
  }
  *out = ctx;
  return KVM_OK;
}

int kvm_set_option(kvm_ctx* ctx, int32_t option, int64_t value) {
  if (!ctx) return KVM_E_ARG;
  switch (option) {
    case KVM_OPT_CNSM_PATH:
      if (value != KVM_CNSM_STREAM && value != KVM_CNSM_RELAY) return fail(ctx, KVM_E_ARG, "cnsm path %lld", (long long)value);
      ctx->opt_relay = value == KVM_CNSM_RELAY;
      return KVM_OK;
    case KVM_OPT_STREAM_FLAG_ALL:
      ctx->opt_force_all = value != 0;
      return KVM_OK;
    case KVM_OPT_PLAN_CACHE:
      ctx->opt_plan_cache = value != 0;
      ctx->norm_cache.valid = false;
      ctx->splan.valid = false;
      return KVM_OK;
    default:
      return fail(ctx, KVM_E_ARG, "unknown option %d", (int)option);
  }
}

void kvm_destroy(kvm_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  DevBuf* dev[] = {&ctx->series_buf, &ctx->arena, &ctx->qarena, &ctx->counters, &ctx->wl_off, &ctx->wl_ex, &ctx->wl_ex2,
                   &ctx->region_count, &ctx->tile_prefix, &ctx->cand_off, &ctx->cand_mean, &ctx->cand_std, &ctx->cand2_off, &ctx->cand2_mean, &ctx->cand2_std,
                   &ctx->ans_off, &ctx->ans_dist, &ctx->seg_b, &ctx->seg_first, &ctx->seg_last, &ctx->chain_count,
                   &ctx->chain_prefix, &ctx->run_key, &ctx->run_b, &ctx->run_first, &ctx->run_last, &ctx->sarena,
                   &ctx->need_bits, &ctx->chain_last, &ctx->flagged, &ctx->x_off, &ctx->x_ex, &ctx->x_ex2, &ctx->bmax,
                   &ctx->cand2_lb, &ctx->cand_lb, &ctx->env_lo, &ctx->env_up, &ctx->g_send, &ctx->g_recv};
  for (DevBuf* b : dev) b->release();
  kvm_comm_release(ctx);
  ctx->g_hsend.release();
  ctx->g_hrecv.release();
  for (BatchSlot& sl : ctx->slots) sl.release();
  for (auto& w : ctx->wm) {
    DevBuf* wb[] = {&w.runs, &w.seg_off, &w.seg_cnt, &w.need_bits, &w.chain_last, &w.flagged, &w.x_off, &w.x_ex, &w.x_ex2};
    for (DevBuf* b : wb) b->release();
  }
  ctx->wm_counters.release();
  ctx->wm_host.release();
  for (cudaStream_t st : ctx->aux)
    if (st) cudaStreamDestroy(st);
  if (ctx->ev_set) cudaEventDestroy(ctx->ev_set);
  ctx->batch_gate.release();
  PinBuf* pin[] = {&ctx->stage, &ctx->stage2, &ctx->h_counters, &ctx->h_off, &ctx->h_dist, &ctx->h_key, &ctx->h_first, &ctx->h_last,
                   &ctx->h_b, &ctx->sstage};
  for (PinBuf* b : pin) b->release();
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  for (cudaEvent_t e : ctx->evs)
    if (e) cudaEventDestroy(e);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

static int alloc_series(kvm_ctx* ctx, int64_t n, int64_t first, int64_t count) {
  if (n < 1 || first < 1 || count < 1 || first + count - 1 > n)
    return fail(ctx, KVM_E_ARG, "series range first=%lld count=%lld n=%lld", (long long)first, (long long)count,
                (long long)n);
  if (n > 2147483647LL) return fail(ctx, KVM_E_ARG, "n exceeds the reference's int32 offsets");
  KVM_CUDA(ctx, cudaSetDevice(ctx->device));
  // Zero pads on both sides: (a) the walkers' whole-row TMA copies may start up to one tile before the
  // first sample / end one tile after the last; (b) the index-build path reproduces the reference's zero
  // padding of the last 1000-byte block (K/operator/file/TimeSeriesNodeIterator.java:55-59).
  ctx->series = nullptr;
  ctx->norm_cache.valid = false;  // plans hold shard-relative indices
  ctx->env_rho = -1;
  ctx->env_failed = false;
  KVM_CUDA(ctx, ctx->series_buf.ensure(sizeof(double) * (size_t)(count + kFrontPad + kTailPad)));
  ctx->series = ctx->series_buf.as<double>() + kFrontPad;
  ctx->series_len = count + kFrontPad + kTailPad;
  KVM_CUDA(ctx, cudaMemsetAsync(ctx->series_buf.p, 0, sizeof(double) * kFrontPad, ctx->stream));
  KVM_CUDA(ctx, cudaMemsetAsync(ctx->series + count, 0, sizeof(double) * kTailPad, ctx->stream));
  ctx->n = n;
  ctx->first = first;
  ctx->count = count;
  return KVM_OK;
}

// Block-maximum table of |sample| (and the global maximum): scales the guard bands of the streaming statistics pass
static int measure_absmax(kvm_ctx* ctx) {
  ctx->absmax = NAN;
  ctx->splan.valid = false;
  ctx->n_bmax = (ctx->count + kvm::kBmaxBlock - 1) / kvm::kBmaxBlock;
  KVM_CUDA(ctx, ctx->bmax.ensure(sizeof(double) * (size_t)ctx->n_bmax));
  KVM_CUDA(ctx, ctx->counters.ensure(sizeof(unsigned long long) * kNumCounters));
  KVM_CUDA(ctx, ctx->h_counters.ensure(sizeof(unsigned long long) * kNumCounters));
  KVM_CUDA(ctx, cudaMemsetAsync(ctx->counters.p, 0, sizeof(unsigned long long), ctx->stream));
  blockmax_kernel<<<ctx->n_sms * 8, 256, 0, ctx->stream>>>(ctx->series, (long long)ctx->count, ctx->bmax.as<double>(),
                                                          ctx->n_bmax, ctx->counters.as<unsigned long long>());
  KVM_CUDA(ctx, cudaGetLastError());
  KVM_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters.p, ctx->counters.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
  KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::memcpy(&ctx->absmax, ctx->h_counters.p, sizeof(double));
  return KVM_OK;
}

int kvm_load_series_host(kvm_ctx* ctx, const double* samples, int64_t n, int64_t first, int64_t count) {
  if (!ctx) return KVM_E_ARG;
  if (!samples) return fail(ctx, KVM_E_ARG, "samples is null");
  int rc = alloc_series(ctx, n, first, count);
  if (rc) return rc;
  KVM_CUDA(ctx, cudaMemcpyAsync(ctx->series, samples, sizeof(double) * (size_t)count, cudaMemcpyHostToDevice,
                                ctx->stream));
  KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return measure_absmax(ctx);
}

int kvm_load_series_file(kvm_ctx* ctx, const char* path, int64_t n, int64_t first, int64_t count) {
  if (!ctx) return KVM_E_ARG;
  if (!path) return fail(ctx, KVM_E_ARG, "path is null");
  int rc = alloc_series(ctx, n, first, count);
  if (rc) return rc;
  FILE* f = std::fopen(path, "rb");
  if (!f) return fail(ctx, KVM_E_IO, "cannot open %s", path);
  const size_t chunk = size_t(8) << 20;  // doubles per staged block (64 MiB)
  if (cudaSuccess != ctx->stage.ensure(sizeof(double) * std::min<size_t>(chunk, (size_t)count))) {
    std::fclose(f);
    return fail(ctx, KVM_E_OOM, "pinned staging allocation failed");
  }
  if (fseeko(f, (off_t)(8 * (first - 1)), SEEK_SET) != 0) {
    std::fclose(f);
    return fail(ctx, KVM_E_IO, "seek failed in %s", path);
  }
  int64_t done = 0;
  while (done < count) {
    const size_t c = (size_t)std::min<int64_t>((int64_t)chunk, count - done);
    if (std::fread(ctx->stage.p, 8, c, f) != c) {
      std::fclose(f);
      return fail(ctx, KVM_E_IO, "%s is shorter than %lld doubles", path, (long long)(first - 1 + count));
    }
    cudaError_t e = cudaMemcpyAsync(ctx->series + done, ctx->stage.p, 8 * c, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
      std::fclose(f);
      return fail(ctx, KVM_E_CUDA, "upload failed: %s", cudaGetErrorString(e));
    }
    done += (int64_t)c;
  }
  std::fclose(f);
  bswap64_kernel<<<ctx->n_sms * 8, 256, 0, ctx->stream>>>(reinterpret_cast<unsigned long long*>(ctx->series), (long long)count);
  KVM_CUDA(ctx, cudaGetLastError());
  KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return measure_absmax(ctx);
}

static int verify_ed_impl(kvm_ctx* ctx, const double* q, int32_t m, double epsilon, const int32_t* lr, int32_t K, int32_t shift,
                          kvm_result* out, bool may_defer);

int kvm_verify_ed(kvm_ctx* ctx, const double* q, int32_t m, double epsilon, const int32_t* lr, int32_t K, int32_t shift,
                  kvm_result* out) {
  return verify_ed_impl(ctx, q, m, epsilon, lr, K, shift, out, true);
}

static int verify_ed_impl(kvm_ctx* ctx, const double* q, int32_t m, double epsilon, const int32_t* lr, int32_t K, int32_t shift,
                          kvm_result* out, bool may_defer) {
  const int32_t K_list = K;
  bool unchecked = false;  // regular by its two ends only: the pass over the whole list runs behind the kernel launch
  int32_t c0_list = 0;
  int rc = check_common(ctx, q, m, epsilon, lr, K, out);
  if (rc) return rc;
  if ((rc = begin_call(ctx))) return rc;
  // The plan is built in place in the pinned staging block ([q | cbegin | ncand | tile prefix]): with 1e5 short
  // intervals (the index-pruned shape) the host side otherwise costs more than the kernel.
  auto up256 = [](size_t x) { return (x + 255) & ~size_t(255); };
  const size_t o_q = 0;
  const size_t o_cbegin = up256(sizeof(double) * (size_t)m);
  const size_t o_ncand = up256(o_cbegin + sizeof(int32_t) * (size_t)K);
  const size_t o_tp = up256(o_ncand + sizeof(int32_t) * (size_t)K);
  const size_t bytes = o_tp + sizeof(int32_t) * ((size_t)K + 1);
  ctx->norm_cache.valid = false;  // the arena is about to be overwritten
  KVM_CUDA(ctx, ctx->stage.ensure(bytes + 256));
  KVM_CUDA(ctx, ctx->arena.ensure(bytes + 256));
  unsigned char* st = static_cast<unsigned char*>(ctx->stage.p);
  int32_t* cb = reinterpret_cast<int32_t*>(st + o_cbegin);
  int32_t* nc = reinterpret_cast<int32_t*>(st + o_ncand);
  int32_t* tp = reinterpret_cast<int32_t*>(st + o_tp);
  Plan& P = ctx->plan_scratch;
  // An index-free scan arrives as a regular grid (adjacent intervals of one length, nothing clamped): recognised in one
  // vectorised pass over the list, it is ONE run of window starts and needs no per-interval planning.
  bool regular = false;
  if (K >= 2) {
    const int64_t c0 = (int64_t)lr[1] - lr[0] + 1;
    const int64_t first_begin = (int64_t)lr[0] - shift, last_end = (int64_t)lr[2 * K - 1] - shift + m - 1;
    const int64_t lo = ctx->first, hi = ctx->first + ctx->count - 1;
    bool ends_ok = c0 >= 1 && c0 <= INT32_MAX / 2 && first_begin >= 1 && first_begin >= lo && last_end <= ctx->n && last_end <= hi;
    if (ends_ok && may_defer && K >= 4096) {  // (as in the streaming path: check the list while the kernel runs)
      const int64_t V0 = (int64_t)lr[2 * K - 1] - lr[0] + 1;
      unchecked = V0 > (int64_t)(K - 1) * c0 && V0 <= (int64_t)K * c0 && (int64_t)lr[2 * (K - 1)] == (int64_t)lr[0] + (int64_t)(K - 1) * c0;
      c0_list = (int32_t)c0;
    }
    if (ends_ok && !unchecked) ends_ok = !regular_grid_pass(lr, K, (int32_t)c0);
    if (ends_ok) {
      const int64_t V = (int64_t)lr[2 * K - 1] - lr[0] + 1;
      if (V <= 0x7fff0000LL) {
        regular = true;
        cb[0] = (int32_t)(first_begin - lo);
        nc[0] = (int32_t)V;
        P.cnt_candidate = V;
        P.V = V;
        P.S = V + (int64_t)K * (m - 1);
      }
    }
  }
  if (!regular) {
    P.nsamp.resize(K);
    int64_t totals[3];
    if (plan_pass(lr, K, shift, m, ctx->n, ctx->first, ctx->first + ctx->count - 1, cb, P.nsamp.data(), nc, totals))
      return make_plan(ctx, lr, K, shift, m, &P);  // locates the offending interval and sets the error
    P.cnt_candidate = totals[0];
    P.S = totals[1];
    P.V = totals[2];
  }
  out->cnt_candidate = P.cnt_candidate;
  out->n_verified = P.V;
  out->s_total = P.S;
  if (P.V == 0) return fetch_answers(ctx, 0, out);
  // adjacent intervals are one run for this engine: re-pack the plan for the K2 runs
  K = regular ? 1 : coalesce_runs(cb, nc, K);
  const size_t o_ncand2 = up256(o_cbegin + sizeof(int32_t) * (size_t)K);
  const size_t o_tp2 = up256(o_ncand2 + sizeof(int32_t) * (size_t)K);
  const size_t bytes2 = o_tp2 + sizeof(int32_t) * ((size_t)K + 1);
  std::memmove(st + o_ncand2, nc, sizeof(int32_t) * (size_t)K);  // (o_ncand2 <= o_ncand: moving towards the front)
  nc = reinterpret_cast<int32_t*>(st + o_ncand2);
  tp = reinterpret_cast<int32_t*>(st + o_tp2);
  int64_t n_tiles = 0;
  for (int p = 0; p < K; p++) {
    tp[p] = (int32_t)n_tiles;
    n_tiles += (nc[p] + kEdTile - 1) / kEdTile;
  }
  tp[K] = (int32_t)n_tiles;
  std::memcpy(st + o_q, q, sizeof(double) * (size_t)m);
  KVM_CUDA(ctx, cudaMemcpyAsync(ctx->arena.p, ctx->stage.p, bytes2, cudaMemcpyHostToDevice, ctx->stream));
  ctx->h2d_bytes += (long long)bytes2;
  if ((rc = ensure_answers(ctx, std::max<long long>(ctx->ans_cap, 1 << 16)))) return rc;
  const unsigned char* base = ctx->arena.as<unsigned char>();
  unsigned long long cnt[kNumCounters];
  for (int attempt = 0; attempt < 8; attempt++) {
    if ((rc = zero_counters(ctx))) return rc;
    EdParams E;
    E.T = ctx->series;
    E.q = reinterpret_cast<const double*>(base + o_q);
    E.cbegin = reinterpret_cast<const int32_t*>(base + o_cbegin);
    E.ncand = reinterpret_cast<const int32_t*>(base + o_ncand2);
    E.tile_prefix = reinterpret_cast<const int32_t*>(base + o_tp2);
    E.K = K;
    E.m = m;
    E.eps2 = epsilon * epsilon;
    E.first_global = (int32_t)ctx->first;
    static const int ed_ahead = env_int("KVM_ED_AHEAD", kvm::kEdAhead);  // developer knob
    E.ahead = ed_ahead > 0 ? ed_ahead : (1 << 30);
    E.sink = sink_of(ctx);
    KVM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    ed_verify_kernel<<<(unsigned)n_tiles, kEdThreads, 0, ctx->stream>>>(E);
    KVM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    KVM_CUDA(ctx, cudaGetLastError());
    if (unchecked) {
      unchecked = false;
      if (regular_grid_pass(lr, K_list, c0_list)) {  // not a regular grid after all: discard, plan interval by interval
        KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return verify_ed_impl(ctx, q, m, epsilon, lr, K_list, shift, out, false);
      }
    }
    if ((rc = read_counters(ctx, cnt))) return rc;
    out->kernel_ms += elapsed_ms(ctx);
    out->stage_ms[0] += elapsed_ms(ctx);
    out->n_launches += 1;
    if ((long long)cnt[kCntAnswers] <= ctx->ans_cap) break;
    if (attempt == 7) return fail(ctx, KVM_E_OOM, "answer buffer kept overflowing");
    if ((rc = ensure_answers(ctx, (long long)cnt[kCntAnswers] + 1024))) return rc;
  }
  out->n_exact = P.V;
  return fetch_answers(ctx, (long long)cnt[kCntAnswers], out);
}

int kvm_verify_cnsm_ed(kvm_ctx* ctx, const double* q, int32_t m, double epsilon, double alpha, double beta,
                       const int32_t* lr, int32_t K, int32_t shift, kvm_result* out) {
  return verify_norm(ctx, Mode::kEd, q, m, epsilon, 0, alpha, beta, lr, K, shift, out);
}

namespace {
void swap_slot(kvm_ctx* c, BatchSlot& s) {
  std::swap(c->wl_off, s.wl_off);
  std::swap(c->wl_ex, s.wl_ex);
  std::swap(c->wl_ex2, s.wl_ex2);
  std::swap(c->region_count, s.region_count);
  std::swap(c->tile_prefix, s.tile_prefix);
  std::swap(c->counters, s.counters);
  std::swap(c->qarena, s.qarena);
  std::swap(c->cand_off, s.cand_off);
  std::swap(c->cand_mean, s.cand_mean);
  std::swap(c->cand_std, s.cand_std);
  std::swap(c->ans_off, s.ans_off);
  std::swap(c->ans_dist, s.ans_dist);
  std::swap(c->cand_cap, s.cand_cap);
  std::swap(c->ans_cap, s.ans_cap);
  std::swap(c->h_counters, s.h_counters);
  std::swap(c->h_off, s.h_off);
  std::swap(c->h_dist, s.h_dist);
  std::swap(c->eager_valid, s.eager_valid);
}
}  // namespace

// Query set: Q queries of one length over one interval list.  One statistics pass (cnsm_relay_kernel, kMode 2) gates all
// of them; the evaluator / exact stages then run per query on that query's work list.  Results are what Q calls of
// kvm_verify_cnsm_ed would return (same kernels, same arithmetic); outs[q].offsets / distances stay valid until the next
// batch call on this ctx.
int kvm_verify_cnsm_ed_batch(kvm_ctx* ctx, const double* queries, int32_t n_queries, int32_t m, double epsilon, double alpha,
                             double beta, const int32_t* lr, int32_t K, int32_t shift, kvm_result* outs) {
  if (!ctx) return KVM_E_ARG;
  if (!queries || !outs || n_queries < 1 || n_queries > kMaxBatch)
    return fail(ctx, KVM_E_ARG, "a query set holds 1..%d queries", kMaxBatch);
  const int Q = n_queries;
  int rc = check_common(ctx, queries, m, epsilon, lr, K, &outs[0]);
  if (rc) return rc;
  for (int q = 0; q < Q; q++) std::memset(&outs[q], 0, sizeof(kvm_result));
  if ((rc = begin_call(ctx))) return rc;
  // ---- plan (shared): the single-query engine's plan, built in place in pinned staging
  ctx->norm_cache.valid = false;
  Plan& P = ctx->plan_scratch;
  if ((rc = make_plan(ctx, lr, K, shift, m, &P))) return rc;
  const int n_regions = (K + 31) / 32;
  auto up256 = [](size_t x) { return (x + 255) & ~size_t(255); };
  const size_t o_cbegin = 0, o_nsamp = up256(sizeof(int32_t) * (size_t)K);
  const size_t o_rbase = up256(o_nsamp + sizeof(int32_t) * (size_t)K);
  const size_t o_gate = up256(o_rbase + sizeof(long long) * (size_t)(n_regions + 1));
  const size_t bytes = o_gate + sizeof(BatchGate);
  KVM_CUDA(ctx, ctx->stage.ensure(bytes + 256));
  KVM_CUDA(ctx, ctx->arena.ensure(bytes + 256));
  unsigned char* st = static_cast<unsigned char*>(ctx->stage.p);
  std::memcpy(st + o_cbegin, P.cbegin.data(), sizeof(int32_t) * (size_t)K);
  {
    int32_t* walk_nsamp = reinterpret_cast<int32_t*>(st + o_nsamp);
    long long* region_base = reinterpret_cast<long long*>(st + o_rbase);
    long long acc = 0;
    for (int c = 0; c < K; c++) {
      if ((c & 31) == 0) region_base[c >> 5] = acc;
      acc += P.ncand[c];
      walk_nsamp[c] = P.ncand[c] > 0 ? P.nsamp[c] : 0;
    }
    region_base[n_regions] = acc;
  }
  // ---- per query: statistics, gate keys, buffers
  if (ctx->slots.size() < (size_t)Q) ctx->slots.resize(Q);
  std::vector<NormSetup> S(Q);
  BatchGate* G = reinterpret_cast<BatchGate*>(st + o_gate);
  std::memset(G, 0, sizeof(BatchGate));
  G->n_q = Q;
  const bool nothing = P.V == 0;
  for (int q = 0; q < Q; q++) {
    outs[q].cnt_candidate = P.cnt_candidate;
    outs[q].n_verified = P.V;
    outs[q].s_total = P.S;
    query_stats(queries + (size_t)q * m, m, &S[q].meanQ, &S[q].stdQ);
    S[q].inv_alpha = 1.0 / alpha;
    S[q].n_regions = n_regions;
    S[q].degenerate = !(S[q].stdQ > 0.0) || !(S[q].stdQ < INFINITY);
    BatchSlot& sl = ctx->slots[q];
    if (nothing) continue;
    KVM_CUDA(ctx, sl.wl_off.ensure(sizeof(int32_t) * (size_t)P.V));
    KVM_CUDA(ctx, sl.wl_ex.ensure(sizeof(double) * (size_t)P.V));
    KVM_CUDA(ctx, sl.wl_ex2.ensure(sizeof(double) * (size_t)P.V));
    KVM_CUDA(ctx, sl.region_count.ensure(sizeof(int32_t) * (n_regions + 1)));
    KVM_CUDA(ctx, sl.tile_prefix.ensure(sizeof(int32_t) * (n_regions + 2)));
    KVM_CUDA(ctx, sl.counters.ensure(sizeof(unsigned long long) * kNumCounters));
    KVM_CUDA(ctx, cudaMemsetAsync(sl.counters.p, 0, sizeof(unsigned long long) * kNumCounters, ctx->stream));
    if (S[q].degenerate) {  // no window can pass: an empty key range
      G->mean_klo[q] = G->var_klo[q] = INT32_MAX;
      G->mean_kspan[q] = G->var_kspan[q] = 0;
    } else {
      gate_keys(m, S[q], alpha, beta, &G->mean_klo[q], &G->mean_kspan[q], &G->var_klo[q], &G->var_kspan[q]);
    }
    G->e_off[q] = sl.wl_off.as<int32_t>();
    G->e_ex[q] = sl.wl_ex.as<double>();
    G->e_ex2[q] = sl.wl_ex2.as<double>();
    G->region_count[q] = sl.region_count.as<int32_t>();
    G->tile_prefix[q] = sl.tile_prefix.as<int32_t>();
    G->totals[q] = sl.counters.as<unsigned long long>() + kCntTiles;
  }
  if (nothing) {
    for (int q = 0; q < Q; q++)
      if ((rc = fetch_answers(ctx, 0, &outs[q]))) return rc;
    return KVM_OK;
  }
  KVM_CUDA(ctx, cudaMemcpyAsync(ctx->arena.p, ctx->stage.p, bytes, cudaMemcpyHostToDevice, ctx->stream));
  ctx->h2d_bytes += (long long)bytes;
  const unsigned char* base = ctx->arena.as<unsigned char>();
  // ---- one statistics pass for the set
  if ((rc = zero_counters(ctx))) return rc;
  WalkParams W;
  std::memset(&W, 0, sizeof(W));
  W.T = ctx->series;
  W.cbegin = reinterpret_cast<const int32_t*>(base + o_cbegin);
  W.cnsamp = reinterpret_cast<const int32_t*>(base + o_nsamp);
  W.region_base = reinterpret_cast<const long long*>(base + o_rbase);
  W.K = K;
  W.m = m;
  W.first_global = (int32_t)ctx->first;
  W.dm = (double)m;
  W.idx_hi = (int)((ctx->count + kTailPad - 2) & ~int64_t(1));
  W.done = reinterpret_cast<unsigned int*>(ctx->counters.as<unsigned long long>() + kCntDone);
  W.batch = reinterpret_cast<const BatchGate*>(base + o_gate);
  KVM_CUDA(ctx, cudaEventRecord(ctx->evs[2], ctx->stream));
  if ((m % 2) == 0) cnsm_relay_kernel<kRelayStages, 1, 2><<<n_regions, kRelayThreads, relay_smem_bytes(kRelayStages), ctx->stream>>>(W);
  else cnsm_relay_kernel<kRelayStages, 0, 2><<<n_regions, kRelayThreads, relay_smem_bytes(kRelayStages), ctx->stream>>>(W);
  KVM_CUDA(ctx, cudaEventRecord(ctx->evs[3], ctx->stream));
  KVM_CUDA(ctx, cudaGetLastError());
  float walk_ms = 0.f;
  // ---- while the statistics pass runs: every query's z-normalisation and |z| ordering (K/NormQueryEngine.java:438-448)
  // into one pinned block [q][zq f64 x m | order i32 x m], one copy per query into its own device buffer
  const size_t q_zq = 0, q_order = up256(sizeof(double) * (size_t)m), q_bytes = up256(q_order + sizeof(int32_t) * (size_t)m);
  const size_t o_zq = q_zq, o_order = q_order;
  KVM_CUDA(ctx, ctx->stage2.ensure(q_bytes * (size_t)Q + 256));
  {
    std::vector<double> z(m);
    for (int q = 0; q < Q; q++) {
      if (S[q].degenerate) continue;
      BatchSlot& sl = ctx->slots[q];
      KVM_CUDA(ctx, sl.qarena.ensure(q_bytes + 256));
      unsigned char* dst = static_cast<unsigned char*>(ctx->stage2.p) + q_bytes * (size_t)q;
      double* zq = reinterpret_cast<double*>(dst + q_zq);
      int32_t* order = reinterpret_cast<int32_t*>(dst + q_order);
      const double* qq = queries + (size_t)q * m;
      for (int i = 0; i < m; i++) z[i] = (qq[i] - S[q].meanQ) / S[q].stdQ;
      for (int i = 0; i < m; i++) order[i] = i;
      std::stable_sort(order, order + m, [&](int32_t a, int32_t b) {
        return java_double_compare(std::fabs(z[b]), std::fabs(z[a])) < 0;
      });
      for (int i = 0; i < m; i++) zq[i] = z[order[i]];
      KVM_CUDA(ctx, cudaMemcpyAsync(sl.qarena.p, dst, q_bytes, cudaMemcpyHostToDevice, ctx->stream));
      ctx->h2d_bytes += (long long)q_bytes;
    }
  }
  // ---- per query: the single-query evaluator + exact stages on that query's buffers.  All launches and result
  // copies are enqueued first, then ONE synchronisation; a query whose candidate / answer buffers overflowed is re-run.
  const double eps2 = epsilon * epsilon;
  auto launch_tail = [&](int q) -> int {  // ctx holds slot q's buffers
    const unsigned char* qbase = ctx->qarena.as<unsigned char>();
    EvalParams E;
    E.T = ctx->series;
    E.first_global = (int32_t)ctx->first;
    E.m = m;
    E.e_off = ctx->wl_off.as<int32_t>();
    E.e_ex = ctx->wl_ex.as<double>();
    E.e_ex2 = ctx->wl_ex2.as<double>();
    E.region_base = reinterpret_cast<const long long*>(base + o_rbase);
    E.region_count = ctx->region_count.as<int32_t>();
    E.tile_prefix = ctx->tile_prefix.as<int32_t>();
    E.totals = ctx->counters.as<unsigned long long>() + kCntTiles;
    E.n_regions = n_regions;
    E.zq = reinterpret_cast<const double*>(qbase + o_zq);
    E.order = reinterpret_cast<const int32_t*>(qbase + o_order);
    E.meanQ = S[q].meanQ;
    E.stdQ = S[q].stdQ;
    E.alpha = alpha;
    E.inv_alpha = S[q].inv_alpha;
    E.beta = beta;
    E.eps2 = eps2;
    E.eps2_hi = eps2 * (1.0 + 1e-9) + 1e-18;
    E.out = cands_of(ctx);
    E.gate_pass = ctx->counters.as<unsigned long long>() + kCntGate;
    cnsm_ed_eval_kernel<<<ctx->n_sms * 16, kEvalTile, 0, ctx->stream>>>(E);
    ExactEdParams X;
    X.T = E.T;
    X.first_global = E.first_global;
    X.m = m;
    X.zq = E.zq;
    X.order = E.order;
    X.eps2 = eps2;
    X.eps2_hi = E.eps2_hi;
    X.n_exact = ctx->counters.as<unsigned long long>() + kCntFlag;
    X.in = E.out;
    X.sink = sink_of(ctx);
    X.win_cap = 0;
    X.q_cap = 0;
    cnsm_ed_exact_kernel<false><<<ctx->n_sms * 6, 128, sizeof(double) * kExactChunk * 4, ctx->stream>>>(X);
    KVM_CUDA(ctx, cudaGetLastError());
    outs[q].n_launches += 2;
    return enqueue_counter_read(ctx);
  };
  // the per-query stages are latency-bound kernels that leave the GPU half empty: spread them over a few streams
  constexpr int kAux = 3;
  for (int i = 0; i < kAux; i++)
    if (!ctx->aux[i]) KVM_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->aux[i], cudaStreamNonBlocking));
  if (!ctx->ev_set) KVM_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_set, cudaEventDisableTiming));
  KVM_CUDA(ctx, cudaEventRecord(ctx->ev_set, ctx->stream));  // statistics pass + every query's upload
  for (int i = 0; i < kAux; i++) KVM_CUDA(ctx, cudaStreamWaitEvent(ctx->aux[i], ctx->ev_set, 0));
  cudaStream_t main_stream = ctx->stream;
  for (int q = 0; q < Q; q++) {
    BatchSlot& sl = ctx->slots[q];
    outs[q].n_launches = (q == 0) ? 1 : 0;
    if (S[q].degenerate) {
      sl.off.clear();
      sl.dist.clear();
      continue;
    }
    swap_slot(ctx, sl);
    rc = ensure_answers(ctx, std::max<long long>(ctx->ans_cap, 1 << 16));
    if (rc == KVM_OK) rc = ensure_cands(ctx, std::max<long long>(ctx->cand_cap, 1 << 18));
    ctx->stream = ctx->aux[q % kAux];
    if (rc == KVM_OK) rc = launch_tail(q);
    ctx->stream = main_stream;
    swap_slot(ctx, sl);
    if (rc) return rc;
  }
  for (int i = 0; i < kAux; i++) {  // join
    KVM_CUDA(ctx, cudaEventRecord(ctx->ev_set, ctx->aux[i]));
    KVM_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_set, 0));
  }
  KVM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  float tail_ms = 0.f;
  cudaEventElapsedTime(&tail_ms, ctx->evs[3], ctx->ev1);
  int n_live = 0;
  for (int q = 0; q < Q; q++) n_live += S[q].degenerate ? 0 : 1;
  for (int q = 0; q < Q; q++) {
    if (S[q].degenerate) continue;
    BatchSlot& sl = ctx->slots[q];
    kvm_result* out = &outs[q];
    swap_slot(ctx, sl);
    unsigned long long cnt[kNumCounters];
    std::memcpy(cnt, ctx->h_counters.p, sizeof(cnt));
    rc = KVM_OK;
    for (int attempt = 1; attempt < 8; attempt++) {
      const bool cand_over = (long long)cnt[kCntCand] > ctx->cand_cap, ans_over = (long long)cnt[kCntAnswers] > ctx->ans_cap;
      if (!cand_over && !ans_over) break;
      if (cand_over) rc = ensure_cands(ctx, (long long)cnt[kCntCand] + 1024);
      if (rc == KVM_OK && ans_over) rc = ensure_answers(ctx, (long long)cnt[kCntAnswers] + 1024);
      if (rc) break;
      // re-run this query's tail only: keep the walker's totals, clear the tail's counters
      unsigned long long* c = ctx->counters.as<unsigned long long>();
      cudaMemsetAsync(c + kCntAnswers, 0, sizeof(unsigned long long) * 3, ctx->stream);  // answers, cand, gate
      cudaMemsetAsync(c + kCntFlag, 0, sizeof(unsigned long long), ctx->stream);
      if ((rc = launch_tail(q))) break;
      cudaStreamSynchronize(ctx->stream);
      std::memcpy(cnt, ctx->h_counters.p, sizeof(cnt));
      if (attempt == 7 && ((long long)cnt[kCntCand] > ctx->cand_cap || (long long)cnt[kCntAnswers] > ctx->ans_cap))
        rc = fail(ctx, KVM_E_OOM, "candidate/answer buffers kept overflowing");
    }
    if (rc == KVM_OK) {
      out->n_gate_pass = (int64_t)cnt[kCntGate];
      out->n_exact = (int64_t)cnt[kCntFlag];
      out->stage_ms[1] = n_live ? (double)tail_ms / n_live : 0.0;  // evaluator + exact, averaged over the set
      rc = fetch_answers(ctx, (long long)cnt[kCntAnswers], out);
    }
    if (rc == KVM_OK) {  // fetch_answers points into ctx-owned vectors that the next query overwrites
      sl.off.assign(out->offsets, out->offsets + out->count);
      sl.dist.assign(out->distances, out->distances + out->count);
    }
    swap_slot(ctx, sl);
    if (rc) return rc;
  }
  KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaEventElapsedTime(&walk_ms, ctx->evs[2], ctx->evs[3]);
  for (int q = 0; q < Q; q++) {
    BatchSlot& sl = ctx->slots[q];
    outs[q].offsets = sl.off.data();
    outs[q].distances = sl.dist.data();
    outs[q].count = (int64_t)sl.off.size();
    outs[q].h2d_bytes = (int32_t)std::min<long long>(ctx->h2d_bytes, INT32_MAX);
    outs[q].stage_ms[0] = (double)walk_ms / Q;  // the set's one statistics pass, shared
    outs[q].kernel_ms = outs[q].stage_ms[0] + outs[q].stage_ms[1] + outs[q].stage_ms[2];
  }
  return KVM_OK;
}

int kvm_verify_cnsm_dtw(kvm_ctx* ctx, const double* q, int32_t m, double epsilon, int32_t rho, double alpha, double beta,
                        const int32_t* lr, int32_t K, int32_t shift, kvm_result* out) {
  return verify_norm(ctx, Mode::kDtw, q, m, epsilon, rho, alpha, beta, lr, K, shift, out);
}

// Samples the reference's block iterator feeds to a scan executor: 125-sample nodes, the last one zero padded, and only
// within-node advances counted against n (nextData() of K/experiments/ucr/UcrDtwQueryExecutor.java and
// UcrEdQueryExecutor.java:58-81, as K/IndexBuilder.java:152-180).  For n % 125 != 0 the executors therefore also
// verify windows that run into the zero padding; the series buffer keeps that many zeros behind the data (kTailPad).
static int64_t ucr_samples_fed(int64_t n) {
  int64_t fed = 0, cnt = 0;
  const int64_t n_nodes = (n + 124) / 125;
  for (int64_t b = 0; b < n_nodes; b++) {
    const int64_t within = std::min<int64_t>(124, std::max<int64_t>(0, n - cnt));
    fed += 1 + within;
    cnt += within;
    if (within < 124) break;
  }
  return fed;
}

// One verify_norm call over `lr` with the phantom zeros counted as samples.
static int ucr_scan(kvm_ctx* ctx, Mode mode, const double* q, int32_t m, double epsilon, int32_t rho, double alpha, double beta,
                    int64_t fed, const std::vector<int32_t>& lr, kvm_result* out) {
  if (!out) return fail(ctx, KVM_E_ARG, "null/invalid argument");
  if (lr.empty()) {  // series shorter than the query: nothing to scan
    std::memset(out, 0, sizeof(*out));
    return KVM_OK;
  }
  const int64_t n_real = ctx->n, count_real = ctx->count;
  ctx->n = fed;
  ctx->count = fed;
  ctx->splan.valid = false;
  ctx->norm_cache.valid = false;
  const int rc = verify_norm(ctx, mode, q, m, epsilon, rho, alpha, beta, lr.data(), (int)(lr.size() / 2), 0, out);
  ctx->n = n_real;
  ctx->count = count_real;
  ctx->splan.valid = false;
  ctx->norm_cache.valid = false;
  return rc;
}

int kvm_scan_ucr_dtw(kvm_ctx* ctx, const double* q, int32_t m, double epsilon, int32_t rho, double alpha, double beta,
                     kvm_result* out) {
  if (!ctx) return KVM_E_ARG;
  if (!ctx->series) return fail(ctx, KVM_E_STATE, "no series loaded");
  if (ctx->first != 1 || ctx->count != ctx->n) return fail(ctx, KVM_E_STATE, "the UCR scan needs the whole series on this ctx");
  constexpr int64_t kEpoch = 100000;  // UcrDtwQueryExecutor.java:97
  if (m < 3 || m > kEpoch) return fail(ctx, KVM_E_ARG, "UCR-DTW needs 3 <= m <= EPOCH");
  const int64_t fed = ucr_samples_fed(ctx->n);
  const int64_t per = kEpoch - m + 1, last = fed - m + 1;  // window starts per buffer; last 1-based start
  std::vector<int32_t> lr;
  for (int64_t left = 1; left <= last; left += per) {
    lr.push_back((int32_t)left);
    lr.push_back((int32_t)std::min(left + per - 1, last));
  }
  const int rc = ucr_scan(ctx, Mode::kDtw, q, m, epsilon, rho, alpha, beta, fed, lr, out);
  if (rc) return rc;
  for (int32_t& o : ctx->res_off) o -= 1;  // 0-based offsets, :278
  return KVM_OK;
}

int kvm_scan_ucr_ed(kvm_ctx* ctx, const double* q, int32_t m, double epsilon, double alpha, double beta, kvm_result* out) {
  if (!ctx) return KVM_E_ARG;
  if (!ctx->series) return fail(ctx, KVM_E_STATE, "no series loaded");
  if (ctx->first != 1 || ctx->count != ctx->n) return fail(ctx, KVM_E_STATE, "the UCR scan needs the whole series on this ctx");
  if (m < 1) return fail(ctx, KVM_E_ARG, "UCR-ED needs m >= 1");
  // UcrEdQueryExecutor.java:138-176: ONE statistics chain from the first sample to the last (ex / ex2 are never
  // reset), 1-based offsets (:166) — i.e. the cNSM-ED engine's loop over the single interval [1, fed-m+1].
  const int64_t fed = ucr_samples_fed(ctx->n);
  std::vector<int32_t> lr;
  if (fed - m + 1 >= 1) {
    lr.push_back(1);
    lr.push_back((int32_t)(fed - m + 1));
  }
  return ucr_scan(ctx, Mode::kEd, q, m, epsilon, 0, alpha, beta, fed, lr, out);
}

int kvm_verify_dtw(kvm_ctx* ctx, const double* q, int32_t m, double epsilon, int32_t rho, const int32_t* lr, int32_t K,
                   int32_t shift, kvm_result* out) {
  int rc = check_common(ctx, q, m, epsilon, lr, K, out);
  if (rc) return rc;
  if (rho < 0 || m < 3) return fail(ctx, KVM_E_ARG, "DTW needs rho >= 0 and m >= 3");
  if ((rc = begin_call(ctx))) return rc;
  Plan& P = ctx->plan_scratch;
  if ((rc = make_plan(ctx, lr, K, shift, m, &P))) return rc;
  out->cnt_candidate = P.cnt_candidate;
  out->n_verified = P.V;
  out->s_total = P.S;
  if (P.V == 0) return fetch_answers(ctx, 0, out);
  K = coalesce_runs(P.cbegin.data(), P.ncand.data(), K);  // (the data envelope is the shard's: interval borders do not matter)
  P.cbegin.resize(K);
  P.ncand.resize(K);
  std::vector<double> qv(q, q + m), uq, lq;
  envelope(qv, rho, lq, uq);  // K/QueryEngineDtw.java:362
  int64_t n_tiles = 0;
  std::vector<int32_t>& tp = ctx->tp_scratch;
  tile_prefix_of(P, kEdTile, &n_tiles, tp);
  Arena& A = ctx->arena_scratch;
  A.host.clear();
  const size_t o_q = A.add(q, sizeof(double) * m);
  const size_t o_uq = A.add(uq.data(), sizeof(double) * m);
  const size_t o_lq = A.add(lq.data(), sizeof(double) * m);
  const size_t o_cbegin = A.add(P.cbegin.data(), sizeof(int32_t) * K);
  const size_t o_ncand = A.add(P.ncand.data(), sizeof(int32_t) * K);
  const size_t o_tp = A.add(tp.data(), sizeof(int32_t) * (K + 1));
  if ((rc = upload_arena(ctx, A))) return rc;
  if ((rc = ensure_answers(ctx, std::max<long long>(ctx->ans_cap, 1 << 16)))) return rc;
  if ((rc = ensure_cands(ctx, std::max<long long>(ctx->cand_cap, 1 << 18)))) return rc;
  const unsigned char* base = ctx->arena.as<unsigned char>();
  const double eps2 = epsilon * epsilon;
  unsigned long long cnt[kNumCounters];
  for (int attempt = 0; attempt < 8; attempt++) {
    if ((rc = zero_counters(ctx))) return rc;
    LbRawParams L;
    L.T = ctx->series;
    L.cbegin = reinterpret_cast<const int32_t*>(base + o_cbegin);
    L.ncand = reinterpret_cast<const int32_t*>(base + o_ncand);
    L.tile_prefix = reinterpret_cast<const int32_t*>(base + o_tp);
    L.K = K;
    L.first_global = (int32_t)ctx->first;
    L.Q = LbQuery{reinterpret_cast<const double*>(base + o_q), reinterpret_cast<const double*>(base + o_uq),
                  reinterpret_cast<const double*>(base + o_lq), m, eps2 * (1.0 + 1e-9) + 1e-18};
    L.out = cands_of(ctx);
    DtwParams D;
    D.T = L.T;
    D.first_global = L.first_global;
    D.m = m;
    D.rho = rho;
    D.q = L.Q.q;
    D.uq = L.Q.uq;
    D.lq = L.Q.lq;
    D.eps2 = eps2;
    D.eps2_hi = L.Q.eps2_hi;
    D.n_abandoned = ctx->counters.as<unsigned long long>() + kCntFlag;
    D.n_cells = ctx->counters.as<unsigned long long>() + kCntCells;
    D.in = L.out;
    D.sink = sink_of(ctx);
    KVM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    dtw_lb_raw_kernel<<<(unsigned)n_tiles, kEdThreads, 0, ctx->stream>>>(L);
    KVM_CUDA(ctx, cudaEventRecord(ctx->evs[0], ctx->stream));
    if ((rc = launch_lb_data(ctx, L.Q.q, L.Q.uq, L.Q.lq, m, rho, L.Q.eps2_hi))) return rc;  // LB_Keogh on the data envelope
    D.in = cands2_of(ctx);
    KVM_CUDA(ctx, cudaEventRecord(ctx->evs[1], ctx->stream));
    if ((rc = launch_dtw(ctx, D))) return rc;
    KVM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    KVM_CUDA(ctx, cudaGetLastError());
    if ((rc = read_counters(ctx, cnt))) return rc;
    out->kernel_ms += elapsed_ms(ctx);
    add_stage_ms(ctx, out);
    out->n_launches += 3;
    const bool cand_over = (long long)cnt[kCntCand] > ctx->cand_cap, ans_over = (long long)cnt[kCntAnswers] > ctx->ans_cap;
    if (!cand_over && !ans_over) break;
    if (attempt == 7) return fail(ctx, KVM_E_OOM, "result buffers kept overflowing");
    if (cand_over && (rc = ensure_cands(ctx, (long long)cnt[kCntCand] + 1024))) return rc;
    if (ans_over && (rc = ensure_answers(ctx, (long long)cnt[kCntAnswers] + 1024))) return rc;
  }
  out->n_lb_pass = (int64_t)cnt[kCntCand2];
  out->n_dtw_cells = (int64_t)cnt[kCntCells];
  return fetch_answers(ctx, (long long)cnt[kCntAnswers], out);
}

int kvm_window_mean_runs(kvm_ctx* ctx, int32_t w, kvm_runs* out) {
  if (!ctx) return KVM_E_ARG;
  if (!out) return fail(ctx, KVM_E_ARG, "out is null");
  std::memset(out, 0, sizeof(*out));
  constexpr int kEpoch = 100000;  // K/IndexBuilder.java:136
  if (w < 2 || w > kEpoch) return fail(ctx, KVM_E_ARG, "window width %d", w);
  if (!ctx->series) return fail(ctx, KVM_E_STATE, "no series loaded");
  if (ctx->first != 1 || ctx->count != ctx->n)
    return fail(ctx, KVM_E_RANGE, "index build needs the whole series on this ctx");
  int rc = begin_call(ctx);
  if (rc) return rc;
  const int64_t n = ctx->n;
  // Samples the reference's block iterator feeds (K/IndexBuilder.java:152-180): 125-sample nodes,
  // the last one zero padded; nextData() counts only within-node advances against n.
  int64_t fed = 0;
  {
    const int64_t n_nodes = (n + 124) / 125;
    int64_t cnt = 0;
    for (int64_t b = 0; b < n_nodes; b++) {
      const int64_t within = std::min<int64_t>(124, std::max<int64_t>(0, n - cnt));
      fed += 1 + within;
      cnt += within;
      if (within < 124) break;  // ++cnt <= n failed inside this node
    }
  }
  // chains = the reference's epochs: epoch `it` starts at sample it*(EPOCH-w+1) and restarts the running sum
  std::vector<int32_t> cbegin, cnsamp;
  const int64_t stride = kEpoch - w + 1;
  int64_t n_win = 0;
  for (int64_t it = 0;; it++) {
    const int64_t g0 = it * stride;
    if (g0 + w - 1 >= fed) break;  // ep <= w-1: nothing new to read
    const int64_t ep = std::min<int64_t>(kEpoch, fed - g0);
    const int64_t nwin = std::min<int64_t>(ep - w + 1, n - g0);  // loc = g0 + i - w + 2 <= n
    if (nwin <= 0) break;
    cbegin.push_back((int32_t)g0);
    cnsamp.push_back((int32_t)(w - 1 + nwin));  // samples the chain needs for its nwin windows
    n_win = g0 + nwin;                            // windows are contiguous: loc 1 .. n_win
    if (ep < kEpoch) break;
  }
  const int n_chains = (int)cbegin.size();
  if (n_chains == 0) return KVM_OK;
  const int n_regions = (n_chains + 31) / 32;
  std::vector<long long> region_base(n_regions + 1, 0);
  Arena A;
  const size_t o_cbegin = A.add(cbegin.data(), sizeof(int32_t) * n_chains);
  const size_t o_nsamp = A.add(cnsamp.data(), sizeof(int32_t) * n_chains);
  const size_t o_rbase = A.add(region_base.data(), sizeof(long long) * (n_regions + 1));
  if ((rc = upload_arena(ctx, A))) return rc;
  const int n_tiles = (int)((n_win + kRleTile - 1) / kRleTile);
  KVM_CUDA(ctx, ctx->seg_b.ensure(sizeof(int32_t) * (size_t)n_win));
  KVM_CUDA(ctx, ctx->chain_count.ensure(sizeof(int32_t) * (size_t)n_tiles));
  KVM_CUDA(ctx, ctx->chain_prefix.ensure(sizeof(long long) * (size_t)(n_tiles + 1)));
  KVM_CUDA(ctx, ctx->region_count.ensure(sizeof(int32_t) * (n_regions + 1)));
  if ((rc = zero_counters(ctx))) return rc;
  const unsigned char* base = ctx->arena.as<unsigned char>();
  WalkParams W{};
  W.T = ctx->series;
  W.cbegin = reinterpret_cast<const int32_t*>(base + o_cbegin);
  W.cnsamp = reinterpret_cast<const int32_t*>(base + o_nsamp);
  W.region_base = reinterpret_cast<const long long*>(base + o_rbase);
  W.K = n_chains;
  W.m = w;
  W.first_global = 1;
  W.idx_hi = (int)((ctx->count + kTailPad - 2) & ~int64_t(1));
  W.dm = (double)w;
  W.region_count = ctx->region_count.as<int32_t>();
  W.tile_prefix = nullptr;
  W.totals = nullptr;
  W.done = nullptr;
  W.batch = nullptr;
  W.bucket_out = ctx->seg_b.as<int32_t>();
  W.c20w = 20.0 / (double)w;
  W.overflow = reinterpret_cast<int*>(ctx->counters.as<unsigned long long>() + kCntFlag);
  KVM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
  if ((w % 2) == 0) cnsm_relay_kernel<kRelayStages, 1, 1><<<n_regions, kRelayThreads, relay_smem_bytes(kRelayStages), ctx->stream>>>(W);
  else cnsm_relay_kernel<kRelayStages, 0, 1><<<n_regions, kRelayThreads, relay_smem_bytes(kRelayStages), ctx->stream>>>(W);
  rle_count_kernel<<<n_tiles, 256, 0, ctx->stream>>>(W.bucket_out, (long long)n_win, ctx->chain_count.as<int32_t>());
  rle_scan_kernel<<<1, 1024, 0, ctx->stream>>>(ctx->chain_count.as<int32_t>(), n_tiles, ctx->chain_prefix.as<long long>());
  KVM_CUDA(ctx, cudaGetLastError());
  long long total = 0;
  KVM_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters.p, ctx->chain_prefix.as<long long>() + n_tiles, sizeof(long long),
                                cudaMemcpyDeviceToHost, ctx->stream));
  KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  std::memcpy(&total, ctx->h_counters.p, sizeof(long long));
  KVM_CUDA(ctx, ctx->run_key.ensure(sizeof(double) * (size_t)total));
  KVM_CUDA(ctx, ctx->run_first.ensure(sizeof(int32_t) * (size_t)total));
  rle_emit_kernel<<<n_tiles, 256, 0, ctx->stream>>>(W.bucket_out, (long long)n_win, ctx->chain_prefix.as<long long>(),
                                                   ctx->run_first.as<int32_t>(), ctx->run_key.as<double>());
  KVM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
  KVM_CUDA(ctx, cudaGetLastError());
  unsigned long long cnt[kNumCounters];
  if ((rc = read_counters(ctx, cnt))) return rc;
  if (cnt[kCntFlag]) return fail(ctx, KVM_E_ARG, "window mean outside the supported key range (|mean| < 1e8)");
  out->kernel_ms = elapsed_ms(ctx);
  out->n_launches = 4;
  KVM_CUDA(ctx, ctx->h_key.ensure(sizeof(double) * (size_t)total));
  KVM_CUDA(ctx, ctx->h_first.ensure(sizeof(int32_t) * (size_t)total));
  KVM_CUDA(ctx, cudaMemcpyAsync(ctx->h_key.p, ctx->run_key.p, sizeof(double) * total, cudaMemcpyDeviceToHost, ctx->stream));
  KVM_CUDA(ctx, cudaMemcpyAsync(ctx->h_first.p, ctx->run_first.p, sizeof(int32_t) * total, cudaMemcpyDeviceToHost, ctx->stream));
  KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  // split at 255 positions (K/IndexBuilder.java:268, MAXIMUM_DIFF - 1)
  const double* hk = ctx->h_key.as<double>();
  const int32_t* hs = ctx->h_first.as<int32_t>();
  ctx->run_key_v.clear();
  ctx->run_first_v.clear();
  ctx->run_last_v.clear();
  ctx->run_key_v.reserve((size_t)total + 16);
  ctx->run_first_v.reserve((size_t)total + 16);
  ctx->run_last_v.reserve((size_t)total + 16);
  for (long long i = 0; i < total; i++) {
    const int64_t first = (int64_t)hs[i] + 1;                                   // 1-based loc
    const int64_t last = (i + 1 < total) ? (int64_t)hs[i + 1] : (int64_t)n_win;  // next run's first - 1
    for (int64_t f = first; f <= last; f += 255) {
      ctx->run_key_v.push_back(hk[i]);
      ctx->run_first_v.push_back((int32_t)f);
      ctx->run_last_v.push_back((int32_t)std::min<int64_t>(f + 254, last));
    }
  }
  out->count = (int64_t)ctx->run_key_v.size();
  out->keys = ctx->run_key_v.data();
  out->first = ctx->run_first_v.data();
  out->last = ctx->run_last_v.data();
  return KVM_OK;
}

// Window positions and epoch chains of one width, as the reference's block iterator produces them
// (K/IndexBuilder.java:152-180,194-301): `fed` = samples it feeds (125-sample nodes, the last one zero padded).
static void wmean_geometry(int64_t n, int w, int64_t* n_win_out, int64_t* n_chains_out) {
  constexpr int64_t kEpoch = 100000;
  int64_t fed = 0;
  {
    const int64_t n_nodes = (n + 124) / 125;
    int64_t cnt = 0;
    for (int64_t b = 0; b < n_nodes; b++) {
      const int64_t within = std::min<int64_t>(124, std::max<int64_t>(0, n - cnt));
      fed += 1 + within;
      cnt += within;
      if (within < 124) break;
    }
  }
  const int64_t stride = kEpoch - w + 1;
  int64_t n_win = 0, n_chains = 0;
  for (int64_t it = 0;; it++) {
    const int64_t g0 = it * stride;
    if (g0 + w - 1 >= fed) break;
    const int64_t ep = std::min<int64_t>(kEpoch, fed - g0);
    const int64_t nwin = std::min<int64_t>(ep - w + 1, n - g0);
    if (nwin <= 0) break;
    n_win = g0 + nwin;
    n_chains = it + 1;
    if (ep < kEpoch) break;
  }
  *n_win_out = n_win;
  *n_chains_out = n_chains;
}

// All window widths of one index build in ONE pass over the series (wmean_kernels.cuh).  outs[q] is what
// kvm_window_mean_runs(ctx, widths[q], ...) returns; its arrays stay valid until the next window-mean call on this ctx.
int kvm_window_mean_runs_all(kvm_ctx* ctx, const int32_t* widths, int32_t n_widths, kvm_runs* outs) {
  if (!ctx) return KVM_E_ARG;
  if (!widths || !outs || n_widths < 1 || n_widths > kvm::kMaxWidths)
    return fail(ctx, KVM_E_ARG, "1..%d window widths per call", kvm::kMaxWidths);
  for (int q = 0; q < n_widths; q++) std::memset(&outs[q], 0, sizeof(kvm_runs));
  constexpr int kEpoch = 100000;
  int w_max = 0;
  for (int q = 0; q < n_widths; q++) {
    if (widths[q] < 2 || widths[q] > kEpoch) return fail(ctx, KVM_E_ARG, "window width %d", widths[q]);
    w_max = std::max(w_max, widths[q]);
  }
  if (!ctx->series) return fail(ctx, KVM_E_STATE, "no series loaded");
  if (ctx->first != 1 || ctx->count != ctx->n) return fail(ctx, KVM_E_RANGE, "index build needs the whole series on this ctx");
  constexpr int NT = 256;
  const size_t smem = kvm::wmean_smem_bytes(NT, w_max);
  auto fallback = [&]() -> int {  // width by width through the exact walker
    for (int q = 0; q < n_widths; q++) {
      kvm_runs r;
      int rc = kvm_window_mean_runs(ctx, widths[q], &r);
      if (rc) return rc;
      kvm_ctx::WmeanSlot& S = ctx->wm[q];
      S.keys.assign(r.keys, r.keys + r.count);
      S.first.assign(r.first, r.first + r.count);
      S.last.assign(r.last, r.last + r.count);
      outs[q] = r;
      outs[q].keys = S.keys.data();
      outs[q].first = S.first.data();
      outs[q].last = S.last.data();
    }
    return KVM_OK;
  };
  if (smem > 110 * 1024 || !std::isfinite(ctx->absmax)) return fallback();
  int rc = begin_call(ctx);
  if (rc) return rc;
  const int64_t n = ctx->n;
  int64_t n_win[kvm::kMaxWidths] = {0}, n_chains[kvm::kMaxWidths] = {0};
  int64_t total_pos = 0;
  for (int q = 0; q < n_widths; q++) {
    wmean_geometry(n, widths[q], &n_win[q], &n_chains[q]);
    total_pos = std::max(total_pos, n_win[q]);
  }
  if (total_pos == 0) return KVM_OK;
  const int W = kvm::kGroup * NT;
  const int n_tiles = (int)((total_pos + W - 1) / W);
  const size_t n_segs = (size_t)n_tiles * (NT / 32);
  // counters: per width [n_runs, n_flagged, x_count], then the overflow flag
  const size_t n_cnt = 3 * kvm::kMaxWidths + 1;
  KVM_CUDA(ctx, ctx->wm_counters.ensure(sizeof(unsigned long long) * n_cnt));
  KVM_CUDA(ctx, ctx->wm_host.ensure(sizeof(unsigned long long) * n_cnt));
  unsigned long long* cnt_d = ctx->wm_counters.as<unsigned long long>();
  static bool attr_set = false;
  if (!attr_set) {
    KVM_CUDA(ctx, cudaFuncSetAttribute(kvm::wmean_stream_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    attr_set = true;
  }
  unsigned long long cnt[n_cnt];
  double kernel_ms = 0;
  int launches = 0;
  for (int attempt = 0; attempt < 6; attempt++) {
    kvm::WmeanParams P{};
    P.T = ctx->series;
    P.n_widths = n_widths;
    P.w_max = w_max;
    P.total_pos = (int)total_pos;
    P.epoch = kEpoch;
    P.bmax = ctx->bmax.as<double>();
    P.n_bmax = (int)ctx->n_bmax;
    P.overflow = reinterpret_cast<int*>(cnt_d + 3 * kvm::kMaxWidths);
    for (int q = 0; q < n_widths; q++) {
      kvm_ctx::WmeanSlot& S = ctx->wm[q];
      const int w = widths[q];
      if (S.run_cap == 0) {
        S.run_cap = n_win[q] / 3 + n_win[q] / 33 + 4096;
        KVM_CUDA(ctx, S.runs.ensure(sizeof(int2) * (size_t)S.run_cap));
        S.run_cap = (long long)(S.runs.cap / sizeof(int2));
      }
      if (n_segs > S.segs_n) {
        KVM_CUDA(ctx, S.seg_off.ensure(sizeof(int) * n_segs));
        KVM_CUDA(ctx, S.seg_cnt.ensure(sizeof(int) * n_segs));
        S.segs_n = n_segs;
      }
      const size_t words = (size_t)(n_win[q] + 63) / 32 + 4;
      if (words > S.need_words) {
        KVM_CUDA(ctx, S.need_bits.ensure(sizeof(unsigned) * words));
        S.need_words = S.need_bits.cap / sizeof(unsigned);
        KVM_CUDA(ctx, cudaMemsetAsync(S.need_bits.p, 0, S.need_bits.cap, ctx->stream));
      }
      if ((size_t)n_chains[q] > S.chains_n) {
        KVM_CUDA(ctx, S.chain_last.ensure(sizeof(int32_t) * (size_t)n_chains[q]));
        KVM_CUDA(ctx, S.flagged.ensure(sizeof(int32_t) * (size_t)n_chains[q]));
        S.chains_n = std::min(S.chain_last.cap, S.flagged.cap) / sizeof(int32_t);
        KVM_CUDA(ctx, cudaMemsetAsync(S.chain_last.p, 0xff, S.chain_last.cap, ctx->stream));
      }
      if (S.x_cap == 0) {
        S.x_cap = 1 << 14;
        KVM_CUDA(ctx, S.x_off.ensure(sizeof(int32_t) * (size_t)S.x_cap));
        KVM_CUDA(ctx, S.x_ex.ensure(sizeof(double) * (size_t)S.x_cap));
        KVM_CUDA(ctx, S.x_ex2.ensure(sizeof(double) * (size_t)S.x_cap));
      }
      kvm::WmeanWidth& Wq = P.W[q];
      Wq.w = w;
      Wq.n_win = (int)n_win[q];
      Wq.c20w = 20.0 / (double)w;
      {
        // the chain does 2 roundings of size <= u*w*A per sample it consumes; the stream fewer than 320 on values
        // <= (tile samples)*A (as stream_guard()); both doubled
        const double u = kUlpHalf;
        Wq.cd_chain = 2.0 * (2.0 * u * (double)w * 1.01) * 1.000001;
        Wq.cd1 = 2.0 * (320.0 * u * ((double)W + w_max)) * 1.000001;
      }
      Wq.runs = S.runs.as<int2>();
      Wq.run_cap = S.run_cap;
      Wq.n_runs = cnt_d + 3 * q;
      Wq.seg_off = S.seg_off.as<int>();
      Wq.seg_cnt = S.seg_cnt.as<int>();
      Wq.need_bits = S.need_bits.as<unsigned>();
      Wq.chain_last = S.chain_last.as<int32_t>();
      Wq.flagged = S.flagged.as<int32_t>();
      Wq.n_flagged = cnt_d + 3 * q + 1;
    }
    KVM_CUDA(ctx, cudaMemsetAsync(cnt_d, 0, sizeof(unsigned long long) * n_cnt, ctx->stream));
    KVM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    kvm::wmean_stream_kernel<NT><<<n_tiles, NT, smem, ctx->stream>>>(P);
    KVM_CUDA(ctx, cudaGetLastError());
    launches += 1;
    {  // exact re-walk of the epochs that hold an ambiguous window: all widths in one launch
      kvm::RewalkBatch B{};
      int64_t most = 1;
      for (int q = 0; q < n_widths; q++) {
        kvm_ctx::WmeanSlot& S = ctx->wm[q];
        kvm::RewalkParams& R = B.set[q];
        R.T = ctx->series;
        R.chains.regular = 1;
        R.chains.s_base = 0;
        R.chains.chunk = kEpoch - widths[q] + 1;
        R.chains.total_win = (int32_t)n_win[q];
        R.chains.n_chains = (int)n_chains[q];
        R.m = widths[q];
        R.first_global = 1;
        R.need_bits = S.need_bits.as<unsigned>();
        R.chain_last = S.chain_last.as<int32_t>();
        R.flagged = S.flagged.as<int32_t>();
        R.n_flagged = cnt_d + 3 * q + 1;
        R.out = kvm::XList{S.x_off.as<int32_t>(), S.x_ex.as<double>(), S.x_ex2.as<double>(), cnt_d + 3 * q + 2, S.x_cap};
        most = std::max(most, n_chains[q]);
      }
      kvm::chain_rewalk_batch_kernel<<<(unsigned)(std::min<int64_t>(most, ctx->n_sms * 4) * n_widths), 32, 0, ctx->stream>>>(B, n_widths);
      KVM_CUDA(ctx, cudaGetLastError());
      launches += 1;
    }
    KVM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    KVM_CUDA(ctx, cudaMemcpyAsync(ctx->wm_host.p, cnt_d, sizeof(unsigned long long) * n_cnt, cudaMemcpyDeviceToHost, ctx->stream));
    KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::memcpy(cnt, ctx->wm_host.p, sizeof(unsigned long long) * n_cnt);
    kernel_ms += elapsed_ms(ctx);
    if (cnt[3 * kvm::kMaxWidths]) return fallback();  // bucket range / guard: the exact walker handles it
    bool again = false;
    for (int q = 0; q < n_widths; q++) {
      kvm_ctx::WmeanSlot& S = ctx->wm[q];
      if ((long long)cnt[3 * q] > S.run_cap) {
        S.run_cap = (long long)cnt[3 * q] + 4096;
        KVM_CUDA(ctx, S.runs.ensure(sizeof(int2) * (size_t)S.run_cap));
        again = true;
      }
      if ((long long)cnt[3 * q + 2] > S.x_cap) {
        S.x_cap = (long long)cnt[3 * q + 2] + 1024;
        KVM_CUDA(ctx, S.x_off.ensure(sizeof(int32_t) * (size_t)S.x_cap));
        KVM_CUDA(ctx, S.x_ex.ensure(sizeof(double) * (size_t)S.x_cap));
        KVM_CUDA(ctx, S.x_ex2.ensure(sizeof(double) * (size_t)S.x_cap));
        again = true;
      }
    }
    if (!again) break;
    if (attempt == 5) return fail(ctx, KVM_E_OOM, "run buffers kept overflowing");
  }
  // ---- host: stitch the warps' slices in position order, resolve the ambiguous windows with the reference's
  // arithmetic, merge equal neighbours, split at 255 positions (K/IndexBuilder.java:268, MAXIMUM_DIFF - 1).  The widths
  // are independent: each one is stitched on its own host thread while the next one's runs are still being copied.
  struct WmHost {
    std::vector<int2> runs;
    std::vector<int> seg_off, seg_cnt;
    std::vector<int32_t> x_off;
    std::vector<double> x_ex;
    int bad_pos = -1;  // an ambiguous window the re-walk did not deliver
  };
  std::vector<WmHost> host((size_t)n_widths);
  std::vector<std::thread> workers;
  struct Joiner {
    std::vector<std::thread>& w;
    ~Joiner() {
      for (std::thread& t : w)
        if (t.joinable()) t.join();
    }
  } joiner{workers};
  auto stitch = [&](int q) {
    kvm_ctx::WmeanSlot& S = ctx->wm[q];
    WmHost& H = host[(size_t)q];
    const int w = widths[q];
    const size_t n_x = H.x_off.size();
    // exact buckets of the ambiguous windows: b = floor(2 * fl(fl(ex / w) * 10)), K/utils/MeanIntervalUtils.java:51-61
    std::vector<std::pair<int32_t, int>> exact(n_x);
    for (size_t i = 0; i < n_x; i++) {
      const double v = (H.x_ex[i] / (double)w) * 10.0;
      exact[i] = {H.x_off[i] - 1, (int)std::floor(v + v)};  // position = loc - 1
    }
    std::sort(exact.begin(), exact.end());
    S.keys.clear();
    S.first.clear();
    S.last.clear();
    const size_t est = H.runs.size() + (size_t)n_win[q] / 255 + 16;
    S.keys.reserve(est);
    S.first.reserve(est);
    S.last.reserve(est);
    auto close_run = [&](int b, int64_t first_pos, int64_t last_pos) {
      const double int_value = std::floor((double)b * 0.5);
      const double ret = (b & 1) ? int_value + 0.5 : int_value;
      const double key = ret * 0.1;
      for (int64_t f = first_pos; f <= last_pos; f += 255) {
        S.keys.push_back(key);
        S.first.push_back((int32_t)(f + 1));
        S.last.push_back((int32_t)(std::min<int64_t>(f + 254, last_pos) + 1));
      }
    };
    bool open = false;
    int cur_b = 0;
    int64_t cur_first = 0;
    size_t xi = 0;
    for (size_t sg = 0; sg < n_segs; sg++) {
      const int c = H.seg_cnt[sg];
      if (c <= 0) continue;
      const int2* r = H.runs.data() + H.seg_off[sg];
      for (int i = 0; i < c; i++) {
        int b = r[i].y;
        const int pos = r[i].x;
        if (b == kvm::kAmbiguous) {
          while (xi < exact.size() && exact[xi].first < pos) xi++;
          if (xi >= exact.size() || exact[xi].first != pos) {
            H.bad_pos = pos;
            return;
          }
          b = exact[xi].second;
        }
        if (open && b == cur_b) continue;  // same key as the run before: one run
        if (open) close_run(cur_b, cur_first, (int64_t)pos - 1);
        open = true;
        cur_b = b;
        cur_first = pos;
      }
    }
    if (open) close_run(cur_b, cur_first, n_win[q] - 1);
    outs[q].count = (int64_t)S.keys.size();
    outs[q].keys = S.keys.data();
    outs[q].first = S.first.data();
    outs[q].last = S.last.data();
    outs[q].kernel_ms = kernel_ms;
    outs[q].n_launches = launches;
    outs[q].reserved = (int32_t)std::min<unsigned long long>(cnt[3 * q + 1], INT32_MAX);  // epochs re-walked exactly
  };
  for (int q = 0; q < n_widths; q++) {
    kvm_ctx::WmeanSlot& S = ctx->wm[q];
    WmHost& H = host[(size_t)q];
    const long long n_runs = (long long)cnt[3 * q], n_x = (long long)cnt[3 * q + 2];
    H.runs.resize((size_t)n_runs);
    H.seg_off.resize(n_segs);
    H.seg_cnt.resize(n_segs);
    H.x_off.resize((size_t)n_x);
    H.x_ex.resize((size_t)n_x);
    KVM_CUDA(ctx, cudaMemcpyAsync(H.runs.data(), S.runs.p, sizeof(int2) * (size_t)n_runs, cudaMemcpyDeviceToHost, ctx->stream));
    KVM_CUDA(ctx, cudaMemcpyAsync(H.seg_off.data(), S.seg_off.p, sizeof(int) * n_segs, cudaMemcpyDeviceToHost, ctx->stream));
    KVM_CUDA(ctx, cudaMemcpyAsync(H.seg_cnt.data(), S.seg_cnt.p, sizeof(int) * n_segs, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_x) {
      KVM_CUDA(ctx, cudaMemcpyAsync(H.x_off.data(), S.x_off.p, sizeof(int32_t) * (size_t)n_x, cudaMemcpyDeviceToHost, ctx->stream));
      KVM_CUDA(ctx, cudaMemcpyAsync(H.x_ex.data(), S.x_ex.p, sizeof(double) * (size_t)n_x, cudaMemcpyDeviceToHost, ctx->stream));
    }
    KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    workers.emplace_back(stitch, q);
  }
  for (std::thread& t : workers) t.join();
  for (int q = 0; q < n_widths; q++)
    if (host[(size_t)q].bad_pos >= 0)
      return fail(ctx, KVM_E_CUDA, "window-mean pass: ambiguous window %d of width %d was not re-walked", host[(size_t)q].bad_pos, widths[q]);
  return KVM_OK;
}

// DtwUtils.lowerUpperLemire on the device: the envelope of samples [first, first + len - 1] (1-based) of the loaded
// series, clamped at the ends of that region exactly as the reference clamps at the ends of its read buffer.
int kvm_envelope(kvm_ctx* ctx, int32_t r, int64_t first, int32_t len, double* lower, double* upper) {
  if (!ctx) return KVM_E_ARG;
  if (!lower || !upper || len < 1 || r < 0) return fail(ctx, KVM_E_ARG, "null/invalid argument");
  if (r > 512) return fail(ctx, KVM_E_ARG, "envelope radius %d exceeds the supported maximum 512", r);
  if (!ctx->series) return fail(ctx, KVM_E_STATE, "no series loaded");
  if (first < ctx->first || first + len - 1 > ctx->first + ctx->count - 1)
    return fail(ctx, KVM_E_RANGE, "samples [%lld,%lld] are not on this ctx", (long long)first, (long long)(first + len - 1));
  int rc = begin_call(ctx);
  if (rc) return rc;
  KVM_CUDA(ctx, ctx->run_key.ensure(sizeof(double) * 2 * (size_t)len));
  double* lo_d = ctx->run_key.as<double>();
  double* up_d = lo_d + len;
  const size_t smem = sizeof(long long) * 2 * (size_t)(kvm::kEnvTile + 2 * r);
  static bool attr = false;
  if (!attr) {
    KVM_CUDA(ctx, cudaFuncSetAttribute(kvm::envelope_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  kvm::envelope_kernel<<<(len + kvm::kEnvTile - 1) / kvm::kEnvTile, 256, smem, ctx->stream>>>(ctx->series + (first - ctx->first), len, r,
                                                                                            lo_d, up_d);
  KVM_CUDA(ctx, cudaGetLastError());
  KVM_CUDA(ctx, cudaMemcpyAsync(lower, lo_d, sizeof(double) * (size_t)len, cudaMemcpyDeviceToHost, ctx->stream));
  KVM_CUDA(ctx, cudaMemcpyAsync(upper, up_d, sizeof(double) * (size_t)len, cudaMemcpyDeviceToHost, ctx->stream));
  KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return KVM_OK;
}

// ---- phase-1 tail on the host (phase1.hpp): plain arrays in / out, no ctx, no GPU ----------------------------------
static void to_ivs(const int32_t* lr, const double* eps, int64_t k, std::vector<kvm_phase1::Iv>& v) {
  v.resize((size_t)k);
  for (int64_t i = 0; i < k; i++) v[i] = kvm_phase1::Iv{lr[2 * i], lr[2 * i + 1], eps ? eps[i] : 0.0};
}
static int from_ivs(const std::vector<kvm_phase1::Iv>& v, int32_t* lr_out, double* eps_out, int64_t cap, int64_t* k_out) {
  *k_out = (int64_t)v.size();
  if ((int64_t)v.size() > cap) return KVM_E_ARG;
  for (size_t i = 0; i < v.size(); i++) {
    lr_out[2 * i] = v[i].left;
    lr_out[2 * i + 1] = v[i].right;
    if (eps_out) eps_out[i] = v[i].eps;
  }
  return KVM_OK;
}

int kvm_intervals_sort_merge(const int32_t* lr, const double* eps, int64_t k, int32_t mode, int32_t* lr_out, double* eps_out,
                             int64_t cap, int64_t* k_out, int64_t* cnt_disjoint, int64_t* cnt_offsets) {
  if (k < 0 || (k > 0 && !lr) || !lr_out || !k_out || mode < 0 || mode > 2) return KVM_E_ARG;
  std::vector<kvm_phase1::Iv> v, out;
  to_ivs(lr, eps, k, v);
  kvm_phase1::sort_merge(v, mode, out, cnt_disjoint, cnt_offsets);
  return from_ivs(out, lr_out, eps_out, cap, k_out);
}

int kvm_intervals_intersect(const int32_t* cs_lr, const double* cs_eps, int64_t k1, const int32_t* csi_lr, const double* csi_eps,
                            int64_t k2, double eps2, int32_t delta_w, int32_t* lr_out, double* eps_out, int64_t cap, int64_t* k_out,
                            double* min_eps) {
  if (k1 < 0 || k2 < 0 || (k1 > 0 && (!cs_lr || !cs_eps)) || (k2 > 0 && (!csi_lr || !csi_eps)) || !lr_out || !eps_out || !k_out)
    return KVM_E_ARG;
  std::vector<kvm_phase1::Iv> a, b, out;
  to_ivs(cs_lr, cs_eps, k1, a);
  to_ivs(csi_lr, csi_eps, k2, b);
  const double me = kvm_phase1::intersect(a, b, eps2, delta_w, out);
  if (min_eps) *min_eps = me;
  return from_ivs(out, lr_out, eps_out, cap, k_out);
}

int kvm_intervals_first_segment(const int32_t* lr, const double* eps, int64_t k, int32_t order, int32_t w0, int32_t length, int32_t n,
                                int32_t delta_w, int32_t* lr_out, double* eps_out, int64_t cap, int64_t* k_out, double* min_eps) {
  if (k < 0 || (k > 0 && (!lr || !eps)) || !lr_out || !eps_out || !k_out) return KVM_E_ARG;
  std::vector<kvm_phase1::Iv> v, out;
  to_ivs(lr, eps, k, v);
  const double me = kvm_phase1::first_segment(v, order, w0, length, n, delta_w, out);
  if (min_eps) *min_eps = me;
  return from_ivs(out, lr_out, eps_out, cap, k_out);
}

// ---- the same tail for the cNSM engines (K/NormQueryEngine.java:313-397, 788-896 and NormQueryEngineDtw.java:326-426,
// 926-1046): kvm_norm_interval = kvm_phase1::NormIv
static_assert(sizeof(kvm_norm_interval) == sizeof(kvm_phase1::NormIv) && sizeof(kvm_norm_interval) == 48, "kvm_norm_interval layout");
int kvm_norm_intervals_sort_merge(const kvm_norm_interval* in, int64_t k, int32_t mode, kvm_norm_interval* out, int64_t cap,
                                  int64_t* k_out, int64_t* cnt_disjoint, int64_t* cnt_offsets) {
  if (k < 0 || (k > 0 && !in) || !out || !k_out || mode < 0 || mode > 2 || k > (int64_t)0xffffffffLL) return KVM_E_ARG;
  // same layout (static_assert above): the lists are read and written in place, no staging copies
  const kvm_phase1::NormIv* v = reinterpret_cast<const kvm_phase1::NormIv*>(in);
  int64_t n_out = 0;
  kvm_phase1::norm_sort_merge_core(
      v, (size_t)k, mode,
      [&](const kvm_phase1::NormIv& x) {
        if (n_out < cap) out[n_out] = kvm_norm_interval{x.left, x.right, x.ex, x.ex2, x.exu, x.ex2u, x.bp};
        n_out++;
      },
      cnt_disjoint, cnt_offsets);
  *k_out = n_out;
  return n_out > cap ? KVM_E_ARG : KVM_OK;
}

// (writes results straight into the caller's array; counts on past `cap`)
struct NormSink {
  kvm_norm_interval* out;
  int64_t cap, n = 0;
  void operator()(const kvm_phase1::NormIv& x) {
    if (n < cap) out[n] = kvm_norm_interval{x.left, x.right, x.ex, x.ex2, x.exu, x.ex2u, x.bp};
    n++;
  }
};

int kvm_norm_intervals_intersect(const kvm_norm_interval* cs, int64_t k1, const kvm_norm_interval* csi, int64_t k2, int32_t pre_length,
                                 int32_t w0, int32_t query_length, double mean_q, double std_q, double alpha, double beta,
                                 int32_t delta_w, int32_t dtw, kvm_norm_interval* out, int64_t cap, int64_t* k_out) {
  if (k1 < 0 || k2 < 0 || (k1 > 0 && !cs) || (k2 > 0 && !csi) || !out || !k_out || pre_length < 1 || w0 < 1 || query_length < 1)
    return KVM_E_ARG;
  NormSink sink{out, cap};
  kvm_phase1::norm_intersect_core(reinterpret_cast<const kvm_phase1::NormIv*>(cs), (size_t)k1,
                                  reinterpret_cast<const kvm_phase1::NormIv*>(csi), (size_t)k2, pre_length, w0, query_length, mean_q,
                                  std_q, alpha, beta, delta_w, dtw != 0, sink);
  *k_out = sink.n;
  return sink.n > cap ? KVM_E_ARG : KVM_OK;
}

int kvm_norm_intervals_first_segment(const kvm_norm_interval* in, int64_t k, int32_t order, int32_t w0, int32_t length, int32_t n,
                                     int32_t delta_w, kvm_norm_interval* out, int64_t cap, int64_t* k_out) {
  if (k < 0 || (k > 0 && !in) || !out || !k_out) return KVM_E_ARG;
  NormSink sink{out, cap};
  kvm_phase1::norm_first_segment_core(reinterpret_cast<const kvm_phase1::NormIv*>(in), (size_t)k, order, w0, length, n, delta_w, sink);
  *k_out = sink.n;
  return sink.n > cap ? KVM_E_ARG : KVM_OK;
}

int kvm_index_image_from_runs(const double* keys, const int32_t* first, const int32_t* last, int64_t n_runs,
                              unsigned char** image, kvm_index_info* info) {
  if (!keys || !first || !last || !image || !info || n_runs < 0) return KVM_E_ARG;
  std::memset(info, 0, sizeof(*info));
  *image = nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<unsigned char> file;
  kvm_index::ImageInfo ii;
  if (!kvm_index::build_image(keys, first, last, n_runs, file, &ii)) return KVM_E_RANGE;
  *image = static_cast<unsigned char*>(std::malloc(file.size() ? file.size() : 1));
  if (!*image) return KVM_E_OOM;
  std::memcpy(*image, file.data(), file.size());
  info->file_bytes = (int64_t)file.size();
  info->n_runs = n_runs;
  info->n_intervals = ii.intervals;
  info->n_offsets = ii.offsets;
  info->n_rows_step1 = ii.rows_step1;
  info->n_rows = ii.rows;
  info->host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return KVM_OK;
}

int kvm_index_row_positions(const unsigned char* row, int64_t row_bytes, int32_t* lr_out, int64_t cap, int64_t* k_out) {
  if (!row || row_bytes < 0 || !k_out || (cap > 0 && !lr_out) || cap < 0) return KVM_E_ARG;
  const int64_t k = kvm_index::parse_compact(row, row_bytes, lr_out, cap);
  *k_out = k < 0 ? 0 : k;
  if (k < 0) return KVM_E_RANGE;
  return k > cap ? KVM_E_ARG : KVM_OK;
}

void kvm_image_free(unsigned char* image) { std::free(image); }

int kvm_build_index_file(kvm_ctx* ctx, int32_t w, const char* path, kvm_index_info* info) {
  if (!ctx) return KVM_E_ARG;
  if (!info) return fail(ctx, KVM_E_ARG, "info is null");
  std::memset(info, 0, sizeof(*info));
  kvm_runs runs;
  int rc = kvm_window_mean_runs(ctx, w, &runs);
  if (rc) return rc;
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<unsigned char> file;
  kvm_index::ImageInfo ii;
  if (!kvm_index::build_image(runs.keys, runs.first, runs.last, runs.count, file, &ii))
    return fail(ctx, KVM_E_RANGE, "no window of width %d in this series (the reference throws)", w);
  if (path) {
    FILE* f = std::fopen(path, "wb");
    if (!f) return fail(ctx, KVM_E_IO, "cannot open %s for writing", path);
    const size_t wrote = std::fwrite(file.data(), 1, file.size(), f);
    if (std::fclose(f) != 0 || wrote != file.size()) return fail(ctx, KVM_E_IO, "short write to %s", path);
  }
  info->file_bytes = (int64_t)file.size();
  info->n_runs = runs.count;
  info->n_intervals = ii.intervals;
  info->n_offsets = ii.offsets;
  info->n_rows_step1 = ii.rows_step1;
  info->n_rows = ii.rows;
  info->kernel_ms = runs.kernel_ms;
  info->host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return KVM_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Several GPUs behind one handle (single process): the series is sharded by offset range, every device verifies the
// intervals whose first scanned sample it owns (chains are never split), one host thread per device drives its ctx,
// and the host concatenates the answers — shards own ascending offset ranges, so device order is offset order.
// No series data and no collective crosses GPUs; (the multi-process layout, one rank per GPU with an NCCL all_gather of
// the packed answers, is kvmatch_b200/sharding.py.)
struct kvm_multi {
  std::vector<kvm_ctx*> ctx;
  std::vector<int64_t> own_lo, own_hi;  // owned window starts (1-based, inclusive)
  int64_t n = 0;
  std::string err;
  std::vector<int32_t> off;
  std::vector<double> dist;
};

int kvm_multi_create(kvm_multi** out, const int32_t* device_ids, int32_t n_dev) {
  if (!out) return KVM_E_ARG;
  *out = nullptr;
  if (!device_ids || n_dev < 1 || n_dev > 64) return fail(nullptr, KVM_E_ARG, "1..64 devices");
  kvm_multi* M = new kvm_multi();
  for (int d = 0; d < n_dev; d++) {
    kvm_ctx* c = nullptr;
    const int rc = kvm_create(&c, device_ids[d]);
    if (rc) {
      for (kvm_ctx* x : M->ctx) kvm_destroy(x);
      delete M;
      return rc;  // (message in kvm_last_error(NULL))
    }
    M->ctx.push_back(c);
  }
  *out = M;
  return KVM_OK;
}

void kvm_multi_destroy(kvm_multi* M) {
  if (!M) return;
  for (kvm_ctx* c : M->ctx) kvm_destroy(c);
  delete M;
}

const char* kvm_multi_last_error(const kvm_multi* M) { return M ? M->err.c_str() : g_create_error.c_str(); }

int32_t kvm_multi_devices(const kvm_multi* M) { return M ? (int32_t)M->ctx.size() : 0; }

// Shard `samples` (the whole series, length n) over the devices: device d owns window starts
// [d*per + 1, (d+1)*per] with per = ceil(n / n_dev) rounded up to a multiple of `grid` (pass the chain chunk of an
// index-free scan so that no chain straddles two devices; 1 otherwise) and holds `halo` more samples behind them
// (at least the longest query minus one; longer if single intervals are longer than that).
int kvm_multi_load_series_host(kvm_multi* M, const double* samples, int64_t n, int64_t halo, int64_t grid) {
  if (!M) return KVM_E_ARG;
  if (!samples || n < 1 || halo < 0 || grid < 1) {
    M->err = "null/invalid argument";
    return KVM_E_ARG;
  }
  const int D = (int)M->ctx.size();
  int64_t per = (n + D - 1) / D;
  per = (per + grid - 1) / grid * grid;
  M->own_lo.assign(D, 0);
  M->own_hi.assign(D, -1);
  M->n = n;
  std::vector<int> rcs(D, KVM_OK);
  std::vector<std::thread> th;
  for (int d = 0; d < D; d++) {
    const int64_t lo = (int64_t)d * per + 1, hi = std::min<int64_t>(n, (int64_t)(d + 1) * per);
    if (lo > n) continue;  // more devices than grid cells: this one stays empty
    M->own_lo[d] = lo;
    M->own_hi[d] = hi;
    const int64_t last = std::min<int64_t>(n, hi + halo);
    th.emplace_back([=, &rcs]() { rcs[d] = kvm_load_series_host(M->ctx[d], samples + (lo - 1), n, lo, last - lo + 1); });
  }
  for (std::thread& t : th) t.join();
  for (int d = 0; d < D; d++)
    if (rcs[d]) {
      M->err = kvm_last_error(M->ctx[d]);
      return rcs[d];
    }
  return KVM_OK;
}

// One verification call over all devices.  engine: KVM_ENGINE_ED / _CNSM_ED / _DTW / _CNSM_DTW (the unused parameters
// are ignored).  out is what the single-device entry returns for the whole series: answers in ascending offset order,
// counters summed, kernel_ms / stage_ms = the slowest device's.
int kvm_multi_verify(kvm_multi* M, int32_t engine, const double* q, int32_t m, double epsilon, int32_t rho, double alpha,
                     double beta, const int32_t* lr, int32_t K, int32_t shift, kvm_result* out) {
  if (!M) return KVM_E_ARG;
  if (!out || !q || K < 0 || (K > 0 && !lr) || engine < KVM_ENGINE_ED || engine > KVM_ENGINE_CNSM_DTW) {
    M->err = "null/invalid argument";
    return KVM_E_ARG;
  }
  std::memset(out, 0, sizeof(*out));
  const int D = (int)M->ctx.size();
  // intervals -> devices by their first scanned sample max(left - shift, 1)
  std::vector<std::vector<int32_t>> part(D);
  for (int p = 0; p < K; p++) {
    const int64_t begin = std::max<int64_t>((int64_t)lr[2 * p] - shift, 1);
    int d = 0;
    while (d + 1 < D && M->own_hi[d] >= 0 && begin > M->own_hi[d]) d++;
    if (M->own_hi[d] < 0) {
      M->err = "an interval starts beyond the loaded series";
      return KVM_E_RANGE;
    }
    part[d].push_back(lr[2 * p]);
    part[d].push_back(lr[2 * p + 1]);
  }
  std::vector<kvm_result> res(D);
  std::vector<int> rcs(D, KVM_OK);
  std::vector<std::thread> th;
  for (int d = 0; d < D; d++) {
    std::memset(&res[d], 0, sizeof(kvm_result));
    if (M->own_hi[d] < 0) continue;
    th.emplace_back([&, d]() {
      const int32_t* iv = part[d].data();
      const int32_t k = (int32_t)(part[d].size() / 2);
      kvm_ctx* c = M->ctx[d];
      switch (engine) {
        case KVM_ENGINE_ED: rcs[d] = kvm_verify_ed(c, q, m, epsilon, iv, k, shift, &res[d]); break;
        case KVM_ENGINE_CNSM_ED: rcs[d] = kvm_verify_cnsm_ed(c, q, m, epsilon, alpha, beta, iv, k, shift, &res[d]); break;
        case KVM_ENGINE_DTW: rcs[d] = kvm_verify_dtw(c, q, m, epsilon, rho, iv, k, shift, &res[d]); break;
        default: rcs[d] = kvm_verify_cnsm_dtw(c, q, m, epsilon, rho, alpha, beta, iv, k, shift, &res[d]); break;
      }
    });
  }
  for (std::thread& t : th) t.join();
  for (int d = 0; d < D; d++)
    if (rcs[d]) {
      M->err = std::string("device ") + std::to_string(d) + ": " + kvm_last_error(M->ctx[d]);
      return rcs[d];
    }
  M->off.clear();
  M->dist.clear();
  for (int d = 0; d < D; d++) {
    const kvm_result& r = res[d];
    M->off.insert(M->off.end(), r.offsets, r.offsets + r.count);
    M->dist.insert(M->dist.end(), r.distances, r.distances + r.count);
    out->cnt_candidate += r.cnt_candidate;
    out->n_verified += r.n_verified;
    out->s_total += r.s_total;
    out->n_gate_pass += r.n_gate_pass;
    out->n_lb_pass += r.n_lb_pass;
    out->n_exact += r.n_exact;
    out->n_rewalked += r.n_rewalked;
    out->n_dtw_cells += r.n_dtw_cells;
    out->n_chains_rewalked += r.n_chains_rewalked;
    out->n_launches += r.n_launches;
    out->h2d_bytes += r.h2d_bytes;
    out->kernel_ms = std::max(out->kernel_ms, r.kernel_ms);
    for (int i = 0; i < 4; i++) out->stage_ms[i] = std::max(out->stage_ms[i], r.stage_ms[i]);
  }
  out->count = (int64_t)M->off.size();
  out->offsets = M->off.data();
  out->distances = M->dist.data();
  return KVM_OK;
}

// ---- multi-GPU tail inside the library: NCCL resolved at run time (no link dependency; a process that already
// loaded an NCCL, e.g. through torch, shares that copy) ---------------------------------------------------------------
namespace {
struct NcclId {  // ncclUniqueId: 128 opaque bytes, passed by value
  char b[128];
};
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(NcclId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
      api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
      api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
      api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
      api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
      if (api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy) api.h = h;
    }
  }
  return api.h ? &api : nullptr;
}
constexpr int kNcclFloat64 = 8;  // ncclDouble
constexpr int kGatherHead = 16;  // doubles in front of the answers of one rank
constexpr int kGatherCap = 256;  // answers per rank carried by the first (fixed-size) all-gather

int nccl_fail(kvm_ctx* ctx, NcclApi* N, int rc, const char* what) {
  return fail(ctx, KVM_E_NCCL, "%s: %s", what, N && N->GetErrorString ? N->GetErrorString(rc) : "NCCL error");
}
}  // namespace

static void kvm_comm_release(kvm_ctx* ctx) {
  if (ctx->xattached) {
    for (int r = 0; r < ctx->comm_world; r++)
      if (r != ctx->comm_rank && ctx->xpeer[r]) cudaIpcCloseMemHandle(ctx->xpeer[r]);
    ctx->xattached = false;
  }
  ctx->xbuf.release();
  ctx->xerr.release();
  if (ctx->comm) {
    if (NcclApi* N = nccl_api()) N->CommDestroy(ctx->comm);
    ctx->comm = nullptr;
  }
}

extern "C" {

int kvm_comm_unique_id(unsigned char* id128) {
  NcclApi* N = nccl_api();
  if (!N || !id128) return KVM_E_NCCL;
  NcclId id;
  if (N->GetUniqueId(&id) != 0) return KVM_E_NCCL;
  std::memcpy(id128, id.b, 128);
  return KVM_OK;
}

int kvm_comm_init(kvm_ctx* ctx, const unsigned char* id128, int32_t rank, int32_t world) {
  if (!ctx) return KVM_E_ARG;
  if (!id128 || world < 1 || rank < 0 || rank >= world) return fail(ctx, KVM_E_ARG, "null/invalid argument");
  NcclApi* N = nccl_api();
  if (!N) return fail(ctx, KVM_E_NCCL, "libnccl.so.2 could not be loaded");
  KVM_CUDA(ctx, cudaSetDevice(ctx->device));
  kvm_comm_release(ctx);  // a second init replaces the communicator and drops the peer-memory mappings of the first
  NcclId id;
  std::memcpy(id.b, id128, 128);
  const int rc = N->CommInitRank(&ctx->comm, world, id, rank);
  if (rc != 0) {
    ctx->comm = nullptr;
    return nccl_fail(ctx, N, rc, "ncclCommInitRank");
  }
  ctx->comm_rank = rank;
  ctx->comm_world = world;
  return KVM_OK;
}

int kvm_comm_ipc_handle(kvm_ctx* ctx, unsigned char* handle64) {
  if (!ctx) return KVM_E_ARG;
  if (!handle64) return fail(ctx, KVM_E_ARG, "null/invalid argument");
  if (!ctx->comm) return fail(ctx, KVM_E_STATE, "kvm_comm_init has not been called on this ctx");
  if (ctx->comm_world > 8) return fail(ctx, KVM_E_ARG, "the peer-memory exchange serves up to 8 ranks");
  KVM_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t bytes = sizeof(double) * 2 * (size_t)ctx->comm_world * kvm::kXchgSlot + sizeof(unsigned long long) * 2 * (size_t)ctx->comm_world;
  // (a dedicated allocation: cudaIpcGetMemHandle exports the whole cudaMalloc block)
  ctx->xbuf.release();
  KVM_CUDA(ctx, cudaMalloc(&ctx->xbuf.p, bytes));
  ctx->xbuf.cap = bytes;
  KVM_CUDA(ctx, cudaMemsetAsync(ctx->xbuf.p, 0, bytes, ctx->stream));
  KVM_CUDA(ctx, ctx->xerr.ensure(sizeof(int)));
  KVM_CUDA(ctx, cudaMemsetAsync(ctx->xerr.p, 0, sizeof(int), ctx->stream));
  KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  cudaIpcMemHandle_t h;
  KVM_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->xbuf.p));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  std::memcpy(handle64, &h, 64);
  return KVM_OK;
}

int kvm_comm_ipc_attach(kvm_ctx* ctx, const unsigned char* handles) {
  if (!ctx) return KVM_E_ARG;
  if (!handles) return fail(ctx, KVM_E_ARG, "null/invalid argument");
  if (!ctx->comm || !ctx->xbuf.p) return fail(ctx, KVM_E_STATE, "kvm_comm_ipc_handle has not been called on this ctx");
  KVM_CUDA(ctx, cudaSetDevice(ctx->device));
  for (int r = 0; r < ctx->comm_world; r++) {
    if (r == ctx->comm_rank) {
      ctx->xpeer[r] = ctx->xbuf.p;
      continue;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + 64 * (size_t)r, 64);
    KVM_CUDA(ctx, cudaIpcOpenMemHandle(&ctx->xpeer[r], h, cudaIpcMemLazyEnablePeerAccess));
  }
  ctx->xattached = true;
  ctx->xseq = 0;
  return KVM_OK;
}

int kvm_gather_result(kvm_ctx* ctx, const kvm_result* local, kvm_result* merged, double* best_distance, int32_t* best_offset) {
  if (!ctx) return KVM_E_ARG;
  if (!local || !merged || local == merged) return fail(ctx, KVM_E_ARG, "null/invalid argument (local and merged must be distinct)");
  if (!ctx->comm) return fail(ctx, KVM_E_STATE, "kvm_comm_init has not been called on this ctx");
  NcclApi* N = nccl_api();
  const int W = ctx->comm_world;
  KVM_CUDA(ctx, cudaSetDevice(ctx->device));
  // Best match of this rank: lowest distance, lowest offset among equals (the reference's stable sort by distance over
  // the scan order, K/QueryEngine.java:373-376)
  double bd = INFINITY, bo = 4e18;
  for (int64_t i = 0; i < local->count; i++)
    if (local->distances[i] < bd) {  // (offsets ascend: the first minimum is the lowest offset)
      bd = local->distances[i];
      bo = (double)local->offsets[i];
    }
  double exchange_ms = 0.0;  // device time of the exchange (CUDA events on the ctx stream: copies + collective)
  auto round_trip = [&](size_t len_per_rank, auto fill) -> int {
    KVM_CUDA(ctx, ctx->g_hsend.ensure(sizeof(double) * (len_per_rank + 2)));
    KVM_CUDA(ctx, ctx->g_hrecv.ensure(sizeof(double) * len_per_rank * W));
    KVM_CUDA(ctx, ctx->g_send.ensure(sizeof(double) * len_per_rank));
    KVM_CUDA(ctx, ctx->g_recv.ensure(sizeof(double) * len_per_rank * W));
    fill(static_cast<double*>(ctx->g_hsend.p));
    KVM_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
    static const int p2p_on = env_int("KVM_GATHER_P2P", 1);
    if (p2p_on && ctx->xattached && len_per_rank == (size_t)kvm::kXchgSlot) {
      // round 1 over peer memory: one kernel pushes this rank's block to every rank and waits for theirs
      kvm::XchgParams X{};
      for (int r = 0; r < W; r++) X.peer[r] = static_cast<double*>(ctx->xpeer[r]);
      X.send = static_cast<const double*>(ctx->g_hsend.p);
      X.rank = ctx->comm_rank;
      X.world = W;
      ctx->xseq++;
      X.seq = ctx->xseq;
      X.parity = (int)(ctx->xseq & 1ULL);
      X.err = ctx->xerr.as<int>();
      kvm::xchg_kernel<<<1, 256, 0, ctx->stream>>>(X);
      KVM_CUDA(ctx, cudaGetLastError());
      const double* area = static_cast<const double*>(ctx->xbuf.p) + (size_t)X.parity * W * kvm::kXchgSlot;
      KVM_CUDA(ctx, cudaMemcpyAsync(ctx->g_hrecv.p, area, sizeof(double) * len_per_rank * W, cudaMemcpyDeviceToHost, ctx->stream));
      int* herr = reinterpret_cast<int*>(static_cast<double*>(ctx->g_hsend.p) + len_per_rank);  // (spare pinned slot behind the block)
      KVM_CUDA(ctx, cudaMemcpyAsync(herr, ctx->xerr.p, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
      KVM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
      KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
      exchange_ms += ms;
      if (*herr) return fail(ctx, KVM_E_NCCL, "peer-memory exchange timed out: a rank did not arrive");
      return KVM_OK;
    }
    // The blocks are a few KB: the collective reads and writes the pinned host buffers directly (they are device
    // accessible through unified addressing), which saves the two staging copies around it.  KVM_GATHER_ZEROCOPY=0
    // stages through device buffers instead.
    static const int zero_copy = env_int("KVM_GATHER_ZEROCOPY", 1);
    if (zero_copy) {
      const int rc = N->AllGather(ctx->g_hsend.p, ctx->g_hrecv.p, len_per_rank, kNcclFloat64, ctx->comm, ctx->stream);
      if (rc != 0) return nccl_fail(ctx, N, rc, "ncclAllGather");
    } else {
      KVM_CUDA(ctx, cudaMemcpyAsync(ctx->g_send.p, ctx->g_hsend.p, sizeof(double) * len_per_rank, cudaMemcpyHostToDevice, ctx->stream));
      const int rc = N->AllGather(ctx->g_send.p, ctx->g_recv.p, len_per_rank, kNcclFloat64, ctx->comm, ctx->stream);
      if (rc != 0) return nccl_fail(ctx, N, rc, "ncclAllGather");
      KVM_CUDA(ctx, cudaMemcpyAsync(ctx->g_hrecv.p, ctx->g_recv.p, sizeof(double) * len_per_rank * W, cudaMemcpyDeviceToHost, ctx->stream));
    }
    KVM_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
    KVM_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
    exchange_ms += ms;
    return KVM_OK;
  };
  // ---- round 1, fixed size: header + the first kGatherCap answers of every rank
  const size_t len1 = kGatherHead + 2 * (size_t)kGatherCap;
  int rc = round_trip(len1, [&](double* h) {
    std::memset(h, 0, sizeof(double) * len1);
    h[0] = (double)local->count;
    h[1] = (double)local->n_verified;
    h[2] = (double)local->cnt_candidate;
    h[3] = (double)local->s_total;
    h[4] = (double)local->n_gate_pass;
    h[5] = (double)local->n_lb_pass;
    h[6] = (double)local->n_exact;
    h[7] = (double)local->n_rewalked;
    h[8] = (double)local->n_chains_rewalked;
    h[9] = (double)local->n_dtw_cells;
    h[10] = local->kernel_ms;
    h[11] = bd;
    h[12] = bo;
    h[13] = (double)local->n_launches;
    const int64_t c = std::min<int64_t>(local->count, kGatherCap);
    for (int64_t i = 0; i < c; i++) {
      h[kGatherHead + i] = (double)local->offsets[i];
      h[kGatherHead + kGatherCap + i] = local->distances[i];
    }
  });
  if (rc) return rc;
  const double* a = static_cast<const double*>(ctx->g_hrecv.p);
  int64_t total = 0, max_count = 0;
  std::memset(merged, 0, sizeof(*merged));
  double gbd = INFINITY, gbo = 4e18;
  for (int r = 0; r < W; r++) {
    const double* h = a + (size_t)r * len1;
    total += (int64_t)h[0];
    max_count = std::max<int64_t>(max_count, (int64_t)h[0]);
    merged->n_verified += (int64_t)h[1];
    merged->cnt_candidate += (int64_t)h[2];
    merged->s_total += (int64_t)h[3];
    merged->n_gate_pass += (int64_t)h[4];
    merged->n_lb_pass += (int64_t)h[5];
    merged->n_exact += (int64_t)h[6];
    merged->n_rewalked += (int64_t)h[7];
    merged->n_chains_rewalked += (int64_t)h[8];
    merged->n_dtw_cells += (int64_t)h[9];
    merged->kernel_ms = std::max(merged->kernel_ms, h[10]);
    if (h[11] < gbd || (h[11] == gbd && h[12] < gbo)) {
      gbd = h[11];
      gbo = h[12];
    }
    merged->n_launches += (int32_t)h[13];
  }
  ctx->g_off.clear();
  ctx->g_dist.clear();
  ctx->g_off.reserve((size_t)total);
  ctx->g_dist.reserve((size_t)total);
  if (max_count <= kGatherCap) {
    for (int r = 0; r < W; r++) {  // ranks own ascending offset ranges: concatenation is the scan order
      const double* h = a + (size_t)r * len1;
      for (int64_t i = 0; i < (int64_t)h[0]; i++) {
        ctx->g_off.push_back((int32_t)h[kGatherHead + i]);
        ctx->g_dist.push_back(h[kGatherHead + kGatherCap + i]);
      }
    }
  } else {
    // ---- round 2 (a rank holds more answers than the fixed block carries; every rank sees the same max_count, so
    // every rank takes this branch): all answers, padded to the largest count
    std::vector<int64_t> counts(W);
    for (int r = 0; r < W; r++) counts[r] = (int64_t)a[(size_t)r * len1];
    const size_t len2 = 2 * (size_t)max_count;
    rc = round_trip(len2, [&](double* h) {
      for (int64_t i = 0; i < local->count; i++) {
        h[i] = (double)local->offsets[i];
        h[max_count + i] = local->distances[i];
      }
    });
    if (rc) return rc;
    const double* b = static_cast<const double*>(ctx->g_hrecv.p);
    for (int r = 0; r < W; r++)
      for (int64_t i = 0; i < counts[r]; i++) {
        ctx->g_off.push_back((int32_t)b[(size_t)r * len2 + i]);
        ctx->g_dist.push_back(b[(size_t)r * len2 + max_count + i]);
      }
  }
  merged->stage_ms[0] = exchange_ms;
  merged->count = total;
  merged->offsets = ctx->g_off.data();
  merged->distances = ctx->g_dist.data();
  if (best_distance) *best_distance = gbd;
  if (best_offset) *best_offset = total > 0 ? (int32_t)gbo : 0;
  return KVM_OK;
}

}  // extern "C"

void kvm_result_free(kvm_ctx* ctx, kvm_result* r) {
  (void)ctx;
  if (!r) return;
  r->count = 0;
  r->offsets = nullptr;
  r->distances = nullptr;
}

void kvm_runs_free(kvm_ctx* ctx, kvm_runs* r) {
  (void)ctx;
  if (!r) return;
  r->count = 0;
  r->keys = nullptr;
  r->first = r->last = nullptr;
}

}  // extern "C"
