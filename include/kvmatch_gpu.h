/* kvmatch_gpu.h — C ABI of libkvmatch_gpu.so: KV-match phase-2 candidate verification and
 * IndexBuilder's sliding-window mean pass on one NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary for ONE hot path of DSM-fudan/KV-match.  The reference has no
 * native interface (phase 2 is inline Java), so each entry point below states the reference
 * loop it replaces.  K/ = src/main/java/cn/edu/fudan/dsm/kvmatch/ in the reference tree.
 *
 * Conventions
 *   - plain C, no C++ types, no exceptions; every call returns KVM_OK (0) or a negative code and
 *     records a message retrievable with kvm_last_error().
 *   - offsets are the reference's: 1-based int32 positions in the whole series of length n.
 *   - arithmetic is IEEE binary64 without FMA contraction wherever a value is reported or decides
 *     an answer, so results equal the reference's Java loops (see DESIGN.md "Parity").
 *   - inputs are caller-owned and only read during the call.  Result buffers are library-owned
 *     pinned host memory, valid until the next call on the same ctx or kvm_result_free().
 *   - one kvm_ctx = one GPU = one engine instance; a ctx is not re-entrant; distinct ctxs are
 *     independent.  Multi-GPU: one process (or ctx) per GPU, each holding an offset range of the
 *     series plus a halo (kvm_load_series_host's `first`), answers merged by the host.
 *   - NO CPU FALLBACK: without a usable sm_100 device kvm_create fails with KVM_E_NODEVICE.
 */
#ifndef KVMATCH_GPU_H_
#define KVMATCH_GPU_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KVM_ABI_VERSION 4 /* 4: + kvm_norm_intervals_*, kvm_index_row_positions (additions only) */

enum {
  KVM_OK = 0,
  KVM_E_NODEVICE = -1, /* no CUDA device / not compute capability 10.x */
  KVM_E_ARG = -2,      /* invalid argument */
  KVM_E_OOM = -3,      /* device or pinned-host allocation failed */
  KVM_E_CUDA = -4,     /* CUDA runtime error (message has the detail) */
  KVM_E_IO = -5,       /* series file could not be read */
  KVM_E_STATE = -6,    /* no series loaded */
  KVM_E_RANGE = -7,    /* an interval needs samples outside [1,n] or outside this ctx's shard; the
                          reference throws IllegalArgumentException for the former
                          (K/operator/file/TimeSeriesFileOperator.java:55-57) */
  KVM_E_NCCL = -8      /* NCCL could not be loaded or a collective failed (kvm_comm_*, kvm_gather_result) */
};

typedef struct kvm_ctx kvm_ctx;

/* Answers of one verification call, in ascending offset order (= the reference's scan order; the
 * Java caller then stable-sorts by distance, K/QueryEngine.java:373). */
typedef struct kvm_result {
  int64_t count;           /* #answers */
  const int32_t* offsets;  /* 1-based window starts */
  const double* distances; /* sqrt(dist^2), K/QueryEngine.java:359 */
  int64_t cnt_candidate;   /* sum(right-left+1) over the intervals, unclamped: the reference's #candidates */
  int64_t n_verified;      /* window starts examined (after clamping to [1,n]) */
  int64_t s_total;         /* series samples those windows cover, each counted once per interval */
  int64_t n_gate_pass;     /* cNSM: windows passing the exact alpha/beta gate */
  int64_t n_lb_pass;       /* DTW: windows surviving the GPU lower bounds, i.e. full DTWs computed */
  int64_t n_exact;         /* ED: windows re-evaluated by the sequential reference-order path; DTW: candidates that reached the band DTW (survivors of the corner probe, when it ran) */
  double kernel_ms;        /* device time of this call's kernels (CUDA events on the ctx stream) */
  double stage_ms[4];      /* the same, per stage.  cNSM engines: [0] streaming statistics pass (gate + in-stream lower
                              bound), [1] exact re-walk of the flagged chains, [2] exact stage (cNSM-ED: exact gate +
                              reference-order sum; cNSM-DTW: exact gate + lower bounds), [3] cNSM-DTW: banded DTW.
                              RSM engines: [0] ED scan / raw LB scan, [2] banded DTW. */
  int32_t n_launches;      /* kernels launched by this call */
  int32_t h2d_bytes;       /* bytes this call copied host->device (query, and the interval plan unless the previous
                              call's plan was reused) */
  int64_t n_rewalked;        /* cNSM: windows whose chain sums were recomputed exactly (ambiguous gate or surviving the
                                in-stream lower bound) */
  int64_t n_chains_rewalked; /* cNSM: statistic chains (merged intervals) walked exactly for them */
  int64_t n_dtw_cells;       /* DTW: band cells evaluated by the DTW kernel (5 FP64 operations each) */
} kvm_result;

/* IndexBuilder step-1 output for one window width w: the (key, first, last) intervals in the order
 * K/IndexBuilder.java:268-286 appends them; last-first <= 254. */
typedef struct kvm_runs {
  int64_t count;
  const double* keys;   /* MeanIntervalUtils.toRound(mean), K/utils/MeanIntervalUtils.java:51-61 */
  const int32_t* first; /* 1-based */
  const int32_t* last;
  double kernel_ms;
  int32_t n_launches;
  int32_t reserved;
} kvm_runs;

int kvm_abi_version(void);

/* Create a context on CUDA device `device_id`.  Replaces the engines' constructor choice of a
 * TimeSeriesOperator (K/QueryEngine.java:70-96): the series lives in HBM for the ctx lifetime. */
int kvm_create(kvm_ctx** out, int device_id);
void kvm_destroy(kvm_ctx* ctx);
const char* kvm_last_error(const kvm_ctx* ctx); /* ctx may be NULL: last kvm_create error */

/* Per-ctx options (defaults in parentheses; the environment variables KVM_CNSM_PATH=relay, KVM_STREAM_FORCE_ALL=1,
 * KVM_PLAN_CACHE=0 set the defaults of new contexts).  None of them changes a result.
 *   KVM_OPT_CNSM_PATH        KVM_CNSM_STREAM (default): the cNSM statistics run as an HBM stream with guard bands and an
 *                            exact re-walk of the ambiguous / surviving windows; KVM_CNSM_RELAY: every chain is walked
 *                            exactly (the round-1 kernel; also what queries too long for the stream's tile use)
 *   KVM_OPT_STREAM_FLAG_ALL  1: the stream flags every window inside its outer gate band, so the exact stages decide
 *                            everything (a self-check of the guard bands; default 0)
 *   KVM_OPT_PLAN_CACHE       1 (default): an interval list identical to the previous call's reuses its device plan */
enum { KVM_OPT_CNSM_PATH = 1, KVM_OPT_STREAM_FLAG_ALL = 2, KVM_OPT_PLAN_CACHE = 3 };
enum { KVM_CNSM_STREAM = 0, KVM_CNSM_RELAY = 1 };
int kvm_set_option(kvm_ctx* ctx, int32_t option, int64_t value);

/* Load samples [first, first+count-1] (1-based) of a series whose total length is n.
 * first=1,count=n loads everything (single GPU).  Replaces TimeSeriesOperator.readTimeSeries
 * (K/operator/TimeSeriesOperator.java:38) as the data feed of phase 2. */
int kvm_load_series_host(kvm_ctx* ctx, const double* samples, int64_t n, int64_t first, int64_t count);
/* Same, from a reference data file files/data-N: N big-endian IEEE-754 doubles, no header
 * (K/DataGenerator.java:102-113); the byte swap runs on the device. */
int kvm_load_series_file(kvm_ctx* ctx, const char* path, int64_t n, int64_t first, int64_t count);

/* Phase-2 verification.  q: raw query (length m); lr: K merged intervals (left,right) as produced
 * by sortAndMergeIntervals (K/QueryEngine.java:664-693), sorted by left; shift =
 * (lastSegment-1)*25.  Each interval p is scanned over samples
 * [max(left-shift,1), min(right-shift+m-1, n)]; running statistics restart per interval.
 *
 * kvm_verify_ed       replaces K/QueryEngine.java:341-363        (RSM-ED)
 * kvm_verify_cnsm_ed  replaces K/NormQueryEngine.java:432-528    (cNSM-ED)
 * kvm_verify_dtw      replaces K/QueryEngineDtw.java:349-452     (RSM-DTW, rho = Sakoe-Chiba radius)
 * kvm_verify_cnsm_dtw replaces K/NormQueryEngineDtw.java:457-603 (cNSM-DTW)
 */
int kvm_verify_ed(kvm_ctx* ctx, const double* q, int32_t m, double epsilon, const int32_t* lr, int32_t K,
                  int32_t shift, kvm_result* out);
int kvm_verify_cnsm_ed(kvm_ctx* ctx, const double* q, int32_t m, double epsilon, double alpha, double beta,
                       const int32_t* lr, int32_t K, int32_t shift, kvm_result* out);
int kvm_verify_dtw(kvm_ctx* ctx, const double* q, int32_t m, double epsilon, int32_t rho, const int32_t* lr,
                   int32_t K, int32_t shift, kvm_result* out);
int kvm_verify_cnsm_dtw(kvm_ctx* ctx, const double* q, int32_t m, double epsilon, int32_t rho, double alpha,
                        double beta, const int32_t* lr, int32_t K, int32_t shift, kvm_result* out);

/* Query set (an addition to the reference's per-query engines): n_queries (<= 16) raw queries of one length m, stored
 * back to back, verified over ONE interval list with the cNSM-ED semantics of kvm_verify_cnsm_ed.  The window
 * statistics (K/NormQueryEngine.java:498-499,523-524) do not depend on the query, so one statistics pass serves the
 * whole set; outs[q] is what kvm_verify_cnsm_ed(queries + q*m, ...) returns.  outs[q].offsets / distances stay valid
 * until the next query-set call on this ctx (kvm_result_free is not needed for them).  What the reference's experiment
 * drivers do query by query (K/experiments/NormQueryTestGroupBySelectivity.java:110-127). */
int kvm_verify_cnsm_ed_batch(kvm_ctx* ctx, const double* queries, int32_t n_queries, int32_t m, double epsilon,
                             double alpha, double beta, const int32_t* lr, int32_t K, int32_t shift, kvm_result* outs);

/* Index-free full scan with the semantics of the reference's UCR-DTW baseline executor
 * (K/experiments/ucr/UcrDtwQueryExecutor.java:84-314): the series is streamed in EPOCH = 100000-sample buffers that
 * overlap by m-1 (:97-131), the running statistics restart per buffer (:133-134), every window goes through the
 * alpha/beta gate and the LB_Kim / LB_Keogh / DTW cascade, and answers carry 0-BASED offsets (:278).  Equivalent to
 * kvm_verify_cnsm_dtw over the intervals [it*(EPOCH-m+1)+1, ...] with shift 0.  Needs the whole series on this ctx.
 */
int kvm_scan_ucr_dtw(kvm_ctx* ctx, const double* q, int32_t m, double epsilon, int32_t rho, double alpha, double beta,
                     kvm_result* out);

/* The ED baseline executor (K/experiments/ucr/UcrEdQueryExecutor.java:101-183): every window of the series through the
 * alpha/beta gate and the |zQ|-ordered early-abandoning distance, with ONE statistics chain that is never restarted
 * (:138-176) and 1-BASED offsets (:166).  Equivalent to kvm_verify_cnsm_ed over the single interval [1, n-m+1]: the
 * streaming pass screens every window in parallel, but each window that needs the chain's exact sums is re-walked from
 * sample 1 by one warp (about 16 cycles per sample: ~10 ms per 1e6 samples up to the last such window), which is the
 * price of bit-identical statistics on an unbroken chain; on long series prefer kvm_verify_cnsm_ed on a chain grid.
 * The executor's distance loop runs `while sum < eps^2` (:91) where the engines use `<=` (K/NormQueryEngine.java:516):
 * they differ only when a partial sum equals eps^2 exactly (the eps-tie exemption).  Needs the whole series on this ctx.
 * The PAA variants (PaaUcrEdQueryExecutor / PaaUcrDtwQueryExecutor) add LB_PAA and a triangle-inequality skip, both
 * pruning only: for eps >= 1 their answer set is this scan's (out of scope). */
int kvm_scan_ucr_ed(kvm_ctx* ctx, const double* q, int32_t m, double epsilon, double alpha, double beta, kvm_result* out);

/* DtwUtils.lowerUpperLemire on the device (K/utils/DtwUtils.java:50-91, called on the data buffer at
 * K/QueryEngineDtw.java:397-399): lower[i] / upper[i] = min / max of samples [max(first, first+i-r) ..
 * min(first+len-1, first+i+r)] (1-based `first`, i = 0..len-1; r <= 512).  The verification entries do not need it
 * (their third lower bound, LB_Keogh on the data envelope, K/utils/DtwUtils.java:238-257, forms the envelope of each
 * surviving window on the fly); it is the reference's buffer-level operation offered as such.  Caller-owned outputs. */
int kvm_envelope(kvm_ctx* ctx, int32_t r, int64_t first, int32_t len, double* lower, double* upper);

/* IndexBuilder step 1 for window width w (K/IndexBuilder.java:194-301): sliding mean with the
 * reference's EPOCH=100000 restart structure, toRound key, run-length intervals split at 255.
 * Needs the whole series on this ctx (first == 1, count == n). */
int kvm_window_mean_runs(kvm_ctx* ctx, int32_t w, kvm_runs* out);

/* The same for all window widths of an index build at once (the reference's Sigma = {25,50,100,200,400},
 * K/IndexBuilder.java:98-120 runs one SingleIndexBuilder pass per width: its own "TODO: naive"): ONE pass over the
 * series serves up to 5 widths.  outs[q] is what kvm_window_mean_runs(ctx, widths[q], ...) returns; its arrays stay
 * valid until the next window-mean call on this ctx; outs[q].reserved = epochs of that width re-walked exactly. */
int kvm_window_mean_runs_all(kvm_ctx* ctx, const int32_t* widths, int32_t n_widths, kvm_runs* outs);

/* The whole single-width index build: window-mean pass on the GPU, then IndexBuilder step 2 (adjacent-row merge,
 * K/IndexBuilder.java:308-345, K/utils/IndexNodeUtils.java:30-90) and the file image IndexFileOperator.writeAll
 * produces (K/operator/file/IndexFileOperator.java:127-164; row codec K/common/entity/IndexNode.java:51-96; statistic
 * table K/utils/ByteUtils.java:84-100) on the host.  Writes `path` (the reference names it files/index-<N>-<w>);
 * path == NULL only fills `info`.  Replaces SingleIndexBuilder.run() (K/IndexBuilder.java:186-347) for one w. */
typedef struct kvm_index_info {
  int64_t file_bytes;
  int64_t n_runs;        /* step-1 (key, first, last) runs */
  int64_t n_intervals;   /* intervals after the step-2 merge */
  int64_t n_offsets;     /* window positions covered (= n - w + 1 on real data) */
  int32_t n_rows_step1;  /* distinct keys */
  int32_t n_rows;        /* rows after the merge */
  double kernel_ms;      /* GPU time of the window-mean pass */
  double host_ms;        /* step 2 + encoding + write */
} kvm_index_info;
int kvm_build_index_file(kvm_ctx* ctx, int32_t w, const char* path, kvm_index_info* info);
/* Host-only half of the above (no GPU, no ctx): step 2 and the file image from runs the caller already holds, in the
 * order IndexBuilder step 1 appends them.  *image is malloc'ed by the library; release it with kvm_image_free. */
int kvm_index_image_from_runs(const double* keys, const int32_t* first, const int32_t* last, int64_t n_runs,
                              unsigned char** image, kvm_index_info* info);
void kvm_image_free(unsigned char* image);
/* The reader's half (host only): the (left, right) pairs of ONE index row from its compact bytes, the 8-byte key
 * excluded (IndexNode.parseBytesCompact, K/common/entity/IndexNode.java:108-128).  *k_out is always set; KVM_E_ARG if
 * the row holds more than `cap` pairs (row_bytes / 2 always suffices), KVM_E_RANGE on a truncated row. */
int kvm_index_row_positions(const unsigned char* row, int64_t row_bytes, int32_t* lr_out, int64_t cap, int64_t* k_out);

/* ---- phase-1 tail on the host (no ctx, no GPU): the interval algebra between index probing and verification, on
 * plain arrays so that the candidate list reaches kvm_verify_* without boxed Java lists.  Intervals are (left, right)
 * int32 pairs with one lower-bound value (the reference's Interval.epsilon) each; outputs are caller-owned with room
 * for `cap` intervals (*k_out is always set; KVM_E_ARG if it exceeds cap).
 *   kvm_intervals_sort_merge     mode 0 = sortButNotMergeIntervals          K/QueryEngine.java:593-622
 *                                mode 1 = sortButNotMergeIntervalsAndCount   :624-662 (cnt_disjoint, cnt_offsets)
 *                                mode 2 = sortAndMergeIntervals              :664-693 (the list handed to phase 2)
 *   kvm_intervals_intersect      CS ∩ CS_i with summed bounds <= eps2, shifted by delta_w          :282-308
 *   kvm_intervals_first_segment  the first segment's positions clamped to window starts in [1, n-length+1]  :264-280 */
int kvm_intervals_sort_merge(const int32_t* lr, const double* eps, int64_t k, int32_t mode, int32_t* lr_out, double* eps_out,
                             int64_t cap, int64_t* k_out, int64_t* cnt_disjoint, int64_t* cnt_offsets);
int kvm_intervals_intersect(const int32_t* cs_lr, const double* cs_eps, int64_t k1, const int32_t* csi_lr, const double* csi_eps,
                            int64_t k2, double eps2, int32_t delta_w, int32_t* lr_out, double* eps_out, int64_t cap,
                            int64_t* k_out, double* min_eps);
int kvm_intervals_first_segment(const int32_t* lr, const double* eps, int64_t k, int32_t order, int32_t w0, int32_t length,
                                int32_t n, int32_t delta_w, int32_t* lr_out, double* eps_out, int64_t cap, int64_t* k_out,
                                double* min_eps);

/* The same tail for the cNSM engines (K/NormQueryEngine.java, K/NormQueryEngineDtw.java): an interval carries the lower
 * and (DTW engine) upper sums of the segments seen so far (K/common/NormInterval.java exLower / ex2Lower / exUpper /
 * ex2Upper, in blocks of w0 points) and the bit set of beta partitions its index rows fell into.
 *   kvm_norm_intervals_sort_merge     mode 0 = sortButNotMergeIntervals         K/NormQueryEngine.java:788-823 (Dtw :926-967)
 *                                     mode 1 = sortButNotMergeIntervalsAndCount  :825-869 (Dtw :969-1019)
 *                                     mode 2 = sortAndMergeIntervals             :871-896 (Dtw :1021-1046), handed to phase 2
 *   kvm_norm_intervals_intersect      CS ∩ CS_i with the beta-partition and variance filters (ENABLE_BETA_PARTITION,
 *                                     ENABLE_STD_FILTER), shifted by delta_w; pre_length = blocks covered so far;
 *                                     dtw = 0: NormQueryEngine :333-397, dtw = 1: NormQueryEngineDtw :349-425
 *   kvm_norm_intervals_first_segment  the first segment's positions clamped to window starts in [1, n-length+1]
 *                                     :313-332 (Dtw :326-348) */
typedef struct kvm_norm_interval {
  int32_t left, right;
  double ex_lower, ex2_lower;
  double ex_upper, ex2_upper; /* DTW engine only; zero otherwise */
  int64_t beta_partitions;
} kvm_norm_interval;
int kvm_norm_intervals_sort_merge(const kvm_norm_interval* in, int64_t k, int32_t mode, kvm_norm_interval* out, int64_t cap,
                                  int64_t* k_out, int64_t* cnt_disjoint, int64_t* cnt_offsets);
int kvm_norm_intervals_intersect(const kvm_norm_interval* cs, int64_t k1, const kvm_norm_interval* csi, int64_t k2,
                                 int32_t pre_length, int32_t w0, int32_t query_length, double mean_q, double std_q, double alpha,
                                 double beta, int32_t delta_w, int32_t dtw, kvm_norm_interval* out, int64_t cap, int64_t* k_out);
int kvm_norm_intervals_first_segment(const kvm_norm_interval* in, int64_t k, int32_t order, int32_t w0, int32_t length, int32_t n,
                                     int32_t delta_w, kvm_norm_interval* out, int64_t cap, int64_t* k_out);

/* ---- several GPUs behind one handle (one process) -------------------------------------------------
 * The series is sharded by offset range: device d owns window starts [d*per+1, (d+1)*per] (per = ceil(n / n_dev)
 * rounded up to a multiple of `grid`) and holds `halo` more samples behind them.  Every interval is verified by the
 * device that owns its first scanned sample max(left-shift, 1) — statistic chains are never split, so results are
 * bit-identical to one device's — and must end within that device's halo (KVM_E_RANGE otherwise).  One host thread
 * per device drives its context; the host concatenates the answers (device order = offset order), sums the
 * counters and reports the slowest device's times.  No series data and no collective crosses GPUs.  (One process
 * per GPU with an NCCL all_gather of the packed answers is the other supported layout: kvmatch_b200/sharding.py.)
 * Precedent in the reference: the (w-1)-point overlap of K/mapreduce/BuildIndexMapReduce.java:216-221. */
typedef struct kvm_multi kvm_multi;
enum { KVM_ENGINE_ED = 0, KVM_ENGINE_CNSM_ED = 1, KVM_ENGINE_DTW = 2, KVM_ENGINE_CNSM_DTW = 3 };
int kvm_multi_create(kvm_multi** out, const int32_t* device_ids, int32_t n_dev);
void kvm_multi_destroy(kvm_multi* m);
const char* kvm_multi_last_error(const kvm_multi* m);
int32_t kvm_multi_devices(const kvm_multi* m);
int kvm_multi_load_series_host(kvm_multi* m, const double* samples, int64_t n, int64_t halo, int64_t grid);
/* engine selects kvm_verify_ed / _cnsm_ed / _dtw / _cnsm_dtw semantics (parameters the engine does not take are
 * ignored); out->offsets / distances stay valid until the next call on this handle. */
int kvm_multi_verify(kvm_multi* m, int32_t engine, const double* q, int32_t mq, double epsilon, int32_t rho, double alpha,
                     double beta, const int32_t* lr, int32_t K, int32_t shift, kvm_result* out);

/* ---- One process per GPU: the multi-GPU tail inside the library --------------------------------
 * The series shards by offset range, every rank verifies its own slice with the entries above, and the only exchange is
 * the tail: per-rank counts, the sparse answers and the best match.  The reference has no counterpart (single JVM); its
 * MapReduce index build shards the same way (K/mapreduce/BuildIndexMapReduce.java:216-221).
 * kvm_comm_unique_id: rank 0 obtains a 128-byte NCCL id and distributes it to the other ranks by any means (the
 * Python host code broadcasts it through torch.distributed; a Java driver would send it over its own RPC).
 * kvm_comm_init: every rank joins the communicator with its ctx (ncclCommInitRank on the ctx's device).  NCCL is
 * resolved with dlopen("libnccl.so.2") at the first call: the library has no link-time dependency on it.
 * kvm_gather_result: COLLECTIVE — every rank passes the result of its own verify call; ONE fixed-size ncclAllGather on
 * the ctx's stream carries [count, counters, best match, first 256 answers] of every rank (a second, padded one only
 * when some rank holds more); every rank receives the merged result: answers of all ranks in ascending offset order
 * (ranks must own ascending, disjoint offset ranges), counters summed, kernel_ms = max over ranks, stage_ms[0] = device
 * time of the exchange itself (CUDA events around the copies and the collective), and the reference's
 * `Best:` line (lowest distance, lowest offset among equals, K/QueryEngine.java:373-376) in best_distance /
 * best_offset (may be NULL).  merged->offsets / distances are library-owned, valid until the next call on the ctx. */
int kvm_comm_unique_id(unsigned char* id128);
int kvm_comm_init(kvm_ctx* ctx, const unsigned char* id128, int32_t rank, int32_t world);
int kvm_gather_result(kvm_ctx* ctx, const kvm_result* local, kvm_result* merged, double* best_distance,
                      int32_t* best_offset);
/* Optional fast path of kvm_gather_result for ranks on one node (up to 8, NVLink / NVSwitch peer access): the fixed-size
 * round runs as ONE kernel over peer memory instead of an ncclAllGather — every rank writes its 4 KB block straight into
 * every other rank's exchange buffer and waits for theirs (sequence numbers behind a system-scope fence).
 * kvm_comm_ipc_handle: after kvm_comm_init, every rank exports the 64-byte CUDA IPC handle of its exchange buffer;
 * the handles travel like the NCCL id; kvm_comm_ipc_attach(ctx, handles: world x 64 bytes in rank order) maps the
 * peers' buffers.  Results are identical to the NCCL path; a rank that never arrives makes the others return KVM_E_NCCL
 * after ~3 s instead of hanging.  The overflow round (a rank with more than 256 answers) still uses NCCL. */
int kvm_comm_ipc_handle(kvm_ctx* ctx, unsigned char* handle64);
int kvm_comm_ipc_attach(kvm_ctx* ctx, const unsigned char* handles);

void kvm_result_free(kvm_ctx* ctx, kvm_result* r);
void kvm_runs_free(kvm_ctx* ctx, kvm_runs* r);

#ifdef __cplusplus
}
#endif
#endif /* KVMATCH_GPU_H_ */
