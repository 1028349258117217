// kvm_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A single-threaded C++17 restatement of the KV-match phase-2 verification loops and of
// IndexBuilder's sliding-window mean pass, written so that every floating-point operation
// happens in the same order and with the same rounding as the reference's Java code
// (IEEE binary64, no FMA contraction: build with -ffp-contract=off, no fast-math).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library.  The product (libkvmatch_gpu.so) never links or calls it.
//
// PARITY UNPINNED: the reference (Java 8, no JVM in this image) ships no tests, fixtures or
// golden vectors for this path, and cannot be executed here.  What pins this file instead:
// the reference's self-checks (self-match distance exactly 0.0, README "Best: 123456,
// distance: 0.0", the rounding examples in MeanIntervalUtils' doc comments), a second,
// independent numpy restatement in tests/, and algebraic invariants (LB <= DTW, rho=0 DTW ==
// squared ED, Lemire envelope == clamped sliding min/max).
//
// K/ = /root/reference/src/main/java/cn/edu/fudan/dsm/kvmatch/   (citations are file:line)
//
// Assumptions that cannot be confirmed without a JVM: Math.pow(10,1)==10.0 and
// Math.pow(10,-1)==0.1 exactly; HotSpot evaluates double expressions in strict binary64;
// Arrays.sort(Object[]) is a stable sort.

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

namespace {

constexpr double kInf = 1e20;                 // K/utils/DtwUtils.java:24
constexpr int kWu0 = 25;                      // WuList[0], K/QueryEngine.java:51
constexpr int kMaxScanDataLength = 40000;     // K/NormQueryEngine.java:60
constexpr int kEpoch = 100000;                // K/IndexBuilder.java:136, UcrDtwQueryExecutor.java:97
constexpr int kRowBytes = 1000;               // TimeSeriesNode.ROW_LENGTH used as a BYTE count by the
                                              // file iterator, K/operator/file/TimeSeriesFileOperator.java:100
constexpr int kMaximumDiff = 256;             // K/common/entity/IndexNode.java:31

inline double sqdist(double a, double b) { return (a - b) * (a - b); }   // DtwUtils.java:38-40
inline double dmin(double a, double b) { return (a < b) ? a : b; }       // DtwUtils.java:34-36

inline uint64_t double_equals_bits(double v) {
  // Double.equals compares doubleToLongBits, which canonicalises NaN.
  if (v != v) return 0x7ff8000000000000ULL;
  uint64_t b;
  std::memcpy(&b, &v, 8);
  return b;
}

// Double.compare(a, b)
inline int java_double_compare(double a, double b) {
  if (a < b) return -1;
  if (a > b) return 1;
  uint64_t x = double_equals_bits(a), y = double_equals_bits(b);
  int64_t sx = (int64_t)x, sy = (int64_t)y;
  return (sx == sy) ? 0 : (sx < sy ? -1 : 1);
}

// ---- K/common/CircularArray.java:38-109 : the index deque used by the Lemire envelope ----
class IndexDeque {
 public:
  explicit IndexDeque(int capacity) : cap_(capacity), size_(0), f_(0), r_(capacity - 1), dq_(capacity, 0) {}
  void push_back(int v) {
    dq_[r_] = v;
    if (--r_ < 0) r_ = cap_ - 1;
    ++size_;
  }
  void pop_front() {
    if (--f_ < 0) f_ = cap_ - 1;
    --size_;
  }
  void pop_back() {
    r_ = (r_ + 1) % cap_;
    --size_;
  }
  int front() const {
    int a = f_ - 1;
    if (a < 0) a = cap_ - 1;
    return dq_[a];
  }
  int back() const { return dq_[(r_ + 1) % cap_]; }
  bool empty() const { return size_ == 0; }

 private:
  int cap_, size_, f_, r_;
  std::vector<int> dq_;
};

// ---- K/utils/DtwUtils.java:50-91 (array) and :93-134 (List) — same algorithm ----
// Returns false where the reference would throw ArrayIndexOutOfBounds (len < r + 1).
bool lemire_envelope(const double* t, int len, int r, double* l, double* u) {
  if (r < 0 || len < r + 1 || len < 1) return false;
  IndexDeque du(2 * r + 2), dl(2 * r + 2);
  du.push_back(0);
  dl.push_back(0);
  for (int i = 1; i < len; i++) {
    if (i > r) {
      u[i - r - 1] = t[du.front()];
      l[i - r - 1] = t[dl.front()];
    }
    if (t[i] > t[i - 1]) {
      du.pop_back();
      while (!du.empty() && t[i] > t[du.back()]) du.pop_back();
    } else {
      dl.pop_back();
      while (!dl.empty() && t[i] < t[dl.back()]) dl.pop_back();
    }
    du.push_back(i);
    dl.push_back(i);
    if (i == 2 * r + 1 + du.front()) {
      du.pop_front();
    } else if (i == 2 * r + 1 + dl.front()) {
      dl.pop_front();
    }
  }
  for (int i = len; i < len + r + 1; i++) {
    u[i - r - 1] = t[du.front()];
    l[i - r - 1] = t[dl.front()];
    if (i - du.front() >= 2 * r + 1) du.pop_front();
    if (i - dl.front() >= 2 * r + 1) dl.pop_front();
  }
  return true;
}

// ---- K/utils/DtwUtils.java:149-189 ----
double lb_kim_hierarchy(const double* t, const double* q, int j, int len, double mean, double std_, double bsf) {
  double d, lb;
  double x0 = (t[j] - mean) / std_;
  double y0 = (t[len - 1 + j] - mean) / std_;
  lb = sqdist(x0, q[0]) + sqdist(y0, q[len - 1]);
  if (lb >= bsf) return lb;

  double x1 = (t[j + 1] - mean) / std_;
  d = dmin(sqdist(x1, q[0]), sqdist(x0, q[1]));
  d = dmin(d, sqdist(x1, q[1]));
  lb += d;
  if (lb >= bsf) return lb;

  double y1 = (t[len - 2 + j] - mean) / std_;
  d = dmin(sqdist(y1, q[len - 1]), sqdist(y0, q[len - 2]));
  d = dmin(d, sqdist(y1, q[len - 2]));
  lb += d;
  if (lb >= bsf) return lb;

  double x2 = (t[j + 2] - mean) / std_;
  d = dmin(sqdist(x0, q[2]), sqdist(x1, q[2]));
  d = dmin(d, sqdist(x2, q[2]));
  d = dmin(d, sqdist(x2, q[1]));
  d = dmin(d, sqdist(x2, q[0]));
  lb += d;
  if (lb >= bsf) return lb;

  double y2 = (t[len - 3 + j] - mean) / std_;
  d = dmin(sqdist(y0, q[len - 3]), sqdist(y1, q[len - 3]));
  d = dmin(d, sqdist(y2, q[len - 3]));
  d = dmin(d, sqdist(y2, q[len - 2]));
  d = dmin(d, sqdist(y2, q[len - 1]));
  lb += d;
  return lb;
}

// ---- K/utils/DtwUtils.java:206-222 ----
double lb_keogh_cumulative(const int* order, const double* t, const double* uo, const double* lo, double* cb, int j,
                           int len, double mean, double std_, double bsf, int64_t* terms) {
  double lb = 0;
  for (int i = 0; i < len && lb < bsf; i++) {
    double x = (t[order[i] + j] - mean) / std_;
    double d = 0;
    if (x > uo[i]) {
      d = sqdist(x, uo[i]);
    } else if (x < lo[i]) {
      d = sqdist(x, lo[i]);
    }
    lb += d;
    cb[order[i]] = d;
    if (terms) ++*terms;
  }
  return lb;
}

// ---- K/utils/DtwUtils.java:238-257 ----
double lb_keogh_data_cumulative(const int* order, const double* qo, double* cb, int I, const double* l,
                                const double* u, int len, double mean, double std_, double bsf, int64_t* terms) {
  double lb = 0;
  for (int i = 0; i < len && lb < bsf; i++) {
    double uu = (u[order[i] + I] - mean) / std_;
    double ll = (l[order[i] + I] - mean) / std_;
    double d = 0;
    if (qo[i] > uu) {
      d = sqdist(qo[i], uu);
    } else if (qo[i] < ll) {
      d = sqdist(qo[i], ll);
    }
    lb += d;
    cb[order[i]] = d;
    if (terms) ++*terms;
  }
  return lb;
}

// ---- K/utils/DtwUtils.java:269-337 ----
double banded_dtw(const double* A, const double* B, const double* cb, int m, int r, double bsf, int64_t* cells) {
  std::vector<double> row_a(2 * r + 1, kInf), row_b(2 * r + 1, kInf);
  double* cost = row_a.data();
  double* cost_prev = row_b.data();
  int k = 0;
  for (int i = 0; i < m; i++) {
    k = std::max(0, r - i);
    double min_cost = kInf;
    for (int j = std::max(0, i - r); j <= std::min(m - 1, i + r); j++, k++) {
      if (i == 0 && j == 0) {
        cost[k] = sqdist(A[0], B[0]);
        min_cost = cost[k];
        if (cells) ++*cells;
        continue;
      }
      double x, y, z;
      y = (j - 1 < 0 || k - 1 < 0) ? kInf : cost[k - 1];
      x = (i - 1 < 0 || k + 1 > 2 * r) ? kInf : cost_prev[k + 1];
      z = (i - 1 < 0 || j - 1 < 0) ? kInf : cost_prev[k];
      cost[k] = dmin(dmin(x, y), z) + sqdist(A[i], B[j]);
      if (cost[k] < min_cost) min_cost = cost[k];
      if (cells) ++*cells;
    }
    if (i + r < m - 1 && min_cost + cb[i + r + 1] >= bsf) return min_cost + cb[i + r + 1];
    std::swap(cost, cost_prev);
  }
  k--;
  return cost_prev[k];
}

// ---- K/utils/MeanIntervalUtils.java:51-61 ----
double to_round(double value) {
  value *= 10.0;  // Math.pow(10, posOfD - 1), posOfD == 2
  double int_value = std::floor(value);
  double diff = value - int_value;
  double ret = int_value;
  if (java_double_compare(diff, 0.5) >= 0) ret += 0.5;
  ret *= 0.1;     // Math.pow(10, -posOfD + 1)
  return ret;
}

// The data feed.  The engines call TimeSeriesOperator.readTimeSeries(left, length)
// (K/operator/file/TimeSeriesFileOperator.java:54-96): 1-based, throws if out of range.
struct SeriesView {
  const double* d;
  int64_t n;
  bool read(int64_t left, int64_t length, const double** out) const {
    if (left < 1 || left + length - 1 > n || length < 1) return false;  // :55-57 IllegalArgumentException
    *out = d + (left - 1);
    return true;
  }
};

// The streaming feed used by IndexBuilder and the UCR executors: TimeSeriesNodeIterator hands out
// nodes of ROW_LENGTH *bytes* (125 doubles, the last one zero padded), and nextData() only bumps
// its counter when it stays inside a node (K/IndexBuilder.java:152-180,
// K/operator/file/TimeSeriesNodeIterator.java:41-79).
class BlockFeed {
 public:
  BlockFeed(const double* data, int64_t n_file, int64_t limit) : data_(data), n_file_(n_file), limit_(limit) {}
  bool next() {
    if (index_ + 1 < node_len_) {
      ++index_;
      return ++cnt_ <= limit_;
    }
    if (pos_ >= n_file_ * 8) return false;
    int64_t got = std::min<int64_t>(kRowBytes, n_file_ * 8 - pos_);
    int64_t first = pos_ / 8;
    node_len_ = kRowBytes / 8;
    for (int i = 0; i < node_len_; i++) node_[i] = (i < got / 8) ? data_[first + i] : 0.0;
    pos_ += got;
    index_ = 0;
    return true;
  }
  double cur() const { return node_[index_]; }

 private:
  const double* data_;
  int64_t n_file_, limit_;
  int64_t pos_ = 0, cnt_ = 0;
  int index_ = 0, node_len_ = 0;
  double node_[kRowBytes / 8] = {0};
};

struct Answers {
  std::vector<int32_t> off;
  std::vector<double> dist;
};

}  // namespace

extern "C" {

struct kvo_result {
  int64_t count;
  int32_t* offsets;      // 1-based window starts, scan order (0-based for kvo_ucr_dtw, as the reference)
  double* distances;     // sqrt(dist^2)
  int64_t cnt_candidate; // sum(right-left+1), unclamped — the reference's #candidates
  int64_t n_verified;    // window starts actually examined
  int64_t s_total;       // series samples touched (clamped), each counted once per interval
  int64_t n_gate_pass;   // cNSM alpha/beta gate passes
  int64_t n_kim_pass;    // candidates surviving LB_Kim
  int64_t n_keogh_pass;  // candidates surviving LB_Keogh (query envelope)
  int64_t n_dtw;         // candidates reaching dtw()
  int64_t terms;         // distance / lower-bound terms evaluated
  int64_t dtw_cells;     // DTW cells evaluated
};

struct kvo_runs {
  int64_t count;
  double* keys;
  int32_t* first;
  int32_t* last;
};

enum { KVO_OK = 0, KVO_E_ARG = -1, KVO_E_REF_THROWS = -2, KVO_E_IO = -3 };

static void fill_result(kvo_result* r, const Answers& a) {
  r->count = (int64_t)a.off.size();
  r->offsets = (int32_t*)std::malloc(sizeof(int32_t) * std::max<size_t>(1, a.off.size()));
  r->distances = (double*)std::malloc(sizeof(double) * std::max<size_t>(1, a.dist.size()));
  if (!a.off.empty()) {
    std::memcpy(r->offsets, a.off.data(), sizeof(int32_t) * a.off.size());
    std::memcpy(r->distances, a.dist.data(), sizeof(double) * a.dist.size());
  }
}

void kvo_result_free(kvo_result* r) {
  if (!r) return;
  std::free(r->offsets);
  std::free(r->distances);
  r->offsets = nullptr;
  r->distances = nullptr;
  r->count = 0;
}

void kvo_runs_free(kvo_runs* r) {
  if (!r) return;
  std::free(r->keys);
  std::free(r->first);
  std::free(r->last);
  r->keys = nullptr;
  r->first = r->last = nullptr;
  r->count = 0;
}

double kvo_to_round(double v) { return to_round(v); }

int kvo_lower_upper_lemire(const double* t, int len, int r, double* l, double* u) {
  return lemire_envelope(t, len, r, l, u) ? KVO_OK : KVO_E_REF_THROWS;
}

double kvo_lb_kim(const double* t, const double* q, int j, int len, double mean, double std_, double bsf) {
  return lb_kim_hierarchy(t, q, j, len, mean, std_, bsf);
}

double kvo_lb_keogh(const int* order, const double* t, const double* uo, const double* lo, double* cb, int j, int len,
                    double mean, double std_, double bsf) {
  return lb_keogh_cumulative(order, t, uo, lo, cb, j, len, mean, std_, bsf, nullptr);
}

double kvo_lb_keogh_data(const int* order, const double* qo, double* cb, int I, const double* l, const double* u,
                         int len, double mean, double std_, double bsf) {
  return lb_keogh_data_cumulative(order, qo, cb, I, l, u, len, mean, std_, bsf, nullptr);
}

double kvo_dtw(const double* A, const double* B, const double* cb, int m, int r, double bsf) {
  return banded_dtw(A, B, cb, m, r, bsf, nullptr);
}

// Query statistics, K/NormQueryEngine.java:192-198 (sequential sums).
void kvo_query_stats(const double* q, int m, double* meanQ, double* stdQ) {
  double ex = 0, ex2 = 0;
  for (int i = 0; i < m; i++) {
    ex += q[i];
    ex2 += q[i] * q[i];
  }
  *meanQ = ex / m;
  *stdQ = std::sqrt(ex2 / m - *meanQ * *meanQ);
}

// z-normalised query sorted by |z| descending, stable: K/NormQueryEngine.java:438-452.
void kvo_sorted_query(const double* q, int m, double* zq_sorted, int32_t* order) {
  double meanQ, stdQ;
  kvo_query_stats(q, m, &meanQ, &stdQ);
  std::vector<double> z(m);
  for (int i = 0; i < m; i++) z[i] = (q[i] - meanQ) / stdQ;
  std::vector<int32_t> idx(m);
  for (int i = 0; i < m; i++) idx[i] = i;
  std::stable_sort(idx.begin(), idx.end(), [&](int32_t a, int32_t b) {
    // comparator(o1,o2) = Double.compare(|o2|, |o1|) ; o1 sorts first when that is < 0
    return java_double_compare(std::fabs(z[b]), std::fabs(z[a])) < 0;
  });
  for (int i = 0; i < m; i++) {
    zq_sorted[i] = z[idx[i]];
    order[i] = idx[i];
  }
}

// ---------------------------------------------------------------------------------------------
// a1  RSM-ED phase 2 — K/QueryEngine.java:341-363
// ---------------------------------------------------------------------------------------------
int kvo_verify_ed(const double* series, int64_t n, const double* q, int m, double epsilon, const int32_t* lr, int K,
                  int shift, kvo_result* out) {
  if (!series || !q || !out || m < 1 || K < 0 || (K > 0 && !lr)) return KVO_E_ARG;
  std::memset(out, 0, sizeof(*out));
  SeriesView ts{series, n};
  Answers ans;
  const double eps2 = epsilon * epsilon;
  for (int p = 0; p < K; p++) {
    int64_t left = lr[2 * p], right = lr[2 * p + 1];
    out->cnt_candidate += right - left + 1;
    int64_t begin = left - shift;
    int64_t end = right - shift + m - 1;
    if (begin < 1) begin = 1;
    if (end > n) end = n;
    const double* data;
    int64_t size = end - begin + 1;
    if (!ts.read(begin, size, &data)) return KVO_E_REF_THROWS;
    out->s_total += size;
    for (int64_t i = 0; i + m - 1 < size; i++) {
      double distance = 0;
      for (int j = 0; j < m && distance <= eps2; j++) {
        distance += (data[i + j] - q[j]) * (data[i + j] - q[j]);
        ++out->terms;
      }
      ++out->n_verified;
      if (distance <= eps2) {
        ans.off.push_back((int32_t)(begin + i));
        ans.dist.push_back(std::sqrt(distance));
      }
    }
  }
  fill_result(out, ans);
  return KVO_OK;
}

// Grouping of merged intervals into reads of <= MAX_SCAN_DATA_LENGTH points,
// K/NormQueryEngine.java:454-482 (identical in K/NormQueryEngineDtw.java:490-518).
struct ReadGroup {
  int begin_idx, end_idx;
  int64_t begin, end;
};

static void next_group(const int32_t* lr, int K, int shift, int m, int64_t n, int* idx_io, ReadGroup* g) {
  int idx = *idx_io;
  int begin_idx = idx, end_idx = idx;
  int64_t begin = (int64_t)lr[2 * idx] - shift;
  int64_t end = (int64_t)lr[2 * idx + 1] - shift + m - 1;
  if (begin < 1) begin = 1;
  int64_t length = end - begin + 1;
  idx++;
  while (idx < K) {
    begin = (int64_t)lr[2 * idx] - shift;
    int64_t new_length = length + begin - end - 1;
    end = (int64_t)lr[2 * idx + 1] - shift + m - 1;
    if (end > n) end = n;
    new_length += end - begin + 1;
    if (new_length > kMaxScanDataLength) break;
    end_idx = idx;
    length = new_length;
    idx++;
  }
  begin = (int64_t)lr[2 * begin_idx] - shift;
  end = (int64_t)lr[2 * end_idx + 1] - shift + m - 1;
  if (begin < 1) begin = 1;
  if (end > n) end = n;
  g->begin_idx = begin_idx;
  g->end_idx = end_idx;
  g->begin = begin;
  g->end = end;
  *idx_io = idx;
}

// ---------------------------------------------------------------------------------------------
// a2  cNSM-ED phase 2 — K/NormQueryEngine.java:192-198 (query stats) and :432-528
// ---------------------------------------------------------------------------------------------
int kvo_verify_cnsm_ed(const double* series, int64_t n, const double* q, int m, double epsilon, double alpha,
                       double beta, const int32_t* lr, int K, int shift, kvo_result* out) {
  if (!series || !q || !out || m < 1 || K < 0 || (K > 0 && !lr)) return KVO_E_ARG;
  std::memset(out, 0, sizeof(*out));
  SeriesView ts{series, n};
  Answers ans;
  const double eps2 = epsilon * epsilon;

  double meanQ, stdQ;
  kvo_query_stats(q, m, &meanQ, &stdQ);
  std::vector<double> zQ(m);
  std::vector<int32_t> order(m);
  kvo_sorted_query(q, m, zQ.data(), order.data());

  std::vector<double> T(2 * (size_t)m);
  int idx = 0;
  while (idx < K) {
    ReadGroup g;
    next_group(lr, K, shift, m, n, &idx, &g);
    const double* data;
    int64_t size = g.end - g.begin + 1;
    if (!ts.read(g.begin, size, &data)) return KVO_E_REF_THROWS;
    for (int idx1 = g.begin_idx; idx1 <= g.end_idx; idx1++) {
      out->cnt_candidate += (int64_t)lr[2 * idx1 + 1] - lr[2 * idx1] + 1;
      double ex = 0, ex2 = 0;
      std::fill(T.begin(), T.end(), 0.0);
      int64_t begin1 = (int64_t)lr[2 * idx1] - shift - g.begin;
      int64_t end1 = (int64_t)lr[2 * idx1 + 1] - shift + m - 1 - g.begin;
      if (begin1 < 0) begin1 = 0;
      if (end1 > size - 1) end1 = size - 1;
      if (end1 >= begin1) out->s_total += end1 - begin1 + 1;
      for (int64_t i = begin1; i <= end1; i++) {
        double d = data[i];
        ex += d;
        ex2 += d * d;
        T[i % m] = d;
        T[(i % m) + m] = d;
        if (i - begin1 >= m - 1) {
          int j = (int)((i + 1) % m);
          double mean = ex / m;
          double std_ = std::sqrt(ex2 / m - mean * mean);
          ++out->n_verified;
          if (std::fabs(mean - meanQ) <= beta && (std_ / stdQ) <= alpha && (std_ / stdQ) >= 1.0 / alpha) {
            ++out->n_gate_pass;
            double dist = 0;
            for (int k = 0; k < m && dist <= eps2; k++) {
              double x = (T[order[k] + j] - mean) / std_;
              dist += (x - zQ[k]) * (x - zQ[k]);
              ++out->terms;
            }
            if (dist <= eps2) {
              ans.off.push_back((int32_t)(g.begin + i - m + 1));
              ans.dist.push_back(std::sqrt(dist));
            }
          }
          ex -= T[j];
          ex2 -= T[j] * T[j];
        }
      }
    }
  }
  fill_result(out, ans);
  return KVO_OK;
}

// The LB cascade + DTW for one candidate; shared by a3/a4 (K/QueryEngineDtw.java:411-446,
// K/NormQueryEngineDtw.java:555-592).  cb1/cb2 are freshly zeroed per candidate as in the engines.
struct DtwScratch {
  std::vector<double> cb1, cb2, cb, win;
  explicit DtwScratch(int m) : cb1(m), cb2(m), cb(m), win(m) {}
};

static bool cascade_and_dtw(const double* T, const double* Q, const int32_t* order, const double* qo, const double* uo,
                            const double* lo, const double* lBuff, const double* uBuff, int64_t I, int j, int m,
                            int rho, double mean, double std_, double eps2, bool normalise, DtwScratch& s,
                            kvo_result* out, double* dist_out) {
  double lbKim = lb_kim_hierarchy(T, Q, j, m, mean, std_, eps2);
  if (!(lbKim <= eps2)) return false;
  ++out->n_kim_pass;
  std::fill(s.cb1.begin(), s.cb1.end(), 0.0);
  double lbK = lb_keogh_cumulative(order, T, uo, lo, s.cb1.data(), j, m, mean, std_, eps2, &out->terms);
  if (!(lbK <= eps2)) return false;
  ++out->n_keogh_pass;
  if (normalise) {
    for (int k = 0; k < m; k++) s.win[k] = (T[k + j] - mean) / std_;  // NormQueryEngineDtw.java:564-567
  }
  std::fill(s.cb2.begin(), s.cb2.end(), 0.0);
  double lbK2 = lb_keogh_data_cumulative(order, qo, s.cb2.data(), (int)I, lBuff, uBuff, m, mean, std_, eps2, &out->terms);
  if (!(lbK2 <= eps2)) return false;
  if (!normalise) std::memcpy(s.win.data(), T + j, sizeof(double) * m);  // QueryEngineDtw.java:426-427
  const std::vector<double>& src = (lbK > lbK2) ? s.cb1 : s.cb2;
  s.cb[m - 1] = src[m - 1];
  for (int k = m - 2; k >= 0; k--) s.cb[k] = s.cb[k + 1] + src[k];
  ++out->n_dtw;
  double dist = banded_dtw(s.win.data(), Q, s.cb.data(), m, rho, eps2, &out->dtw_cells);
  *dist_out = dist;
  return dist <= eps2;
}

// ---------------------------------------------------------------------------------------------
// a3  RSM-DTW phase 2 — K/QueryEngineDtw.java:349-452  (query order = identity, :368-371)
// ---------------------------------------------------------------------------------------------
int kvo_verify_dtw(const double* series, int64_t n, const double* q, int m, double epsilon, int rho,
                   const int32_t* lr, int K, int shift, kvo_result* out) {
  if (!series || !q || !out || m < 3 || rho < 0 || K < 0 || (K > 0 && !lr)) return KVO_E_ARG;
  std::memset(out, 0, sizeof(*out));
  SeriesView ts{series, n};
  Answers ans;
  const double eps2 = epsilon * epsilon;

  std::vector<double> Q(q, q + m), u(m), l(m);
  if (!lemire_envelope(Q.data(), m, rho, l.data(), u.data())) return KVO_E_REF_THROWS;
  std::vector<int32_t> order(m);
  std::vector<double> qo(m), uo(m), lo(m);
  for (int i = 0; i < m; i++) {
    order[i] = i;
    qo[i] = Q[i];
    uo[i] = u[i];
    lo[i] = l[i];
  }
  DtwScratch scratch(m);
  std::vector<double> T(2 * (size_t)m);
  std::vector<double> uBuff, lBuff;
  for (int p = 0; p < K; p++) {
    int64_t left = lr[2 * p], right = lr[2 * p + 1];
    out->cnt_candidate += right - left + 1;
    int64_t begin = left - shift;
    int64_t end = right - shift + m - 1;
    if (begin < 1) begin = 1;
    if (end > n) end = n;
    const double* data;
    int64_t size = end - begin + 1;
    if (!ts.read(begin, size, &data)) return KVO_E_REF_THROWS;
    out->s_total += size;
    std::fill(T.begin(), T.end(), 0.0);
    uBuff.assign(size, 0.0);
    lBuff.assign(size, 0.0);
    if (!lemire_envelope(data, (int)size, rho, lBuff.data(), uBuff.data())) return KVO_E_REF_THROWS;
    for (int64_t i = 0; i < size; i++) {
      double d = data[i];
      T[i % m] = d;
      T[(i % m) + m] = d;
      if (i >= m - 1) {
        int j = (int)((i + 1) % m);
        ++out->n_verified;
        double dist;
        if (cascade_and_dtw(T.data(), Q.data(), order.data(), qo.data(), uo.data(), lo.data(), lBuff.data(),
                            uBuff.data(), i - m + 1, j, m, rho, 0.0, 1.0, eps2, false, scratch, out, &dist)) {
          ans.off.push_back((int32_t)(begin + i - m + 1));
          ans.dist.push_back(std::sqrt(dist));
        }
      }
    }
  }
  fill_result(out, ans);
  return KVO_OK;
}

// ---------------------------------------------------------------------------------------------
// a4  cNSM-DTW phase 2 — K/NormQueryEngineDtw.java:205-211 and :457-603.
// DOCUMENTED DEVIATION: the reference sorts the query with comparator (int)(|b|-|a|)
// (:475-478), which is not a valid total order; its TimSort output cannot be reproduced
// without a JVM.  `order` only steers where the LB loops stop early — it changes neither
// the answer set nor any reported distance — so identity order is used here.
// ---------------------------------------------------------------------------------------------
int kvo_verify_cnsm_dtw(const double* series, int64_t n, const double* q, int m, double epsilon, int rho, double alpha,
                        double beta, const int32_t* lr, int K, int shift, kvo_result* out) {
  if (!series || !q || !out || m < 3 || rho < 0 || K < 0 || (K > 0 && !lr)) return KVO_E_ARG;
  std::memset(out, 0, sizeof(*out));
  SeriesView ts{series, n};
  Answers ans;
  const double eps2 = epsilon * epsilon;

  double meanQ, stdQ;
  kvo_query_stats(q, m, &meanQ, &stdQ);
  std::vector<double> zQ(m), u(m), l(m);
  for (int i = 0; i < m; i++) zQ[i] = (q[i] - meanQ) / stdQ;
  if (!lemire_envelope(zQ.data(), m, rho, l.data(), u.data())) return KVO_E_REF_THROWS;
  std::vector<int32_t> order(m);
  std::vector<double> qo(m), uo(m), lo(m);
  for (int i = 0; i < m; i++) {
    order[i] = i;
    qo[i] = zQ[i];
    uo[i] = u[i];
    lo[i] = l[i];
  }
  DtwScratch scratch(m);
  std::vector<double> T(2 * (size_t)m);
  std::vector<double> uBuff, lBuff;
  int idx = 0;
  while (idx < K) {
    ReadGroup g;
    next_group(lr, K, shift, m, n, &idx, &g);
    const double* data;
    int64_t size = g.end - g.begin + 1;
    if (!ts.read(g.begin, size, &data)) return KVO_E_REF_THROWS;
    uBuff.assign(size, 0.0);
    lBuff.assign(size, 0.0);
    if (!lemire_envelope(data, (int)size, rho, lBuff.data(), uBuff.data())) return KVO_E_REF_THROWS;
    for (int idx1 = g.begin_idx; idx1 <= g.end_idx; idx1++) {
      out->cnt_candidate += (int64_t)lr[2 * idx1 + 1] - lr[2 * idx1] + 1;
      double ex = 0, ex2 = 0;
      std::fill(T.begin(), T.end(), 0.0);
      int64_t begin1 = (int64_t)lr[2 * idx1] - shift - g.begin;
      int64_t end1 = (int64_t)lr[2 * idx1 + 1] - shift + m - 1 - g.begin;
      if (begin1 < 0) begin1 = 0;
      if (end1 > size - 1) end1 = size - 1;
      if (end1 >= begin1) out->s_total += end1 - begin1 + 1;
      for (int64_t i = begin1; i <= end1; i++) {
        double d = data[i];
        ex += d;
        ex2 += d * d;
        T[i % m] = d;
        T[(i % m) + m] = d;
        if (i - begin1 >= m - 1) {
          int j = (int)((i + 1) % m);
          double mean = ex / m;
          double std_ = std::sqrt(ex2 / m - mean * mean);
          ++out->n_verified;
          if (std::fabs(mean - meanQ) <= beta && (std_ / stdQ) <= alpha && (std_ / stdQ) >= 1.0 / alpha) {
            ++out->n_gate_pass;
            double dist;
            if (cascade_and_dtw(T.data(), zQ.data(), order.data(), qo.data(), uo.data(), lo.data(), lBuff.data(),
                                uBuff.data(), i - m + 1, j, m, rho, mean, std_, eps2, true, scratch, out, &dist)) {
              ans.off.push_back((int32_t)(g.begin + i - m + 1));
              ans.dist.push_back(std::sqrt(dist));
            }
          }
          ex -= T[j];
          ex2 -= T[j] * T[j];
        }
      }
    }
  }
  fill_result(out, ans);
  return KVO_OK;
}

// ---------------------------------------------------------------------------------------------
// a10/a11  IndexBuilder.SingleIndexBuilder.run step 1 — K/IndexBuilder.java:186-301.
// `n_file` = number of doubles in files/data-N (normally == n); the block iterator pads the last
// 1000-byte block with zeros, which this restatement reproduces.
// Output: the (key, first, last) intervals in the order the reference appends them.
// ---------------------------------------------------------------------------------------------
int kvo_window_mean_runs(const double* series, int64_t n_file, int64_t n, int w, kvo_runs* out) {
  if (!series || !out || w < 2 || w > kEpoch || n < 1) return KVO_E_ARG;
  std::memset(out, 0, sizeof(*out));
  BlockFeed feed(series, n_file, n);
  std::vector<double> buffer(kEpoch, 0.0), t(2 * (size_t)w, 0.0);
  std::vector<double> keys;
  std::vector<int32_t> firsts, lasts;
  bool done = false, have_last = false;
  double last_key = 0;
  int it = 0, ep;
  while (!done) {
    if (it == 0) {
      for (int k = 0; k < w - 1; k++) {
        if (feed.next()) buffer[k] = feed.cur();
      }
    } else {
      for (int k = 0; k < w - 1; k++) buffer[k] = buffer[kEpoch - w + 1 + k];
    }
    ep = w - 1;
    while (ep < kEpoch) {
      if (feed.next()) {
        buffer[ep] = feed.cur();
        ep++;
      } else {
        break;
      }
    }
    if (ep <= w - 1) {
      done = true;
    } else {
      double ex = 0, ex2 = 0;
      for (int i = 0; i < ep; i++) {
        double d = buffer[i];
        ex += d;
        ex2 += d * d;
        t[i % w] = d;
        t[(i % w) + w] = d;
        if (i >= w - 1) {
          double mean = ex / w;
          int j = (i + 1) % w;
          int64_t loc = (int64_t)it * (kEpoch - w + 1) + i - w + 1 + 1;
          if (loc > n) {
            done = true;
            break;
          }
          double cur = to_round(mean);
          if (!have_last || double_equals_bits(last_key) != double_equals_bits(cur) ||
              loc - firsts.back() == kMaximumDiff - 1) {
            keys.push_back(cur);
            firsts.push_back((int32_t)loc);
            lasts.push_back((int32_t)loc);
            last_key = cur;
            have_last = true;
          } else {
            lasts.back() = (int32_t)loc;
          }
          ex -= t[j];
          ex2 -= t[j] * t[j];
        }
      }
      if (ep < kEpoch) {
        done = true;
      } else {
        it++;
      }
    }
  }
  out->count = (int64_t)keys.size();
  size_t c = std::max<size_t>(1, keys.size());
  out->keys = (double*)std::malloc(sizeof(double) * c);
  out->first = (int32_t*)std::malloc(sizeof(int32_t) * c);
  out->last = (int32_t*)std::malloc(sizeof(int32_t) * c);
  if (!keys.empty()) {
    std::memcpy(out->keys, keys.data(), sizeof(double) * keys.size());
    std::memcpy(out->first, firsts.data(), sizeof(int32_t) * keys.size());
    std::memcpy(out->last, lasts.data(), sizeof(int32_t) * keys.size());
  }
  return KVO_OK;
}

// ---------------------------------------------------------------------------------------------
// Index-free full scans (BASELINE config 5 semantics).
// cNSM-ED stream, statistics chain NEVER reset: K/experiments/ucr/UcrEdQueryExecutor.java:101-183.
// Offsets are 1-based (:166).
// ---------------------------------------------------------------------------------------------
int kvo_ucr_ed(const double* series, int64_t n_file, int64_t N, const double* q, int M, double epsilon, double alpha,
               double beta, kvo_result* out) {
  if (!series || !q || !out || M < 1) return KVO_E_ARG;
  std::memset(out, 0, sizeof(*out));
  Answers ans;
  const double eps2 = epsilon * epsilon;
  double meanQ, stdQ;
  kvo_query_stats(q, M, &meanQ, &stdQ);
  std::vector<double> Q(M);
  std::vector<int32_t> order(M);
  kvo_sorted_query(q, M, Q.data(), order.data());
  std::vector<double> t(2 * (size_t)M, 0.0);
  BlockFeed feed(series, n_file, N);
  double ex = 0, ex2 = 0;
  int64_t i = 0;
  while (feed.next()) {
    double d = feed.cur();
    ex += d;
    ex2 += d * d;
    t[i % M] = d;
    t[(i % M) + M] = d;
    ++out->s_total;
    if (i >= M - 1) {
      int j = (int)((i + 1) % M);
      double mean = ex / M;
      double std_ = ex2 / M;
      std_ = std::sqrt(std_ - mean * mean);
      ++out->n_verified;
      if (std::fabs(mean - meanQ) <= beta && (std_ / stdQ) <= alpha && (std_ / stdQ) >= 1.0 / alpha) {
        ++out->n_gate_pass;
        double sum = 0;  // distance(): loop condition is `sum < bsf` here (:91), unlike the engine's `<=`
        for (int k = 0; k < M && sum < eps2; k++) {
          double x = (t[order[k] + j] - mean) / std_;
          sum += (x - Q[k]) * (x - Q[k]);
          ++out->terms;
        }
        if (sum <= eps2) {
          ans.off.push_back((int32_t)(i - M + 2));
          ans.dist.push_back(std::sqrt(sum));
        }
      }
      ex -= t[j];
      ex2 -= t[j] * t[j];
    }
    i++;
  }
  fill_result(out, ans);
  return KVO_OK;
}

// cNSM-DTW stream with EPOCH restarts: K/experiments/ucr/UcrDtwQueryExecutor.java:84-314.
// Offsets are 0-based (:278).  cb1/cb2 persist across candidates (:109-110).  Identity query
// order (same deviation as kvo_verify_cnsm_dtw).
int kvo_ucr_dtw(const double* series, int64_t n_file, int64_t N, const double* query, int M, double epsilon, int rho,
                double alpha, double beta, kvo_result* out) {
  if (!series || !query || !out || M < 3 || rho < 0 || M > kEpoch) return KVO_E_ARG;
  std::memset(out, 0, sizeof(*out));
  Answers ans;
  const double eps2 = epsilon * epsilon;
  double meanQ, stdQ;
  kvo_query_stats(query, M, &meanQ, &stdQ);
  std::vector<double> q(M), u(M), l(M), qo(M), uo(M), lo(M), cb(M, 0.0), cb1(M, 0.0), cb2(M, 0.0), tz(M);
  std::vector<double> t(2 * (size_t)M, 0.0), buffer(kEpoch, 0.0), u_buff(kEpoch, 0.0), l_buff(kEpoch, 0.0);
  std::vector<int32_t> order(M);
  for (int i = 0; i < M; i++) q[i] = (query[i] - meanQ) / stdQ;
  if (!lemire_envelope(q.data(), M, rho, l.data(), u.data())) return KVO_E_REF_THROWS;
  for (int i = 0; i < M; i++) {
    order[i] = i;
    qo[i] = q[i];
    uo[i] = u[i];
    lo[i] = l[i];
  }
  BlockFeed feed(series, n_file, N);
  bool done = false;
  int it = 0, ep;
  while (!done) {
    if (it == 0) {
      for (int k = 0; k < M - 1; k++) {
        if (feed.next()) buffer[k] = feed.cur();
      }
    } else {
      for (int k = 0; k < M - 1; k++) buffer[k] = buffer[kEpoch - M + 1 + k];
    }
    ep = M - 1;
    while (ep < kEpoch) {
      if (!feed.next()) break;
      buffer[ep] = feed.cur();
      ep++;
    }
    if (ep <= M - 1) {
      done = true;
    } else {
      if (!lemire_envelope(buffer.data(), ep, rho, l_buff.data(), u_buff.data())) return KVO_E_REF_THROWS;
      out->s_total += (it == 0) ? ep : ep - (M - 1);
      double ex = 0, ex2 = 0;
      for (int i = 0; i < ep; i++) {
        double d = buffer[i];
        ex += d;
        ex2 += d * d;
        t[i % M] = d;
        t[(i % M) + M] = d;
        if (i >= M - 1) {
          double mean = ex / M;
          double std_ = ex2 / M;
          std_ = std::sqrt(std_ - mean * mean);
          int j = (i + 1) % M;
          int I = i - (M - 1);
          ++out->n_verified;
          if (std::fabs(mean - meanQ) <= beta && (std_ / stdQ) <= alpha && (std_ / stdQ) >= 1.0 / alpha) {
            ++out->n_gate_pass;
            double lb_kim = lb_kim_hierarchy(t.data(), q.data(), j, M, mean, std_, eps2);
            if (lb_kim <= eps2) {
              ++out->n_kim_pass;
              double lb_k = lb_keogh_cumulative(order.data(), t.data(), uo.data(), lo.data(), cb1.data(), j, M, mean,
                                                std_, eps2, &out->terms);
              if (lb_k <= eps2) {
                ++out->n_keogh_pass;
                for (int k = 0; k < M; k++) tz[k] = (t[k + j] - mean) / std_;
                double lb_k2 = lb_keogh_data_cumulative(order.data(), qo.data(), cb2.data(), I, l_buff.data(),
                                                        u_buff.data(), M, mean, std_, eps2, &out->terms);
                if (lb_k2 <= eps2) {
                  const std::vector<double>& src = (lb_k > lb_k2) ? cb1 : cb2;
                  cb[M - 1] = src[M - 1];
                  for (int k = M - 2; k >= 0; k--) cb[k] = cb[k + 1] + src[k];
                  ++out->n_dtw;
                  double dist = banded_dtw(tz.data(), q.data(), cb.data(), M, rho, eps2, &out->dtw_cells);
                  if (dist <= eps2) {
                    ans.off.push_back((int32_t)((int64_t)it * (kEpoch - M + 1) + i - M + 1));
                    ans.dist.push_back(std::sqrt(dist));
                  }
                }
              }
            }
          }
          ex -= t[j];
          ex2 -= t[j] * t[j];
        }
      }
      if (ep < kEpoch) {
        done = true;
      } else {
        it++;
      }
    }
  }
  fill_result(out, ans);
  return KVO_OK;
}

// ---------------------------------------------------------------------------------------------
// a12  data file codec: files/data-N holds N big-endian IEEE-754 doubles, no header
// (K/DataGenerator.java:102-113 DataOutputStream.writeDouble; reader
// K/operator/file/TimeSeriesFileOperator.java:54-96).
// ---------------------------------------------------------------------------------------------
int kvo_write_series_be(const char* path, const double* series, int64_t n) {
  FILE* f = std::fopen(path, "wb");
  if (!f) return KVO_E_IO;
  std::vector<unsigned char> buf(8 * 4096);
  int64_t done = 0;
  while (done < n) {
    int64_t c = std::min<int64_t>(4096, n - done);
    for (int64_t i = 0; i < c; i++) {
      uint64_t b;
      std::memcpy(&b, &series[done + i], 8);
      for (int k = 0; k < 8; k++) buf[8 * i + k] = (unsigned char)(b >> (56 - 8 * k));
    }
    if (std::fwrite(buf.data(), 8, (size_t)c, f) != (size_t)c) {
      std::fclose(f);
      return KVO_E_IO;
    }
    done += c;
  }
  std::fclose(f);
  return KVO_OK;
}

int kvo_read_series_be(const char* path, double* series, int64_t n) {
  FILE* f = std::fopen(path, "rb");
  if (!f) return KVO_E_IO;
  std::vector<unsigned char> buf(8 * 4096);
  int64_t done = 0;
  while (done < n) {
    int64_t c = std::min<int64_t>(4096, n - done);
    if (std::fread(buf.data(), 8, (size_t)c, f) != (size_t)c) {
      std::fclose(f);
      return KVO_E_IO;
    }
    for (int64_t i = 0; i < c; i++) {
      uint64_t b = 0;
      for (int k = 0; k < 8; k++) b = (b << 8) | buf[8 * i + k];
      std::memcpy(&series[done + i], &b, 8);
    }
    done += c;
  }
  std::fclose(f);
  return KVO_OK;
}


// ---------------------------------------------------------------------------------------------
// IndexBuilder step 2 and the index file image.
//   step 2            K/IndexBuilder.java:308-345 (row merge criterion :325-329)
//   mergeIndexNode    K/utils/IndexNodeUtils.java:30-90 (incl. the copied, mutated `last` pairs)
//   toBytesCompact    K/common/entity/IndexNode.java:51-96
//   file layout       K/operator/file/IndexFileOperator.java:127-164, K/utils/ByteUtils.java:84-100
// Returns the bytes IndexFileOperator.writeAll would put into files/index-N-w.
// ---------------------------------------------------------------------------------------------
namespace {
typedef std::pair<int32_t, int32_t> IvPair;
typedef std::vector<IvPair> IndexNodePositions;

void add_interval(IndexNodePositions& node, IvPair position) {  // IndexNodeUtils.java:82-90
  while (position.second - position.first >= kMaximumDiff) {
    int32_t new_first = position.first + kMaximumDiff - 1;
    node.push_back(IvPair(position.first, new_first));
    position.first = new_first + 1;
  }
  node.push_back(position);
}

IndexNodePositions merge_index_node(const IndexNodePositions& node1, const IndexNodePositions& node2) {  // :30-80
  IndexNodePositions ret;
  size_t index1 = 0, index2 = 0;
  bool has1 = false, has2 = false;  // `last1 == null` / `last2 == null`
  IvPair last1, last2;
  while (index1 < node1.size() && index2 < node2.size()) {
    if (!has1) { last1 = node1[index1]; has1 = true; }
    if (!has2) { last2 = node2[index2]; has2 = true; }
    if (last1.second + 1 < last2.first) {
      add_interval(ret, last1);
      index1++;
      has1 = false;
    } else if (last2.second + 1 < last1.first) {
      add_interval(ret, last2);
      index2++;
      has2 = false;
    } else {
      if (last1.second < last2.second) {
        if (last1.first < last2.first) last2.first = last1.first;
        index1++;
        has1 = false;
      } else {
        if (last2.first < last1.first) last1.first = last2.first;
        index2++;
        has2 = false;
      }
    }
  }
  for (size_t i = index1; i < node1.size(); i++) {
    if (!has1) { last1 = node1[i]; has1 = true; }
    add_interval(ret, last1);
    has1 = false;
  }
  for (size_t i = index2; i < node2.size(); i++) {
    if (!has2) { last2 = node2[i]; has2 = true; }
    add_interval(ret, last2);
    has2 = false;
  }
  return ret;
}

void put_be32(std::vector<unsigned char>& b, size_t at, int32_t v) {  // hbase Bytes.toBytes(int): big-endian
  for (int k = 0; k < 4; k++) b[at + k] = (unsigned char)(((uint32_t)v) >> (24 - 8 * k));
}
void push_be32(std::vector<unsigned char>& b, int32_t v) {
  b.resize(b.size() + 4);
  put_be32(b, b.size() - 4, v);
}
void push_be_double(std::vector<unsigned char>& b, double d) {  // Bytes.toBytes(double): big-endian raw long bits
  uint64_t u;
  std::memcpy(&u, &d, 8);
  for (int k = 0; k < 8; k++) b.push_back((unsigned char)(u >> (56 - 8 * k)));
}

std::vector<unsigned char> to_bytes_compact(const IndexNodePositions& positions) {  // IndexNode.java:51-96
  std::vector<unsigned char> result(4 * positions.size() * 2);
  size_t index = 0, length = 0;
  int count = 0;
  bool is_packing = false;
  while (index < positions.size()) {
    if (!is_packing) {
      put_be32(result, length, positions[index].first);
      length += 4 + 1;  // first: 4 bytes, remain 1 byte for count
      int diff = positions[index].second - positions[index].first;
      result[length++] = (unsigned char)(signed char)(diff - 128);
      is_packing = true;
      count = 1;
    } else {
      int diff = positions[index].first - positions[index - 1].second;
      if (diff < kMaximumDiff && (count - 1) / 2 + 2 < kMaximumDiff) {
        result[length++] = (unsigned char)(signed char)(diff - 128);
        diff = positions[index].second - positions[index].first;
        result[length++] = (unsigned char)(signed char)(diff - 128);
        count += 2;
      } else {
        result[length - count - 1] = (unsigned char)(signed char)((count - 1) / 2 - 128);
        is_packing = false;
        continue;
      }
    }
    index++;
  }
  if (is_packing) result[length - count - 1] = (unsigned char)(signed char)((count - 1) / 2 - 128);
  result.resize(length);
  return result;
}

struct JavaDoubleLess {  // Double.compareTo: numeric, -0.0 < 0.0, NaN last
  bool operator()(double a, double b) const { return java_double_compare(a, b) < 0; }
};
}  // namespace

struct kvo_bytes {
  int64_t count;
  unsigned char* data;
  int32_t n_rows_step1, n_rows;  // rows before / after the step-2 merge
};

int kvo_index_file_image(const double* series, int64_t n_file, int64_t n, int w, kvo_bytes* out) {
  if (!out) return KVO_E_ARG;
  std::memset(out, 0, sizeof(*out));
  kvo_runs runs;
  int rc = kvo_window_mean_runs(series, n_file, n, w, &runs);
  if (rc) return rc;
  // step 1's HashMap<Double, IndexNode> (K/IndexBuilder.java:268-306): positions appended per key in scan order
  std::map<double, IndexNodePositions, JavaDoubleLess> index_node_map;
  for (int64_t i = 0; i < runs.count; i++) index_node_map[runs.keys[i]].push_back(IvPair(runs.first[i], runs.last[i]));
  kvo_runs_free(&runs);
  out->n_rows_step1 = (int32_t)index_node_map.size();
  if (index_node_map.empty()) return KVO_E_REF_THROWS;  // rawStatisticInfo.get(0) throws
  // step 2 (:308-345)
  std::vector<std::pair<double, int>> raw;  // (key, #intervals), sorted by key descending (:317)
  double sum = 0;
  for (auto& e : index_node_map) {
    raw.push_back(std::make_pair(e.first, (int)e.second.size()));
    sum += (double)e.second.size();
  }
  std::stable_sort(raw.begin(), raw.end(), [](const std::pair<double, int>& a, const std::pair<double, int>& b) {
    return java_double_compare(a.first, b.first) > 0;
  });
  const double average = sum / (double)raw.size();  // StatisticInfo.getAverage
  std::map<double, IndexNodePositions, JavaDoubleLess> index_store;          // TreeMap
  std::vector<std::pair<double, std::pair<int32_t, int32_t>>> statistic_info;  // (key, (#intervals, #offsets))
  auto stat_pair = [](const IndexNodePositions& p) {
    int32_t offs = 0;
    for (auto& iv : p) offs += iv.second - iv.first + 1;
    return std::make_pair((int32_t)p.size(), offs);
  };
  IndexNodePositions last = index_node_map[raw[0].first];
  for (size_t i = 1; i < raw.size(); i++) {
    const IndexNodePositions& current = index_node_map[raw[i].first];
    bool is_merged = false;
    if ((double)raw[i].second < average * 1.2) {
      IndexNodePositions merged = merge_index_node(last, current);
      if ((double)merged.size() < (double)(last.size() + current.size()) * 0.8) {
        last = merged;
        is_merged = true;
      }
    }
    if (!is_merged) {
      double key = raw[i - 1].first;
      index_store[key] = last;
      statistic_info.push_back(std::make_pair(key, stat_pair(last)));
      last = current;
    }
  }
  {
    double key = raw[raw.size() - 1].first;
    index_store[key] = last;
    statistic_info.push_back(std::make_pair(key, stat_pair(last)));
  }
  out->n_rows = (int32_t)index_store.size();
  // IndexFileOperator.writeAll (:127-164)
  std::vector<unsigned char> file;
  std::vector<int32_t> offsets;
  for (auto& e : index_store) {
    offsets.push_back((int32_t)file.size());
    push_be_double(file, e.first);
    std::vector<unsigned char> v = to_bytes_compact(e.second);
    file.insert(file.end(), v.begin(), v.end());
  }
  offsets.push_back((int32_t)file.size());
  std::stable_sort(statistic_info.begin(), statistic_info.end(),
                   [](const std::pair<double, std::pair<int32_t, int32_t>>& a,
                      const std::pair<double, std::pair<int32_t, int32_t>>& b) { return a.first < b.first; });  // comparingDouble
  for (size_t i = 0; i < statistic_info.size(); i++) {  // ByteUtils.listTripleToByteArray: cumulative counts
    if (i > 0) {
      statistic_info[i].second.first += statistic_info[i - 1].second.first;
      statistic_info[i].second.second += statistic_info[i - 1].second.second;
    }
    push_be_double(file, statistic_info[i].first);
    push_be32(file, statistic_info[i].second.first);
    push_be32(file, statistic_info[i].second.second);
  }
  offsets.push_back((int32_t)file.size());
  for (int32_t o : offsets) push_be32(file, o);
  out->count = (int64_t)file.size();
  out->data = (unsigned char*)std::malloc(file.size() ? file.size() : 1);
  std::memcpy(out->data, file.data(), file.size());
  return KVO_OK;
}

void kvo_bytes_free(kvo_bytes* b) {
  if (!b) return;
  std::free(b->data);
  b->data = nullptr;
  b->count = 0;
}

}  // extern "C"
