// phase1.hpp — the tail of phase 1 on the host: the interval algebra between index probing and verification
// (K/QueryEngine.java:266-308, 593-693).  Host code on purpose: the lists are thousands of intervals per query segment
// (KBs), far below what a kernel launch costs; what matters is that the candidate list reaches kvm_verify_* without a
// round trip through boxed Java lists.  Plain arrays in, plain arrays out; no GPU, no ctx.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <numeric>
#include <vector>

namespace kvm_phase1 {

struct Iv {
  int32_t left, right;
  double eps;
};

// intervals.sort(Comparator.comparingInt(Interval::getLeft)): List.sort is a stable merge sort
inline void stable_sort_by_left(std::vector<Iv>& v) {
  std::stable_sort(v.begin(), v.end(), [](const Iv& a, const Iv& b) { return a.left < b.left; });
}

// mode 0: sortButNotMergeIntervals (:593-622)          merge only overlaps, or neighbours whose epsilons differ by < 1
// mode 1: sortButNotMergeIntervalsAndCount (:624-662)  the same + the two counts of the phase-2 time estimate
// mode 2: sortAndMergeIntervals (:664-693)             merge overlaps and neighbours (left - 1 <= end)
inline void sort_merge(std::vector<Iv>& v, int mode, std::vector<Iv>& out, int64_t* cnt_disjoint, int64_t* cnt_offsets) {
  out.clear();
  if (cnt_disjoint) *cnt_disjoint = (int64_t)v.size();
  if (cnt_offsets) *cnt_offsets = v.empty() ? 0 : (int64_t)v[0].right - v[0].left + 1;
  if (v.size() <= 1) {  // returned as is (:594-596, :625-627, :665-667)
    out = v;
    return;
  }
  stable_sort_by_left(v);
  int32_t start = v[0].left, end = v[0].right;
  double eps = v[0].eps;
  int64_t disjoint = (int64_t)v.size(), offsets = 0;
  for (size_t i = 1; i < v.size(); i++) {
    const Iv& c = v[i];
    if ((int64_t)c.left - 1 <= end) disjoint--;
    const bool merge = (mode == 2) ? ((int64_t)c.left - 1 <= end)
                                   : ((int64_t)c.left - 1 < end || ((int64_t)c.left - 1 == end && std::fabs(c.eps - eps) < 1));
    if (merge) {
      end = std::max(c.right, end);
      eps = std::min(c.eps, eps);
    } else {
      out.push_back(Iv{start, end, eps});
      offsets += (int64_t)end - start + 1;
      start = c.left;
      end = c.right;
      eps = c.eps;
    }
  }
  out.push_back(Iv{start, end, eps});
  offsets += (int64_t)end - start + 1;
  if (cnt_disjoint) *cnt_disjoint = disjoint;
  if (cnt_offsets) *cnt_offsets = offsets;
}

// CS ∩ CS_i (:282-308): both lists sorted by left and internally disjoint; keeps the overlaps whose summed lower bound
// stays within eps2, shifted by delta_w; returns the smallest summed bound kept (Double.MAX_VALUE if none).
inline double intersect(const std::vector<Iv>& cs, const std::vector<Iv>& csi, double eps2, int32_t delta_w, std::vector<Iv>& out) {
  out.clear();
  double min_eps = 1.7976931348623157e308;
  size_t i1 = 0, i2 = 0;
  while (i1 < cs.size() && i2 < csi.size()) {
    if (cs[i1].right < csi[i2].left) {
      i1++;
    } else if (csi[i2].right < cs[i1].left) {
      i2++;
    } else {
      const double sum = cs[i1].eps + csi[i2].eps;
      const int32_t l = std::max(cs[i1].left, csi[i2].left) + delta_w;
      if (cs[i1].right < csi[i2].right) {
        if (sum <= eps2) {
          out.push_back(Iv{l, cs[i1].right + delta_w, sum});
          if (sum < min_eps) min_eps = sum;
        }
        i1++;
      } else {
        if (sum <= eps2) {
          out.push_back(Iv{l, csi[i2].right + delta_w, sum});
          if (sum < min_eps) min_eps = sum;
        }
        i2++;
      }
    }
  }
  return min_eps;
}

// The first segment's positions clamped to window starts inside the series (:264-280).
inline double first_segment(const std::vector<Iv>& pos, int32_t order, int32_t w0, int32_t length, int32_t n, int32_t delta_w,
                            std::vector<Iv>& out) {
  out.clear();
  double min_eps = 1.7976931348623157e308;
  const int64_t sh = (int64_t)(order - 1) * w0;
  for (const Iv& p : pos) {
    if ((int64_t)p.right - sh + length - 1 > n) {
      if ((int64_t)p.left - sh + length - 1 <= n) out.push_back(Iv{p.left + delta_w, (int32_t)(n - length + 1 + sh + delta_w), p.eps});
    } else if ((int64_t)p.left - sh < 1) {
      if ((int64_t)p.right - sh >= 1) out.push_back(Iv{(int32_t)(1 + sh + delta_w), p.right + delta_w, p.eps});
    } else {
      out.push_back(Iv{p.left + delta_w, p.right + delta_w, p.eps});
    }
    if (p.eps < min_eps) min_eps = p.eps;
  }
  return min_eps;
}

// ---------------------------------------------------------------------------------------------------------------
// cNSM variants (K/NormQueryEngine.java:313-397, 788-896; K/NormQueryEngineDtw.java:326-425, 926-1046): an interval carries
// the lower (and, for DTW, upper) sums of the segments seen so far, in blocks of w0 points, and the bit set of beta
// partitions its rows fell into (K/common/NormInterval.java).  Layout = kvm_norm_interval.
struct NormIv {
  int32_t left, right;
  double ex, ex2;    // exLower, ex2Lower
  double exu, ex2u;  // exUpper, ex2Upper (DTW engine; zero in the ED engine)
  int64_t bp;
};

// Double.compare(a, b) == 0: same value with the same sign of zero, or both NaN
inline bool jcompare_eq(double a, double b) {
  if (a != a || b != b) return a != a && b != b;
  return a == b && std::signbit(a) == std::signbit(b);
}
// Double.compare(a, b) <= 0 (NaN sorts above everything, -0.0 below +0.0)
inline bool jcompare_le(double a, double b) {
  if (a != a) return b != b;
  if (b != b) return true;
  if (a < b) return true;
  if (a > b) return false;
  return !(std::signbit(b) && !std::signbit(a));
}

// Math.min(double, double): NaN if either is NaN, -0.0 below +0.0
inline double jmin(double a, double b) {
  if (a != a) return a;
  if (b != b) return b;
  if (a == 0.0 && b == 0.0) return std::signbit(a) ? a : b;
  return a <= b ? a : b;
}

// mode 0: sortButNotMergeIntervals (:788-823 / Dtw :926-967)          merge overlaps, or neighbours with bit-equal LOWER sums
// mode 1: sortButNotMergeIntervalsAndCount (:825-869 / Dtw :969-1019)  the same + the two counts of the phase-2 time estimate
// mode 2: sortAndMergeIntervals (:871-896 / Dtw :1021-1046)            merge overlaps and neighbours; sums and partitions dropped
// (a merged interval keeps the MINIMUM of every sum, the upper ones included: Dtw :949-950)
// `emit(const NormIv&)` receives the result in order; the input is read in place (no copy): already sorted lists (every
// intersection output) are walked directly, the others through a permutation from a stable LSD radix sort of
// (left, position) pairs — the lists are millions of 48-byte intervals at n = 1e8 and this is where T_1 goes.
template <class Emit>
inline void norm_sort_merge_core(const NormIv* v, size_t n, int mode, Emit&& emit_out, int64_t* cnt_disjoint, int64_t* cnt_offsets) {
  if (cnt_disjoint) *cnt_disjoint = (int64_t)n;
  if (cnt_offsets) *cnt_offsets = n == 0 ? 0 : (int64_t)v[0].right - v[0].left + 1;
  if (n <= 1) {  // returned as is
    if (n == 1) emit_out(v[0]);
    return;
  }
  // List.sort(comparingInt(getLeft)) is a stable merge sort
  bool sorted = true;
  for (size_t i = 1; i < n && sorted; i++) sorted = v[i - 1].left <= v[i].left;
  std::vector<uint64_t> key;
  if (!sorted) {
    key.resize(n);
    std::vector<uint64_t> tmp(n);
    for (size_t i = 0; i < n; i++) key[i] = ((uint64_t)((uint32_t)v[i].left ^ 0x80000000u) << 32) | (uint64_t)i;
    for (int pass = 0; pass < 3; pass++) {  // bits 32..42, 43..53, 54..63 of the key: stable, so equal lefts keep their order
      const int shift = 32 + 11 * pass, bits = pass == 2 ? 10 : 11;
      const uint64_t mask = ((uint64_t)1 << bits) - 1;
      std::vector<size_t> count(((size_t)1 << bits) + 1, 0);
      for (size_t i = 0; i < n; i++) count[((key[i] >> shift) & mask) + 1]++;
      for (size_t b = 1; b < count.size(); b++) count[b] += count[b - 1];
      for (size_t i = 0; i < n; i++) tmp[count[(key[i] >> shift) & mask]++] = key[i];
      key.swap(tmp);
    }
  }
  auto at = [&](size_t i) -> const NormIv& { return sorted ? v[i] : v[(size_t)(key[i] & 0xffffffffu)]; };
  NormIv cur = at(0);
  int64_t disjoint = (int64_t)n, offsets = 0;
  auto emit = [&]() {
    if (mode == 2) {
      emit_out(NormIv{cur.left, cur.right, 0.0, 0.0, 0.0, 0.0, 0});
    } else {
      emit_out(cur);
    }
    offsets += (int64_t)cur.right - cur.left + 1;
  };
  for (size_t i = 1; i < n; i++) {
    const NormIv& c = at(i);
    const int64_t gap = (int64_t)c.left - 1;
    if (gap <= cur.right) disjoint--;
    const bool merge = (mode == 2) ? (gap <= cur.right)
                                   : (gap < cur.right || (gap == cur.right && jcompare_eq(c.ex, cur.ex) && jcompare_eq(c.ex2, cur.ex2)));
    if (merge) {
      cur.right = std::max(c.right, cur.right);
      cur.ex = jmin(c.ex, cur.ex);
      cur.ex2 = jmin(c.ex2, cur.ex2);
      cur.exu = jmin(c.exu, cur.exu);
      cur.ex2u = jmin(c.ex2u, cur.ex2u);
      cur.bp |= c.bp;
    } else {
      emit();
      cur = c;
    }
  }
  emit();
  if (cnt_disjoint) *cnt_disjoint = disjoint;
  if (cnt_offsets) *cnt_offsets = offsets;
}

inline void norm_sort_merge(const std::vector<NormIv>& v, int mode, std::vector<NormIv>& out, int64_t* cnt_disjoint, int64_t* cnt_offsets) {
  out.clear();
  norm_sort_merge_core(v.data(), v.size(), mode, [&](const NormIv& x) { out.push_back(x); }, cnt_disjoint, cnt_offsets);
}

// CS ∩ CS_i (:333-397 / Dtw :349-425) with ENABLE_BETA_PARTITION and ENABLE_STD_FILTER: an overlap survives when the two
// sides share a beta partition and the smallest variance any window with these block sums can have stays within
// (alpha * stdQ)^2.  pre_length = blocks of w0 points covered by the segments so far (this one included).  dtw = the
// DTW engine's form: the same test from the lower sums, then — overwriting it — from the upper sums.
template <class Emit>
inline void norm_intersect_core(const NormIv* cs, size_t n1, const NormIv* csi, size_t n2, int32_t pre_length, int32_t w0,
                                int32_t query_length, double mean_q, double std_q, double alpha, double beta, int32_t delta_w, bool dtw,
                                Emit&& emit_out) {
  const double limit = alpha * alpha * std_q * std_q;
  size_t i1 = 0, i2 = 0;
  while (i1 < n1 && i2 < n2) {
    const NormIv& a = cs[i1];
    const NormIv& b = csi[i2];
    if (a.right < b.left) {
      i1++;
    } else if (b.right < a.left) {
      i2++;
    } else {
      const int64_t common = a.bp & b.bp;
      if (common == 0) {
        if (a.right < b.right) i1++; else i2++;
        continue;
      }
      const double rest = query_length - pre_length * 1.0 * w0;
      const double sum_ex = a.ex + b.ex, sum_ex2 = a.ex2 + b.ex2;
      double sum_exu = 0.0, sum_ex2u = 0.0;
      double mean = sum_ex / pre_length;
      double std2 = 0.0;
      if (mean > mean_q + beta) {
        const double nv = mean_q + beta - (mean - mean_q - beta) * pre_length * w0 / rest;
        mean = mean_q + beta;
        std2 = (sum_ex2 * w0 + (query_length - pre_length * w0) * nv * nv) / query_length - mean * mean;
      }
      if (dtw) {
        sum_exu = a.exu + b.exu;
        sum_ex2u = a.ex2u + b.ex2u;
        double mean_u = sum_exu / pre_length;
        if (mean_u < mean_q - beta) {
          const double nv = mean_q - beta - (mean_q - beta - mean_u) * pre_length * w0 / rest;
          mean_u = mean_q - beta;
          std2 = (sum_ex2u * w0 + (query_length - pre_length * w0) * nv * nv) / query_length - mean_u * mean_u;
        }
      }
      const bool keep = jcompare_le(std2, limit);
      const int32_t l = std::max(a.left, b.left) + delta_w;
      if (a.right < b.right) {
        if (keep) emit_out(NormIv{l, a.right + delta_w, sum_ex, sum_ex2, sum_exu, sum_ex2u, common});
        i1++;
      } else {
        if (keep) emit_out(NormIv{l, b.right + delta_w, sum_ex, sum_ex2, sum_exu, sum_ex2u, common});
        i2++;
      }
    }
  }
}

inline void norm_intersect(const std::vector<NormIv>& cs, const std::vector<NormIv>& csi, int32_t pre_length, int32_t w0,
                           int32_t query_length, double mean_q, double std_q, double alpha, double beta, int32_t delta_w, bool dtw,
                           std::vector<NormIv>& out) {
  out.clear();
  norm_intersect_core(cs.data(), cs.size(), csi.data(), csi.size(), pre_length, w0, query_length, mean_q, std_q, alpha, beta, delta_w, dtw,
                      [&](const NormIv& x) { out.push_back(x); });
}

// The first segment's positions clamped to window starts inside the series (:313-332 / Dtw :326-348).
template <class Emit>
inline void norm_first_segment_core(const NormIv* pos, size_t n_pos, int32_t order, int32_t w0, int32_t length, int32_t n, int32_t delta_w,
                                    Emit&& emit_out) {
  const int64_t sh = (int64_t)(order - 1) * w0;
  for (size_t i = 0; i < n_pos; i++) {
    const NormIv& p = pos[i];
    NormIv o = p;
    if ((int64_t)p.right - sh + length - 1 > n) {
      if ((int64_t)p.left - sh + length - 1 > n) continue;
      o.left = p.left + delta_w;
      o.right = (int32_t)(n - length + 1 + sh + delta_w);
    } else if ((int64_t)p.left - sh < 1) {
      if ((int64_t)p.right - sh < 1) continue;
      o.left = (int32_t)(1 + sh + delta_w);
      o.right = p.right + delta_w;
    } else {
      o.left = p.left + delta_w;
      o.right = p.right + delta_w;
    }
    emit_out(o);
  }
}

inline void norm_first_segment(const std::vector<NormIv>& pos, int32_t order, int32_t w0, int32_t length, int32_t n, int32_t delta_w,
                               std::vector<NormIv>& out) {
  out.clear();
  norm_first_segment_core(pos.data(), pos.size(), order, w0, length, n, delta_w, [&](const NormIv& x) { out.push_back(x); });
}

}  // namespace kvm_phase1
