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

}  // namespace kvm_phase1
