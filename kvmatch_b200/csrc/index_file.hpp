// index_file.hpp — host side of the index build after the window-mean pass: IndexBuilder step 2 (adjacent-row merge)
// and the index file image.  Replaces K/IndexBuilder.java:308-347 and K/operator/file/IndexFileOperator.java:127-164
// (row codec: K/common/entity/IndexNode.java:51-96; statistic table: K/utils/ByteUtils.java:84-100).
// Input: the (key, first, last) runs kvm_window_mean_runs produced, in the order the reference appends them.
// Rows live in flat interval arrays (no per-row containers); the merge walks two rows with the reference's pending-pair
// state machine, because its output is not the canonical interval union (runs of the SAME row that touch stay separate
// unless a run of the other row bridges them), and long results are re-split at MAXIMUM_DIFF = 256.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <numeric>
#include <unordered_map>
#include <vector>

namespace kvm_index {

struct Iv {
  int32_t lo, hi;
};

constexpr int kMaxDiff = 256;  // IndexNode.MAXIMUM_DIFF

inline int java_cmp(double a, double b) {  // Double.compareTo: numeric, then -0.0 < 0.0, NaN last
  if (a < b) return -1;
  if (a > b) return 1;
  int64_t x, y;
  std::memcpy(&x, &a, 8);
  std::memcpy(&y, &b, 8);
  if (a != a) x = 0x7ff8000000000000LL;
  if (b != b) y = 0x7ff8000000000000LL;
  return x == y ? 0 : (x < y ? -1 : 1);
}

inline void emit_split(std::vector<Iv>& out, Iv p) {  // IndexNodeUtils.addInterval (:82-90)
  while (p.hi - p.lo >= kMaxDiff) {
    out.push_back(Iv{p.lo, p.lo + kMaxDiff - 1});
    p.lo += kMaxDiff;
  }
  out.push_back(p);
}

// IndexNodeUtils.mergeIndexNode (:30-80) on two interval spans.
inline void merge_rows(const Iv* a, size_t na, const Iv* b, size_t nb, std::vector<Iv>& out) {
  out.clear();
  size_t i = 0, j = 0;
  Iv pa{0, 0}, pb{0, 0};
  bool live_a = false, live_b = false;
  while (i < na && j < nb) {
    if (!live_a) pa = a[i], live_a = true;
    if (!live_b) pb = b[j], live_b = true;
    if (pa.hi + 1 < pb.lo) {
      emit_split(out, pa), ++i, live_a = false;
    } else if (pb.hi + 1 < pa.lo) {
      emit_split(out, pb), ++j, live_b = false;
    } else if (pa.hi < pb.hi) {  // a's pending pair is swallowed by b's
      pb.lo = std::min(pb.lo, pa.lo);
      ++i, live_a = false;
    } else {
      pa.lo = std::min(pa.lo, pb.lo);
      ++j, live_b = false;
    }
  }
  for (; i < na; ++i, live_a = false) emit_split(out, live_a ? pa : a[i]);
  for (; j < nb; ++j, live_b = false) emit_split(out, live_b ? pb : b[j]);
}

inline void be32(std::vector<unsigned char>& f, int32_t v) {
  const uint32_t u = (uint32_t)v;
  const unsigned char b[4] = {(unsigned char)(u >> 24), (unsigned char)(u >> 16), (unsigned char)(u >> 8), (unsigned char)u};
  f.insert(f.end(), b, b + 4);
}
inline void be_f64(std::vector<unsigned char>& f, double d) {
  uint64_t u;
  std::memcpy(&u, &d, 8);
  for (int s = 56; s >= 0; s -= 8) f.push_back((unsigned char)(u >> s));
}

// IndexNode.toBytesCompact (:51-96): groups of {left i32}{count-128}{len-128}({gap-128}{len-128})*, a group ends when a
// gap does not fit a byte or it already holds 254 follow-up intervals.
inline void append_compact(std::vector<unsigned char>& f, const Iv* p, size_t n) {
  size_t k = 0;
  while (k < n) {
    be32(f, p[k].lo);
    const size_t count_at = f.size();
    f.push_back(0);
    f.push_back((unsigned char)(p[k].hi - p[k].lo - 128));
    int follow = 0;  // (count - 1) / 2 of the reference
    ++k;
    while (k < n) {
      const int gap = p[k].lo - p[k - 1].hi;
      if (!(gap < kMaxDiff && follow + 2 < kMaxDiff)) break;
      f.push_back((unsigned char)(gap - 128));
      f.push_back((unsigned char)(p[k].hi - p[k].lo - 128));
      ++follow;
      ++k;
    }
    f[count_at] = (unsigned char)(follow - 128);
  }
}

// IndexNode.parseBytesCompact (:108-128): the inverse of append_compact over one row's bytes (the key excluded).
// Returns the number of (left, right) pairs, written to lr_out while they fit `cap`; -1 on a truncated row.
inline int64_t parse_compact(const unsigned char* b, int64_t nb, int32_t* lr_out, int64_t cap) {
  int64_t k = 0, idx = 0;
  auto put = [&](int32_t l, int32_t r) {
    if (k < cap) {
      lr_out[2 * k] = l;
      lr_out[2 * k + 1] = r;
    }
    ++k;
  };
  while (idx < nb) {
    if (idx + 6 > nb) return -1;
    int32_t left = (int32_t)(((uint32_t)b[idx] << 24) | ((uint32_t)b[idx + 1] << 16) | ((uint32_t)b[idx + 2] << 8) | (uint32_t)b[idx + 3]);
    const int count = (int)(b[idx + 4] ^ 0x80);  // (signed byte) + 128
    int32_t right = left + (int32_t)(b[idx + 5] ^ 0x80);
    idx += 6;
    put(left, right);
    if (idx + 2 * (int64_t)count > nb) return -1;
    for (int j = 0; j < count; j++) {
      left = right + (int32_t)(b[idx] ^ 0x80);
      right = left + (int32_t)(b[idx + 1] ^ 0x80);
      idx += 2;
      put(left, right);
    }
  }
  return k;
}

struct ImageInfo {
  int32_t rows_step1 = 0, rows = 0;
  int64_t intervals = 0, offsets = 0;
};

// Returns false when there is no run at all (the reference throws on rawStatisticInfo.get(0)).
inline bool build_image(const double* keys, const int32_t* first, const int32_t* last, int64_t n_runs,
                        std::vector<unsigned char>& file, ImageInfo* info) {
  file.clear();
  if (n_runs <= 0) return false;
  // step-1 rows: runs grouped by key, inside a row in append order (the HashMap<Double, IndexNode> of :268-306).
  // Few distinct keys, many runs: hash the key bits to a dense id, counting-sort the runs by id (stable), and order
  // only the distinct keys.
  auto bits_of = [](double d) {
    uint64_t u;
    std::memcpy(&u, &d, 8);
    return (d != d) ? 0x7ff8000000000000ULL : u;  // Double.equals: NaNs are one key
  };
  // open addressing on the key bits (a node-based map costs 30 ns per run here; this one 5): slot = id + 1, 0 = empty
  int cap_bits = 12;
  size_t cap = (size_t)1 << cap_bits;
  std::vector<uint64_t> slot_key(cap);
  std::vector<int32_t> slot_id(cap, 0);
  std::vector<int32_t> run_id((size_t)n_runs);
  std::vector<double> id_key;
  std::vector<int64_t> id_count;
  auto slot_of = [&](uint64_t b) {
    size_t h = (size_t)((b * 0x9E3779B97F4A7C15ULL) >> (64 - cap_bits));  // the product's top bits
    while (slot_id[h] != 0 && slot_key[h] != b) h = (h + 1) & (cap - 1);
    return h;
  };
  for (int64_t r = 0; r < n_runs; r++) {
    const uint64_t b = bits_of(keys[r]);
    size_t h = slot_of(b);
    if (slot_id[h] == 0) {
      if (2 * (id_key.size() + 1) > cap) {  // keep the load under one half
        cap_bits += 2;
        cap = (size_t)1 << cap_bits;
        slot_key.assign(cap, 0);
        slot_id.assign(cap, 0);
        for (size_t i = 0; i < id_key.size(); i++) {
          const uint64_t bi = bits_of(id_key[i]);
          const size_t hi = slot_of(bi);
          slot_key[hi] = bi;
          slot_id[hi] = (int32_t)i + 1;
        }
        h = slot_of(b);
      }
      slot_key[h] = b;
      slot_id[h] = (int32_t)id_key.size() + 1;
      id_key.push_back(keys[r]);
      id_count.push_back(0);
    }
    const int32_t id = slot_id[h] - 1;
    run_id[(size_t)r] = id;
    id_count[(size_t)id]++;
  }
  const size_t n_ids = id_key.size();
  std::vector<int32_t> by_key(n_ids);
  std::iota(by_key.begin(), by_key.end(), 0);
  std::sort(by_key.begin(), by_key.end(), [&](int32_t x, int32_t y) { return java_cmp(id_key[(size_t)x], id_key[(size_t)y]) < 0; });
  std::vector<double> row_key(n_ids);
  std::vector<int64_t> row_begin(n_ids + 1), cursor(n_ids);
  int64_t acc = 0;
  for (size_t k = 0; k < n_ids; k++) {
    row_key[k] = id_key[(size_t)by_key[k]];
    row_begin[k] = acc;
    cursor[(size_t)by_key[k]] = acc;
    acc += id_count[(size_t)by_key[k]];
  }
  row_begin[n_ids] = acc;
  std::vector<Iv> flat((size_t)n_runs);
  for (int64_t r = 0; r < n_runs; r++) flat[(size_t)cursor[(size_t)run_id[(size_t)r]]++] = Iv{first[r], last[r]};
  const int64_t R = (int64_t)row_key.size();
  info->rows_step1 = (int32_t)R;
  const double average = (double)n_runs / (double)R;  // mean #intervals per row (StatisticInfo.getAverage)
  // step 2 (:321-343): from the largest key downwards, fold a row into the running one when it is small
  // (< 1.2 x average) and the fold saves more than 20 % of the intervals; a closed group takes its smallest key.
  struct Out {
    double key;
    size_t begin, end;  // span in `store`
  };
  std::vector<Out> rows;
  std::vector<Iv> store, run(flat.begin() + row_begin[(size_t)R - 1], flat.begin() + row_begin[(size_t)R]), tmp;
  auto close_group = [&](double key) {
    rows.push_back(Out{key, store.size(), store.size() + run.size()});
    store.insert(store.end(), run.begin(), run.end());
  };
  for (int64_t r = R - 2; r >= 0; r--) {
    const Iv* cur = flat.data() + row_begin[(size_t)r];
    const size_t ncur = (size_t)(row_begin[(size_t)r + 1] - row_begin[(size_t)r]);
    bool folded = false;
    if ((double)ncur < average * 1.2) {
      merge_rows(run.data(), run.size(), cur, ncur, tmp);
      if ((double)tmp.size() < (double)(run.size() + ncur) * 0.8) {
        run.swap(tmp);
        folded = true;
      }
    }
    if (!folded) {
      close_group(row_key[(size_t)r + 1]);
      run.assign(cur, cur + ncur);
    }
  }
  close_group(row_key[0]);
  std::reverse(rows.begin(), rows.end());  // ascending keys (TreeMap order)
  info->rows = (int32_t)rows.size();
  // file image: rows, cumulative (key, #intervals, #offsets) table, offset table
  std::vector<int32_t> offs;
  for (const Out& o : rows) {
    offs.push_back((int32_t)file.size());
    be_f64(file, o.key);
    append_compact(file, store.data() + o.begin, o.end - o.begin);
  }
  offs.push_back((int32_t)file.size());
  int32_t cum_iv = 0, cum_off = 0;
  for (const Out& o : rows) {
    cum_iv += (int32_t)(o.end - o.begin);
    for (size_t k = o.begin; k < o.end; k++) cum_off += store[k].hi - store[k].lo + 1;
    be_f64(file, o.key);
    be32(file, cum_iv);
    be32(file, cum_off);
  }
  info->intervals = cum_iv;
  info->offsets = cum_off;
  offs.push_back((int32_t)file.size());
  for (int32_t o : offs) be32(file, o);
  return true;
}

}  // namespace kvm_index
