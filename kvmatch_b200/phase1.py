"""Host-side mirror of KV-match's phase 0 / phase 1 for the RSM-ED engine (K/QueryEngine.java:155-334) over the local-file
index (K/operator/file/IndexFileOperator.java), so that phase-2 verification can be driven with REAL candidate lists.

What is mirrored: the index file reader (offset table, statistic table, range scan, compact interval codec), the query
segmentation DP (determineQueryPlan, :398-503), the per-segment index scan with its distance lower bound
(:505-521, :383-396), and the phase-1 loop (:185-334).  The interval algebra of that loop — sortButNotMergeIntervals,
CS ∩ CS_i, sortAndMergeIntervals — runs in the library (kvm_intervals_*, csrc/phase1.hpp).  Not mirrored: incremental
index visiting (a cache of already scanned rows: same positions, fewer file reads) and the wall-clock driven early
termination (:296-307), which makes the reference's own candidate list irreproducible; it is off.
"""
from __future__ import annotations

import bisect
import ctypes as C
import math
import struct
from dataclasses import dataclass

import numpy as np

from . import _lib

WU_LIST = (25, 50, 100, 200, 400)                      # the enabled widths: the paper's Sigma
WU_ALL = tuple(25 * k for k in range(1, 17))           # K/QueryEngine.java:51-52: WuList = 25..400 step 25, WuEnabled = Sigma


# ---------------------------------------------------------------- MeanIntervalUtils (K/utils/MeanIntervalUtils.java)
def to_round(value: float) -> float:
    value *= 10.0
    int_value = math.floor(value)
    ret = int_value + (0.5 if value - int_value >= 0.5 else 0.0)
    return ret * 0.1


def to_round_stat(value: float, keys) -> float:
    """toRound(value, statisticInfo) :75-85: the largest existing row key <= the rounded value."""
    r = to_round(value)
    i = bisect.bisect_left(keys, r)
    if i < len(keys) and keys[i] == r:
        return r
    i -= 1
    return r - 10000 if i < 0 else keys[i]


def to_upper_stat(rnd: float, keys) -> float:
    """toUpper(round, statisticInfo) :107-117."""
    r = (rnd * 10.0 + 0.5) * 0.1
    i = bisect.bisect_left(keys, r)
    if i < len(keys) and keys[i] == r:
        return r
    return r + 10000 if i >= len(keys) else keys[i]


# ---------------------------------------------------------------- index file reader (IndexFileOperator.java:52-125)
class IndexFile:
    """files/index-<N>-<w>: rows [key f64 BE][compact intervals] in ascending key, the cumulative statistic table, and
    the trailing int32 offset table whose last two entries locate the statistic table and the offset table."""

    def __init__(self, data: bytes):
        self.data = data
        self._buf = np.frombuffer(data, dtype=np.uint8)      # (keeps the bytes addressable for the library's row decoder)
        self._base = self._buf.ctypes.data
        n = len(data)
        last = struct.unpack(">i", data[n - 4:n])[0]                   # readOffsetInfo :52-62
        self.offsets = list(struct.unpack(f">{(n - last) // 4}i", data[last:n]))
        self.n_rows = len(self.offsets) - 2
        self.keys = [struct.unpack(">d", data[o:o + 8])[0] for o in self.offsets[:self.n_rows]]
        s0, s1 = self.offsets[-2], self.offsets[-1]                    # readStatisticInfo :85-91
        self.stat = [struct.unpack(">dii", data[o:o + 16]) for o in range(s0, s1, 16)]
        self.stat_keys = [t[0] for t in self.stat]

    @classmethod
    def open(cls, path: str):
        with open(path, "rb") as f:
            return cls(f.read())

    def row_array(self, i: int):
        """(key, int32 array [k, 2] of (left, right)) of row i: IndexNode.parseBytesCompact, K/common/entity/IndexNode.java:108-128.
        A group is [left i32][count][right - left] followed by `count` (gap, width) byte pairs, every byte stored as
        value - 128 (a signed Java byte); decoded by the library (kvm_index_row_positions, csrc/index_file.hpp)."""
        lo, hi = self.offsets[i], self.offsets[i + 1]
        key = struct.unpack(">d", self.data[lo:lo + 8])[0]
        nb = hi - lo - 8
        out = np.empty((max(nb // 2, 1), 2), dtype=np.int32)
        k = C.c_int64()
        rc = _lib.load().kvm_index_row_positions(self._base + lo + 8, nb, out.ctypes.data, len(out), C.byref(k))
        if rc:
            raise _lib.KvmError(rc, "kvm_index_row_positions")
        return key, out[:k.value]

    def row(self, i: int):
        """(key, [(left, right), ...]) of row i."""
        key, arr = self.row_array(i)
        return key, [tuple(p) for p in arr.tolist()]

    def read_indexes(self, key_from: float, key_to: float):
        """readIndexes :64-83: rows with key_from <= key <= key_to (lowerBound / upperBound on the row keys)."""
        lo = bisect.bisect_left(self.keys, key_from)
        hi = bisect.bisect_right(self.keys, key_to) - 1
        return [self.row(i) for i in range(lo, hi + 1)] if lo < self.n_rows and hi >= 0 else []

    def parts(self):
        """The files that make up this index (one; see ShardedIndexFile)."""
        return [self]

    def read_index_arrays(self, key_from: float, key_to: float):
        """read_indexes with the positions as int32 arrays (the cNSM scans fill structured arrays from them)."""
        lo = bisect.bisect_left(self.keys, key_from)
        hi = bisect.bisect_right(self.keys, key_to) - 1
        return [self.row_array(i) for i in range(lo, hi + 1)] if lo < self.n_rows and hi >= 0 else []


class ShardedIndexFile:
    """One window width's index as SEVERAL files, each covering a contiguous range of window starts (the per-shard layout
    of SURVEY 8(f) f2: a file's offset table holds int32 byte offsets, so one file ends at 2 GiB — about n = 1.5e9 for
    w = 25 on the synthetic series; and one file per GPU shard is what a multi-GPU build writes).  Positions stay global
    1-based window starts.  Every part keeps its own rows (step 2 merges rows per file), so range scans round the lower
    end against each part's own keys; the statistic table the plan DP reads is the parts' tables added up on the union
    of their keys (cumulative counts are sums over disjoint position ranges)."""

    def __init__(self, images):
        self._parts = [IndexFile(b) for b in images]
        keys = sorted({k for p in self._parts for k in p.stat_keys})
        self.stat = []
        for k in keys:
            iv = off = 0
            for p in self._parts:
                i = bisect.bisect_right(p.stat_keys, k) - 1
                if i >= 0:
                    iv += p.stat[i][1]
                    off += p.stat[i][2]
            self.stat.append((k, iv, off))
        self.stat_keys = keys

    def parts(self):
        return self._parts


def open_index(image):
    """One width's index from its file image (bytes) or, for the per-shard layout, the list of its files' images."""
    return IndexFile(image) if isinstance(image, (bytes, bytearray, memoryview)) else ShardedIndexFile(image)


def shard_ranges(n_windows: int, shards: int):
    """`shards` contiguous (lo, hi) ranges of 1-based window starts covering [1, n_windows]."""
    per = -(-n_windows // shards)
    return [(k * per + 1, min((k + 1) * per, n_windows)) for k in range(shards) if k * per < n_windows]


def split_runs(keys, first, last, ranges):
    """Step-1 runs of one width cut at shard boundaries: for each (lo, hi) range of 1-based window starts, the runs (or
    pieces of runs) inside it, in position order.  Keys and positions are those of the single-file index."""
    keys, first, last = np.asarray(keys), np.asarray(first), np.asarray(last)
    out = []
    for lo, hi in ranges:
        a = int(np.searchsorted(last, lo, side="left"))     # first run ending at or after lo
        b = int(np.searchsorted(first, hi, side="right"))   # one past the last run starting at or before hi
        f, l = first[a:b].copy(), last[a:b].copy()
        if len(f):
            f[0] = max(int(f[0]), lo)
            l[-1] = min(int(l[-1]), hi)
        out.append((keys[a:b].copy(), f, l))
    return out


# ---------------------------------------------------------------- interval algebra: the library's host functions
# An interval list is (lr int32 [k, 2], eps float64 [k]) inside this module; the list-of-tuples forms below are the same
# calls for small inputs (tests, notebooks).
def _pack(ivs):
    k = len(ivs)
    lr = np.zeros((max(k, 1), 2), dtype=np.int32)
    eps = np.zeros(max(k, 1))
    for i, (l, r, e) in enumerate(ivs):
        lr[i] = (l, r)
        eps[i] = e
    return lr[:k], eps[:k]


def _unpack(lr, eps):
    return [(int(a), int(b), float(e)) for (a, b), e in zip(lr.tolist(), eps.tolist())]


def _room(k):
    return np.empty((max(k, 1), 2), dtype=np.int32), np.empty(max(k, 1))


def sort_merge_arrays(lr, eps, mode: int):
    """kvm_intervals_sort_merge on arrays -> (lr, eps, cnt_disjoint, cnt_offsets)."""
    L = _lib.load()
    lr, eps = np.ascontiguousarray(lr, dtype=np.int32), np.ascontiguousarray(eps, dtype=np.float64)
    k = len(eps)
    lo, eo = _room(k)
    ko, cd, co = C.c_int64(), C.c_int64(), C.c_int64()
    rc = L.kvm_intervals_sort_merge(lr.ctypes.data, eps.ctypes.data, k, mode, lo.ctypes.data, eo.ctypes.data, len(eo),
                                    C.byref(ko), C.byref(cd), C.byref(co))
    if rc:
        raise _lib.KvmError(rc, "kvm_intervals_sort_merge")
    return lo[:ko.value], eo[:ko.value], cd.value, co.value


def intersect_arrays(cs_lr, cs_eps, csi_lr, csi_eps, eps2: float, delta_w: int):
    """kvm_intervals_intersect on arrays -> (lr, eps, smallest summed bound kept)."""
    L = _lib.load()
    a, ae = np.ascontiguousarray(cs_lr, dtype=np.int32), np.ascontiguousarray(cs_eps, dtype=np.float64)
    b, be = np.ascontiguousarray(csi_lr, dtype=np.int32), np.ascontiguousarray(csi_eps, dtype=np.float64)
    lo, eo = _room(len(ae) + len(be))
    ko, me = C.c_int64(), C.c_double()
    rc = L.kvm_intervals_intersect(a.ctypes.data, ae.ctypes.data, len(ae), b.ctypes.data, be.ctypes.data, len(be), eps2, delta_w,
                                   lo.ctypes.data, eo.ctypes.data, len(eo), C.byref(ko), C.byref(me))
    if rc:
        raise _lib.KvmError(rc, "kvm_intervals_intersect")
    return lo[:ko.value], eo[:ko.value], me.value


def first_segment_arrays(lr, eps, order: int, length: int, n: int, delta_w: int):
    L = _lib.load()
    a, ae = np.ascontiguousarray(lr, dtype=np.int32), np.ascontiguousarray(eps, dtype=np.float64)
    lo, eo = _room(len(ae))
    ko, me = C.c_int64(), C.c_double()
    rc = L.kvm_intervals_first_segment(a.ctypes.data, ae.ctypes.data, len(ae), order, WU_LIST[0], length, n, delta_w, lo.ctypes.data,
                                       eo.ctypes.data, len(eo), C.byref(ko), C.byref(me))
    if rc:
        raise _lib.KvmError(rc, "kvm_intervals_first_segment")
    return lo[:ko.value], eo[:ko.value], me.value


def sort_merge(ivs, mode: int):
    """kvm_intervals_sort_merge -> (intervals, cnt_disjoint, cnt_offsets)."""
    lo, eo, cd, co = sort_merge_arrays(*_pack(ivs), mode)
    return _unpack(lo, eo), cd, co


def intersect(cs, csi, eps2: float, delta_w: int):
    lo, eo, me = intersect_arrays(*_pack(cs), *_pack(csi), eps2, delta_w)
    return _unpack(lo, eo), me


def first_segment(pos, order: int, length: int, n: int, delta_w: int):
    lo, eo, me = first_segment_arrays(*_pack(pos), order, length, n, delta_w)
    return _unpack(lo, eo), me


# ---------------------------------------------------------------- phase 0: determineQueryPlan (K/QueryEngine.java:398-503)
@dataclass
class QuerySegment:
    mean: float
    order: int
    count: int
    wu: int


_KEYS_CACHE: dict = {}


def _keys_of(stat):
    """Row keys of a cumulative statistic table (built once per table: the plan DP looks them up hundreds of times)."""
    hit = _KEYS_CACHE.get(id(stat))
    if hit is None or hit[0] is not stat:
        if len(_KEYS_CACHE) > 64:
            _KEYS_CACHE.clear()
        hit = (stat, [t[0] for t in stat])
        _KEYS_CACHE[id(stat)] = hit
    return hit[1]


def _counts(stat, wu: int, mean: float, epsilon: float):
    """getCountsFromStatisticInfo :382-399 on the cumulative table of width wu."""
    keys = _keys_of(stat)
    rng = epsilon / math.sqrt(wu)
    begin, end = to_round(mean - rng), to_round(mean + rng)

    def search(key):
        i = bisect.bisect_left(keys, key)
        return min(i, len(stat) - 1)
    i = search(begin)
    lower1 = stat[i - 1][1] if i > 0 else 0
    lower2 = stat[i - 1][2] if i > 0 else 0
    i = search(end)
    upper1 = stat[i][1] if i > 0 else 0
    upper2 = stat[i][2] if i > 0 else 0
    return upper1 - lower1, upper2 - lower2


@dataclass
class RangeQuerySegment:
    mean_min: float
    mean_max: float
    order: int
    count: int
    wu: int


def _block_prefix(values, w0: int):
    """Sums of the disjoint blocks of w0 points and their running prefix (:597-606, :627-629)."""
    sums, ex = [], 0.0
    for i, v in enumerate(values):
        ex += float(v)
        if (i + 1) % w0 == 0:
            sums.append(ex)
            ex = 0.0
    prefix = [0.0] * len(sums)
    prefix[0] = sums[0]
    for i in range(1, len(sums)):
        prefix[i] = prefix[i - 1] + sums[i]
    return prefix


def determine_query_plan(q, epsilon: float, stats, counts=None, bounds=None):
    """stats: {width: cumulative statistic table} for the widths of WU_LIST.  `counts(stat, wu, *means)` = the engine's
    getCountsFromStatisticInfo (default: the RSM-ED one); the DP itself is the same in K/QueryEngine.java:398-503,
    K/NormQueryEngine.java:593-671 and, over the block means of the query's lower / upper envelope (`bounds` = (L, U),
    segments become RangeQuerySegments), in K/NormQueryEngineDtw.java:670-799."""
    if counts is None:
        counts = lambda stat, wu, mean: _counts(stat, wu, mean, epsilon)
    w0 = WU_ALL[0]
    m = len(q) // w0
    prefixes = [_block_prefix(q, w0)] if bounds is None else [_block_prefix(bounds[0], w0), _block_prefix(bounds[1], w0)]
    total100 = stats[100][-1][1]
    cost, cost2 = {}, {}

    def means(l, r):
        use = w0 * (r - l + 1)
        return [(p[r] - (p[l - 1] if l > 0 else 0.0)) / use for p in prefixes]

    def get_cost(l, r):
        if (l, r) not in cost:
            use = w0 * (r - l + 1)
            c1, _ = counts(stats[use], use, *means(l, r))
            cost[(l, r)] = math.log(1.0 * c1 / total100) if c1 > 0 else -math.inf
            cost2[(l, r)] = c1
        return cost[(l, r)]
    INF = 1.7976931348623157e308
    dp = [[INF] * (m + 1) for _ in range(m + 1)]
    pre = [[-1] * (m + 1) for _ in range(m + 1)]
    dp[0][0] = 0.0
    for i in range(1, m + 1):
        for j in range(1, min(i, 30) + 1):
            for k in range(1, len(WU_ALL) + 1):
                if i - k < 0:
                    break
                if WU_ALL[k - 1] not in stats:   # WuEnabled
                    continue
                tmp = ((j - 1) * dp[i - k][j - 1] + get_cost(i - k, i - 1)) / j
                if tmp < dp[i][j]:
                    dp[i][j] = tmp
                    pre[i][j] = k
    best, p = INF, -1
    start = (31 - (32 - len(q).bit_length()) - 1) // 2   # (31 - numberOfLeadingZeros(size) - 1) / 2
    for i in range(start, min(m, 30) + 1):
        if dp[m][i] <= best:
            best, p = dp[m][i], i
    queries, index = [], m
    for i in range(p, -1, -1):
        l, r = index - pre[index][i], index - 1
        use = w0 * (r - l + 1)
        if use < 0:
            break
        get_cost(l, r)
        mm = means(l, r)
        queries.append(QuerySegment(mm[0], l + 1, cost2[(l, r)], use) if bounds is None
                       else RangeQuerySegment(mm[0], mm[1], l + 1, cost2[(l, r)], use))
        index -= pre[index][i]
    queries.sort(key=lambda s: s.count)   # ENABLE_QUERY_REORDERING (stable sort)
    return queries


# ---------------------------------------------------------------- phase 1 (K/QueryEngine.java:185-334)
def _scan_rows(idx, lo_value: float, end: float, bound_of):
    """Rows with toRound(lo_value, statisticInfo) <= key <= end + 0.01 as one interval list (every part of a sharded index
    rounds against its own row keys); bound_of(key, upper) = the row's distance lower bound."""
    rows = []
    for part in idx.parts():
        begin = to_round_stat(lo_value, part.stat_keys)
        rows.extend((key, positions, to_upper_stat(key, part.stat_keys)) for key, positions in part.read_index_arrays(begin, end + 0.01))
    lr = np.empty((sum(len(p) for _, p, _ in rows), 2), dtype=np.int32)
    eps = np.empty(len(lr))
    at = 0
    for key, positions, upper in rows:
        k = len(positions)
        lr[at:at + k] = positions
        eps[at:at + k] = bound_of(key, upper)
        at += k
    return lr, eps


def scan_index(idx: IndexFile, seg: QuerySegment, begin: float, end: float):
    """scanIndex :505-521 with getDistanceLowerBound :383-396: positions with wu * lower bound, as (lr, eps) arrays."""
    def bound(key, upper):
        if key > seg.mean:
            delta = (key - seg.mean) * (key - seg.mean)
        elif upper < seg.mean:
            delta = (seg.mean - upper) * (seg.mean - upper)
        else:
            delta = 0.0
        return seg.wu * delta
    return _scan_rows(idx, begin, end, bound)


def _rsm_loop(queries, epsilon: float, length: int, n: int, by_w, scan, seg_range, reset_last_min: bool):
    """The phase-1 loop shared by K/QueryEngine.java:185-334 and K/QueryEngineDtw.java:195-347."""
    v_lr, v_eps = np.empty((0, 2), dtype=np.int32), np.empty(0)
    last_min = 0.0
    range0 = epsilon * epsilon
    for i, seg in enumerate(queries):
        delta_w = 0 if i == len(queries) - 1 else (queries[i + 1].order - seg.order) * WU_LIST[0]
        ix = by_w[seg.wu]
        if reset_last_min and last_min > range0:   # K/QueryEngineDtw.java:210
            last_min = 0.0
        rng = math.sqrt((range0 - last_min) / seg.wu)
        lo, hi = seg_range(seg)
        begin = lo - rng        # (rounded against the statistic table inside the scan)
        end = to_round(hi + rng)
        p_lr, p_eps, _, _ = sort_merge_arrays(*scan(ix, seg, begin, end), 0)
        if i == 0:
            n_lr, n_eps, last_min = first_segment_arrays(p_lr, p_eps, seg.order, length, n, delta_w)
        else:
            n_lr, n_eps, last_min = intersect_arrays(v_lr, v_eps, p_lr, p_eps, range0, delta_w)
        v_lr, v_eps, _, _ = sort_merge_arrays(n_lr, n_eps, 1)
        if len(v_eps) == 0:
            break
    m_lr, _, _, _ = sort_merge_arrays(v_lr, v_eps, 2)
    return [(int(l), int(r)) for l, r in m_lr.tolist()], queries[-1].order, queries


def phase1(q, epsilon: float, n: int, indexes):
    """Candidate intervals of an RSM-ED query: (valid_positions [(left, right)], last_segment, plan).  `indexes` = one
    IndexFile per width of WU_LIST."""
    by_w = dict(zip(WU_LIST, indexes))
    queries = determine_query_plan(q, epsilon, {w: ix.stat for w, ix in by_w.items()})
    return _rsm_loop(queries, epsilon, len(q), n, by_w, scan_index, lambda seg: (seg.mean, seg.mean), False)


# ---------------------------------------------------------------- cNSM-ED: phases 0 / 1 of K/NormQueryEngine.java:177-430
NORM_IV = np.dtype([("left", "<i4"), ("right", "<i4"), ("ex", "<f8"), ("ex2", "<f8"), ("exu", "<f8"), ("ex2u", "<f8"), ("bp", "<i8")])   # = kvm_norm_interval
BETA_PARTITION_WIDTH = 10.0   # :59


def _norm_out(k):
    return np.empty(max(k, 1), dtype=NORM_IV)


def norm_sort_merge(ivs: np.ndarray, mode: int):
    """kvm_norm_intervals_sort_merge -> (intervals, cnt_disjoint, cnt_offsets)."""
    L = _lib.load()
    ivs = np.ascontiguousarray(ivs, dtype=NORM_IV)
    out = _norm_out(len(ivs))
    ko, cd, co = C.c_int64(), C.c_int64(), C.c_int64()
    rc = L.kvm_norm_intervals_sort_merge(ivs.ctypes.data, len(ivs), mode, out.ctypes.data, len(out), C.byref(ko), C.byref(cd), C.byref(co))
    if rc:
        raise _lib.KvmError(rc, "kvm_norm_intervals_sort_merge")
    return out[:ko.value], cd.value, co.value


def norm_intersect(cs, csi, pre_length: int, query_length: int, mean_q: float, std_q: float, alpha: float, beta: float, delta_w: int,
                   dtw: bool = False):
    L = _lib.load()
    cs, csi = np.ascontiguousarray(cs, dtype=NORM_IV), np.ascontiguousarray(csi, dtype=NORM_IV)
    out = _norm_out(len(cs) + len(csi))
    ko = C.c_int64()
    rc = L.kvm_norm_intervals_intersect(cs.ctypes.data, len(cs), csi.ctypes.data, len(csi), pre_length, WU_LIST[0], query_length,
                                        mean_q, std_q, alpha, beta, delta_w, int(dtw), out.ctypes.data, len(out), C.byref(ko))
    if rc:
        raise _lib.KvmError(rc, "kvm_norm_intervals_intersect")
    return out[:ko.value]


def norm_first_segment(pos, order: int, length: int, n: int, delta_w: int):
    L = _lib.load()
    pos = np.ascontiguousarray(pos, dtype=NORM_IV)
    out = _norm_out(len(pos))
    ko = C.c_int64()
    rc = L.kvm_norm_intervals_first_segment(pos.ctypes.data, len(pos), order, WU_LIST[0], length, n, delta_w, out.ctypes.data, len(out),
                                            C.byref(ko))
    if rc:
        raise _lib.KvmError(rc, "kvm_norm_intervals_first_segment")
    return out[:ko.value]


def norm_mean_range(mean: float, wu: int, epsilon: float, alpha: float, beta: float, mean_q: float, std_q: float, lo_shift=0.0, hi_shift=None):
    """The window-mean range a segment of mean `mean` admits under (alpha, beta): the expressions of :225-231 (whole range:
    lo_shift = 0, hi_shift = 2 beta) and :244-250 (one beta partition), operation for operation."""
    if hi_shift is None:
        hi_shift = 2.0 * beta
        begin = 1.0 / alpha * mean + (1 - 1.0 / alpha) * mean_q - beta - math.sqrt(1.0 / (alpha * alpha) * std_q * std_q * epsilon * epsilon / wu)
        begin1 = alpha * mean + (1 - alpha) * mean_q - beta - math.sqrt(alpha * alpha * std_q * std_q * epsilon * epsilon / wu)
        end = alpha * mean + (1 - alpha) * mean_q + beta + math.sqrt(alpha * alpha * std_q * std_q * epsilon * epsilon / wu)
        end1 = 1.0 / alpha * mean + (1 - 1.0 / alpha) * mean_q + beta + math.sqrt(1.0 / (alpha * alpha) * std_q * std_q * epsilon * epsilon / wu)
    else:
        begin = 1.0 / alpha * mean + (1 - 1.0 / alpha) * mean_q - beta + lo_shift - math.sqrt(1.0 / (alpha * alpha) * std_q * std_q * epsilon * epsilon / wu)
        begin1 = alpha * mean + (1 - alpha) * mean_q - beta + lo_shift - math.sqrt(alpha * alpha * std_q * std_q * epsilon * epsilon / wu)
        end = alpha * mean + (1 - alpha) * mean_q - beta + hi_shift + math.sqrt(alpha * alpha * std_q * std_q * epsilon * epsilon / wu)
        end1 = 1.0 / alpha * mean + (1 - 1.0 / alpha) * mean_q - beta + hi_shift + math.sqrt(1.0 / (alpha * alpha) * std_q * std_q * epsilon * epsilon / wu)
    return min(begin, begin1), max(end, end1)


def _counts_norm(stat, wu: int, mean: float, epsilon: float, alpha: float, beta: float, mean_q: float, std_q: float):
    """getCountsFromStatisticInfo :547-572 (both ends through the plain toRound)."""
    keys = _keys_of(stat)
    lo, hi = norm_mean_range(mean, wu, epsilon, alpha, beta, mean_q, std_q)
    begin, end = to_round(lo), to_round(hi)

    def search(key):
        return min(bisect.bisect_left(keys, key), len(stat) - 1)
    i = search(begin)
    lower1 = stat[i - 1][1] if i > 0 else 0
    lower2 = stat[i - 1][2] if i > 0 else 0
    i = search(end)
    upper1 = stat[i][1] if i > 0 else 0
    upper2 = stat[i][2] if i > 0 else 0
    return upper1 - lower1, upper2 - lower2


def beta_partitions(seg: QuerySegment, epsilon: float, alpha: float, beta: float, mean_q: float, std_q: float):
    """:233-254: (int)(2 beta / 10) partitions (at most 64) of the beta range, each with its own rounded mean range.  beta < 5
    gives ZERO partitions, hence empty bit sets and — from the second segment on — an empty candidate set: the reference's
    behaviour, kept."""
    num = min(int(2.0 * beta / BETA_PARTITION_WIDTH), 64)
    parts = []
    for idx in range(num):
        width = 2.0 * beta / num
        lo, hi = norm_mean_range(seg.mean, seg.wu, epsilon, alpha, beta, mean_q, std_q, width * idx, width * (idx + 1))
        parts.append((lo, to_round(hi)))   # (the lower end is rounded against the statistic table inside the scan)
    return parts


def _java_int_shl1(idx: int) -> int:
    """`1 << idx` on a Java int (:692: the shift distance wraps at 32 and bit 31 sign-extends into the long it is or-ed to)."""
    v = (1 << (idx & 31)) & 0xFFFFFFFF
    return v - (1 << 32) if v & 0x80000000 else v


def _scan_rows_norm(idx, lo_value: float, end: float, parts, sums_of) -> np.ndarray:
    """The cNSM scans: rows of every part of the index with their block sums (sums_of(key, upper) -> ex, ex2, exu, ex2u)
    and the beta partitions the row key falls into.  `parts` = (unrounded lower end, rounded upper end) per partition;
    lower ends are rounded against the statistic table of the file being scanned (toRound(value, statisticInfo))."""
    rows = []
    for part in idx.parts():
        begin = to_round_stat(lo_value, part.stat_keys)
        ranges = [(to_round_stat(p_lo, part.stat_keys), p_hi) for p_lo, p_hi in parts]
        for key, positions in part.read_index_arrays(begin, end + 0.01):
            bits = 0
            for j, (p_lo, p_hi) in enumerate(ranges):
                if p_lo > key:
                    break
                if p_lo <= key <= p_hi:
                    bits |= _java_int_shl1(j)
            rows.append((positions, sums_of(key, to_upper_stat(key, part.stat_keys)), bits))
    out = np.zeros(sum(len(p) for p, _, _ in rows), dtype=NORM_IV)
    at = 0
    for positions, sums, bits in rows:
        k = len(positions)
        if k:
            blk = out[at:at + k]
            blk["left"], blk["right"] = positions[:, 0], positions[:, 1]
            blk["ex"], blk["ex2"], blk["exu"], blk["ex2u"] = sums
            blk["bp"] = bits
            at += k
    return out


def scan_index_norm(idx, seg: QuerySegment, lo_value: float, end: float, parts) -> np.ndarray:
    """scanIndex :673-701: positions with the row's lower block sums (key * blocks, key'^2 * blocks with key' = the row's
    upper end for negative keys) and the beta partitions the row key falls into."""
    blocks = seg.wu // WU_ALL[0]

    def sums(key, upper):
        key2 = upper if key < 0 else key
        return key * blocks, key2 * key2 * blocks, 0.0, 0.0
    return _scan_rows_norm(idx, lo_value, end, parts, sums)


def query_statistics(q):
    """Phase 0 :192-198: sequential sums."""
    ex = ex2 = 0.0
    for v in q:
        v = float(v)
        ex += v
        ex2 += v * v
    mean_q = ex / len(q)
    return mean_q, math.sqrt(ex2 / len(q) - mean_q * mean_q)


PROBE_THREADS = 4   # host threads probing the index for the segments of one cNSM query (the library calls drop the GIL)


def _probe_all(probe, queries):
    if PROBE_THREADS <= 1 or len(queries) <= 1:
        return [probe(seg) for seg in queries]
    from concurrent.futures import ThreadPoolExecutor
    _lib.load()   # (resolve the library once, before the threads ask for it)
    with ThreadPoolExecutor(max_workers=min(PROBE_THREADS, len(queries))) as pool:
        return list(pool.map(probe, queries))


def phase1_norm(q, epsilon: float, alpha: float, beta: float, n: int, indexes):
    """Candidate intervals of a cNSM-ED query: (valid_positions [(left, right)], last_segment, plan).  Same deviations as
    phase1(): the index is scanned directly (the reference's incremental-visiting cache returns the same rows where its
    five cases apply and NOTHING where they do not — a range spanning a whole cached range matches no case, :258-309) and
    the wall-clock early termination (:410-421) is off."""
    by_w = dict(zip(WU_LIST, indexes))
    length = len(q)
    mean_q, std_q = query_statistics(q)
    queries = determine_query_plan(q, epsilon, {w: ix.stat for w, ix in by_w.items()},
                                   counts=lambda stat, wu, mean: _counts_norm(stat, wu, mean, epsilon, alpha, beta, mean_q, std_q))
    def probe(seg):   # CS_i of one segment: depends on the segment only, so the segments are probed side by side
        lo, hi = norm_mean_range(seg.mean, seg.wu, epsilon, alpha, beta, mean_q, std_q)
        parts = beta_partitions(seg, epsilon, alpha, beta, mean_q, std_q)
        return norm_sort_merge(scan_index_norm(by_w[seg.wu], seg, lo, to_round(hi), parts), 0)[0]
    probed = _probe_all(probe, queries)
    valid = np.zeros(0, dtype=NORM_IV)
    pre_length = 0
    for i, seg in enumerate(queries):
        delta_w = 0 if i == len(queries) - 1 else (queries[i + 1].order - seg.order) * WU_LIST[0]
        pre_length += seg.wu // WU_ALL[0]
        positions = probed[i]
        if i == 0:
            nxt = norm_first_segment(positions, seg.order, length, n, delta_w)
        else:
            nxt = norm_intersect(valid, positions, pre_length, length, mean_q, std_q, alpha, beta, delta_w)
        valid, _, _ = norm_sort_merge(nxt, 1)
        if len(valid) == 0:
            break
    merged, _, _ = norm_sort_merge(valid, 2)
    return [(int(l), int(r)) for l, r in zip(merged["left"], merged["right"])], queries[-1].order, queries


# ---------------------------------------------------------------- cNSM-DTW: phases 0 / 1 of K/NormQueryEngineDtw.java:190-455
def query_envelope_padded(q, rho: int):
    """:673-715: the raw query padded with rho copies of its first and last value, then the sliding maximum / minimum over
    2 rho + 1 points: U[i] = max q[i-rho .. i+rho], L[i] = min (indexes clamped into the query)."""
    v = np.asarray(q, dtype=np.float64)
    if rho <= 0:
        return v.copy(), v.copy()
    pad = np.concatenate([np.full(rho, v[0]), v, np.full(rho, v[-1])])
    win = np.lib.stride_tricks.sliding_window_view(pad, 2 * rho + 1)
    return win.min(axis=1), win.max(axis=1)


def norm_mean_range_dtw(mean_min: float, mean_max: float, wu: int, epsilon: float, alpha: float, beta: float, mean_q: float, std_q: float,
                        lo_shift=0.0, hi_shift=None):
    """:238-244 (whole beta range) and :259-265 (one beta partition), operation for operation."""
    if hi_shift is None:
        begin = 1.0 / alpha * mean_min + (1 - 1.0 / alpha) * mean_q - beta - 1.0 / alpha * epsilon * std_q / math.sqrt(wu)
        begin1 = alpha * mean_min + (1 - alpha) * mean_q - beta - alpha * epsilon * std_q / math.sqrt(wu)
        end = alpha * mean_max + (1 - alpha) * mean_q + beta + alpha * epsilon * std_q / math.sqrt(wu)
        end1 = 1.0 / alpha * mean_max + (1 - 1.0 / alpha) * mean_q + beta + 1.0 / alpha * epsilon * std_q / math.sqrt(wu)
    else:
        begin = 1.0 / alpha * mean_min + (1 - 1.0 / alpha) * mean_q - beta + lo_shift - 1.0 / alpha * epsilon * std_q / math.sqrt(wu)
        begin1 = alpha * mean_min + (1 - alpha) * mean_q - beta + lo_shift - alpha * epsilon * std_q / math.sqrt(wu)
        end = alpha * mean_max + (1 - alpha) * mean_q - beta + hi_shift + alpha * epsilon * std_q / math.sqrt(wu)
        end1 = 1.0 / alpha * mean_max + (1 - 1.0 / alpha) * mean_q - beta + hi_shift + 1.0 / alpha * epsilon * std_q / math.sqrt(wu)
    return min(begin, begin1), max(end, end1)


def _cumulative_counts(stat, begin: float, end: float):
    """The table lookups shared by every getCountsFromStatisticInfo (e.g. K/NormQueryEngineDtw.java:633-645)."""
    keys = _keys_of(stat)

    def search(key):
        return min(bisect.bisect_left(keys, key), len(stat) - 1)
    i = search(begin)
    lower1 = stat[i - 1][1] if i > 0 else 0
    lower2 = stat[i - 1][2] if i > 0 else 0
    i = search(end)
    upper1 = stat[i][1] if i > 0 else 0
    upper2 = stat[i][2] if i > 0 else 0
    return upper1 - lower1, upper2 - lower2


def scan_index_norm_dtw(idx, seg: RangeQuerySegment, lo_value: float, end: float, parts) -> np.ndarray:
    """scanIndex :802-833: like the ED engine's, plus the row's upper block sums (upper = the next row key)."""
    blocks = seg.wu // WU_ALL[0]

    def sums(key, upper):
        sq_lower = upper * upper if key < 0 else key * key
        sq_upper = key * key if upper < 0 else upper * upper
        return key * blocks, sq_lower * blocks, upper * blocks, sq_upper * blocks
    return _scan_rows_norm(idx, lo_value, end, parts, sums)


def phase1_norm_dtw(q, epsilon: float, rho: int, alpha: float, beta: float, n: int, indexes):
    """Candidate intervals of a cNSM-DTW query: (valid_positions, last_segment, plan); deviations as in phase1_norm()."""
    by_w = dict(zip(WU_LIST, indexes))
    length = len(q)
    mean_q, std_q = query_statistics(q)
    rng = lambda a, b, wu, lo=0.0, hi=None: norm_mean_range_dtw(a, b, wu, epsilon, alpha, beta, mean_q, std_q, lo, hi)

    def counts(stat, wu, mean_min, mean_max):   # getCountsFromStatisticInfo :622-646 (both ends through the plain toRound)
        lo, hi = rng(mean_min, mean_max, wu)
        return _cumulative_counts(stat, to_round(lo), to_round(hi))
    queries = determine_query_plan(q, epsilon, {w: ix.stat for w, ix in by_w.items()}, counts=counts, bounds=query_envelope_padded(q, rho))
    def probe(seg):
        lo, hi = rng(seg.mean_min, seg.mean_max, seg.wu)
        num = min(int(2.0 * beta / BETA_PARTITION_WIDTH), 64)   # :246-268
        parts = []
        for j in range(num):
            width = 2.0 * beta / num
            p_lo, p_hi = rng(seg.mean_min, seg.mean_max, seg.wu, width * j, width * (j + 1))
            parts.append((p_lo, to_round(p_hi)))
        return norm_sort_merge(scan_index_norm_dtw(by_w[seg.wu], seg, lo, to_round(hi), parts), 0)[0]
    probed = _probe_all(probe, queries)
    valid = np.zeros(0, dtype=NORM_IV)
    pre_length = 0
    for i, seg in enumerate(queries):
        delta_w = 0 if i == len(queries) - 1 else (queries[i + 1].order - seg.order) * WU_LIST[0]
        pre_length += seg.wu // WU_ALL[0]
        positions = probed[i]
        if i == 0:
            nxt = norm_first_segment(positions, seg.order, length, n, delta_w)
        else:
            nxt = norm_intersect(valid, positions, pre_length, length, mean_q, std_q, alpha, beta, delta_w, dtw=True)
        valid, _, _ = norm_sort_merge(nxt, 1)
        if len(valid) == 0:
            break
    merged, _, _ = norm_sort_merge(valid, 2)
    return [(int(l), int(r)) for l, r in zip(merged["left"], merged["right"])], queries[-1].order, queries


# ---------------------------------------------------------------- RSM-DTW: phases 0 / 1 of K/QueryEngineDtw.java:172-347
def scan_index_dtw(idx: IndexFile, seg: RangeQuerySegment, begin: float, end: float):
    """scanIndex :647-661 with getDistanceLowerBound :721-734: the row's mean range against the segment's mean range."""
    def bound(key, upper):
        if key > seg.mean_max:
            delta = (key - seg.mean_max) * (key - seg.mean_max)
        elif upper < seg.mean_min:
            delta = (seg.mean_min - upper) * (seg.mean_min - upper)
        else:
            delta = 0.0
        return seg.wu * delta
    return _scan_rows(idx, begin, end, bound)


def phase1_dtw(q, epsilon: float, rho: int, n: int, indexes):
    """Candidate intervals of an RSM-DTW query: (valid_positions, last_segment, plan).  The RSM-ED loop over segments that
    carry the mean range of the query's envelope; the interval algebra is the same (kvm_intervals_*)."""
    by_w = dict(zip(WU_LIST, indexes))

    def counts(stat, wu, mean_min, mean_max):   # getCountsFromStatisticInfo :471-491
        rng = epsilon / math.sqrt(wu)
        return _cumulative_counts(stat, to_round(mean_min - rng), to_round(mean_max + rng))
    queries = determine_query_plan(q, epsilon, {w: ix.stat for w, ix in by_w.items()}, counts=counts, bounds=query_envelope_padded(q, rho))
    return _rsm_loop(queries, epsilon, len(q), n, by_w, scan_index_dtw, lambda seg: (seg.mean_min, seg.mean_max), True)
