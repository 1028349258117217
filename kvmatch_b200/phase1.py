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

    def row(self, i: int):
        """(key, [(left, right), ...]) of row i: IndexNode.parseBytesCompact, K/common/entity/IndexNode.java:108-128."""
        b = self.data[self.offsets[i]:self.offsets[i + 1]]
        key = struct.unpack(">d", b[:8])[0]
        v = memoryview(b)[8:]
        out = []
        idx = 0
        while idx < len(v):
            left = struct.unpack(">i", v[idx:idx + 4])[0]
            idx += 4
            count = int.from_bytes(v[idx:idx + 1], "big", signed=True) + 128
            idx += 1
            right = left + int.from_bytes(v[idx:idx + 1], "big", signed=True) + 128
            idx += 1
            out.append((left, right))
            for _ in range(count):
                left = right + int.from_bytes(v[idx:idx + 1], "big", signed=True) + 128
                right = left + int.from_bytes(v[idx + 1:idx + 2], "big", signed=True) + 128
                idx += 2
                out.append((left, right))
        return key, out

    def read_indexes(self, key_from: float, key_to: float):
        """readIndexes :64-83: rows with key_from <= key <= key_to (lowerBound / upperBound on the row keys)."""
        lo = bisect.bisect_left(self.keys, key_from)
        hi = bisect.bisect_right(self.keys, key_to) - 1
        return [self.row(i) for i in range(lo, hi + 1)] if lo < self.n_rows and hi >= 0 else []


# ---------------------------------------------------------------- interval algebra: the library's host functions
def _pack(ivs):
    k = len(ivs)
    lr = np.zeros((max(k, 1), 2), dtype=np.int32)
    eps = np.zeros(max(k, 1))
    for i, (l, r, e) in enumerate(ivs):
        lr[i] = (l, r)
        eps[i] = e
    return lr, eps, k


def _unpack(lr, eps, k):
    return [(int(lr[i, 0]), int(lr[i, 1]), float(eps[i])) for i in range(k)]


def sort_merge(ivs, mode: int):
    """kvm_intervals_sort_merge -> (intervals, cnt_disjoint, cnt_offsets)."""
    L = _lib.load()
    lr, eps, k = _pack(ivs)
    lo, eo = np.zeros((max(k, 1), 2), dtype=np.int32), np.zeros(max(k, 1))
    ko, cd, co = C.c_int64(), C.c_int64(), C.c_int64()
    rc = L.kvm_intervals_sort_merge(lr.ctypes.data, eps.ctypes.data, k, mode, lo.ctypes.data, eo.ctypes.data, max(k, 1),
                                    C.byref(ko), C.byref(cd), C.byref(co))
    if rc:
        raise _lib.KvmError(rc, "kvm_intervals_sort_merge")
    return _unpack(lo, eo, ko.value), cd.value, co.value


def intersect(cs, csi, eps2: float, delta_w: int):
    L = _lib.load()
    a, ae, k1 = _pack(cs)
    b, be, k2 = _pack(csi)
    cap = max(k1 + k2, 1)
    lo, eo = np.zeros((cap, 2), dtype=np.int32), np.zeros(cap)
    ko, me = C.c_int64(), C.c_double()
    rc = L.kvm_intervals_intersect(a.ctypes.data, ae.ctypes.data, k1, b.ctypes.data, be.ctypes.data, k2, eps2, delta_w,
                                   lo.ctypes.data, eo.ctypes.data, cap, C.byref(ko), C.byref(me))
    if rc:
        raise _lib.KvmError(rc, "kvm_intervals_intersect")
    return _unpack(lo, eo, ko.value), me.value


def first_segment(pos, order: int, length: int, n: int, delta_w: int):
    L = _lib.load()
    a, ae, k = _pack(pos)
    lo, eo = np.zeros((max(k, 1), 2), dtype=np.int32), np.zeros(max(k, 1))
    ko, me = C.c_int64(), C.c_double()
    rc = L.kvm_intervals_first_segment(a.ctypes.data, ae.ctypes.data, k, order, WU_LIST[0], length, n, delta_w, lo.ctypes.data,
                                       eo.ctypes.data, max(k, 1), C.byref(ko), C.byref(me))
    if rc:
        raise _lib.KvmError(rc, "kvm_intervals_first_segment")
    return _unpack(lo, eo, ko.value), me.value


# ---------------------------------------------------------------- phase 0: determineQueryPlan (K/QueryEngine.java:398-503)
@dataclass
class QuerySegment:
    mean: float
    order: int
    count: int
    wu: int


def _counts(stat, wu: int, mean: float, epsilon: float):
    """getCountsFromStatisticInfo :382-399 on the cumulative table of width wu."""
    keys = [t[0] for t in stat]
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


def determine_query_plan(q, epsilon: float, stats):
    """stats: {width: cumulative statistic table} for the widths of WU_LIST."""
    w0 = WU_ALL[0]
    m = len(q) // w0
    sums, ex = [], 0.0
    for i, v in enumerate(q):
        ex += float(v)
        if (i + 1) % w0 == 0:
            sums.append(ex)
            ex = 0.0
    prefix = [0.0] * m
    prefix[0] = sums[0]
    for i in range(1, m):
        prefix[i] = prefix[i - 1] + sums[i]
    total100 = stats[100][-1][1]
    cost, cost2 = {}, {}

    def get_cost(l, r):
        if (l, r) not in cost:
            use = w0 * (r - l + 1)
            mean = (prefix[r] - (prefix[l - 1] if l > 0 else 0.0)) / use
            c1, _ = _counts(stats[use], use, mean, epsilon)
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
        mean = (prefix[r] - (prefix[l - 1] if l > 0 else 0.0)) / use
        get_cost(l, r)
        queries.append(QuerySegment(mean, l + 1, cost2[(l, r)], use))
        index -= pre[index][i]
    queries.sort(key=lambda s: s.count)   # ENABLE_QUERY_REORDERING (stable sort)
    return queries


# ---------------------------------------------------------------- phase 1 (K/QueryEngine.java:185-334)
def scan_index(idx: IndexFile, seg: QuerySegment, begin: float, end: float):
    """scanIndex :505-521 with getDistanceLowerBound :383-396: positions (left, right, wu * lower bound)."""
    out = []
    for key, positions in idx.read_indexes(begin, end + 0.01):
        upper = to_upper_stat(key, idx.stat_keys)
        if key > seg.mean:
            delta = (key - seg.mean) * (key - seg.mean)
        elif upper < seg.mean:
            delta = (seg.mean - upper) * (seg.mean - upper)
        else:
            delta = 0.0
        out.extend((l, r, seg.wu * delta) for l, r in positions)
    return out


def phase1(q, epsilon: float, n: int, indexes):
    """Candidate intervals of an RSM-ED query: (valid_positions [(left, right)], last_segment, plan).  `indexes` = one
    IndexFile per width of WU_LIST."""
    by_w = dict(zip(WU_LIST, indexes))
    length = len(q)
    queries = determine_query_plan(q, epsilon, {w: ix.stat for w, ix in by_w.items()})
    valid = []
    last_min = 0.0
    range0 = epsilon * epsilon
    for i, seg in enumerate(queries):
        delta_w = 0 if i == len(queries) - 1 else (queries[i + 1].order - seg.order) * WU_LIST[0]
        ix = by_w[seg.wu]
        rng = math.sqrt((range0 - last_min) / seg.wu)
        begin = to_round_stat(seg.mean - rng, ix.stat_keys)
        end = to_round(seg.mean + rng)
        positions, _, _ = sort_merge(scan_index(ix, seg, begin, end), 0)
        if i == 0:
            nxt, last_min = first_segment(positions, seg.order, length, n, delta_w)
        else:
            nxt, last_min = intersect(valid, positions, range0, delta_w)
        valid, _, _ = sort_merge(nxt, 1)
        if not valid:
            break
    last_segment = queries[-1].order
    merged, _, _ = sort_merge(valid, 2)
    return [(l, r) for l, r, _ in merged], last_segment, queries
