"""CPU tests of the phase-1 tail: the library's host functions (kvm_intervals_*, csrc/phase1.hpp) against the pure-Python
restatement in oracle/phase1_oracle.py, and the index file reader against the writer (no GPU: host-only entry points)."""
import numpy as np
import pytest

from kvmatch_b200 import _lib, phase1
from oracle import phase1_oracle as po


def random_intervals(rng, k, span=100_000, width=300):
    lefts = rng.integers(1, span, size=k)
    out = []
    for l in lefts:
        out.append((int(l), int(l) + int(rng.integers(0, width)), float(rng.choice([0.0, 0.5, 3.25, 10.0 * rng.random()]))))
    return out


def disjoint_sorted(rng, k, span=100_000):
    cuts = np.sort(rng.choice(np.arange(1, span), size=2 * k, replace=False))
    return [(int(cuts[2 * i]), int(cuts[2 * i + 1]) - 1 if cuts[2 * i + 1] - 1 >= cuts[2 * i] else int(cuts[2 * i]),
             float(rng.random() * 8)) for i in range(k)]


@pytest.mark.parametrize("seed", range(8))
def test_sort_merge_modes(seed):
    rng = np.random.default_rng(seed)
    for k in (0, 1, 2, 17, 400, 5000):
        ivs = random_intervals(rng, k, span=20_000 if seed % 2 else 200_000)
        got, cd, co = phase1.sort_merge(ivs, 0)
        assert got == po.sort_but_not_merge(ivs)
        got, cd, co = phase1.sort_merge(ivs, 1)
        exp, ed, eo = po.sort_but_not_merge(ivs, count=True)
        assert got == exp and (cd, co) == (ed, eo)
        got, _, _ = phase1.sort_merge(ivs, 2)
        assert got == po.sort_and_merge(ivs)
        # the merged list is what phase 2 takes: sorted, disjoint and non-adjacent
        for a, b in zip(got, got[1:]):
            assert a[1] + 1 < b[0]


@pytest.mark.parametrize("seed", range(6))
def test_intersect_and_first_segment(seed):
    rng = np.random.default_rng(100 + seed)
    for k1, k2 in ((0, 5), (5, 0), (40, 60), (1500, 900)):
        cs, csi = disjoint_sorted(rng, k1), disjoint_sorted(rng, k2)
        for eps2, dw in ((1.0, 0), (9.0, 25), (100.0, -50)):
            got, gm = phase1.intersect(cs, csi, eps2, dw)
            exp, em = po.intersect(cs, csi, eps2, dw)
            assert got == exp and gm == em
    pos = disjoint_sorted(rng, 300, span=50_000)
    for order, length, n, dw in ((1, 512, 50_000, 0), (7, 1024, 49_000, 75), (3, 8192, 50_400, -25)):
        got, gm = phase1.first_segment(pos, order, length, n, dw)
        exp, em = po.first_segment(pos, order, 25, length, n, dw)
        assert got == exp and gm == em


def test_index_reader_round_trip():
    """The reader parses what the library's writer produces: keys, positions and the cumulative statistic table."""
    rng = np.random.default_rng(5)
    keys, first, last = [], [], []
    loc = 1
    while loc < 60_000:
        run = int(rng.integers(1, 255))
        keys.append(round(float(rng.integers(-40, 40)) * 0.05, 10))
        first.append(loc)
        last.append(loc + run - 1)
        loc += run
    keys = [phase1.to_round(k) for k in keys]
    image, info = _lib.index_image_from_runs(np.array(keys), np.array(first, dtype=np.int32), np.array(last, dtype=np.int32))
    ix = phase1.IndexFile(image)
    assert ix.n_rows == info.n_rows and ix.keys == sorted(ix.keys)
    covered = sorted(p for i in range(ix.n_rows) for p in ix.row(i)[1])
    assert sum(r - l + 1 for l, r in covered) == info.n_offsets == loc - 1
    for (l1, r1), (l2, r2) in zip(covered, covered[1:]):
        assert r1 < l2
    assert [t[0] for t in ix.stat] == ix.keys and ix.stat[-1][2] == info.n_offsets and ix.stat[-1][1] == info.n_intervals
    lo, hi = ix.keys[len(ix.keys) // 4], ix.keys[3 * len(ix.keys) // 4]
    rows = ix.read_indexes(lo, hi)
    assert [k for k, _ in rows] == [k for k in ix.keys if lo <= k <= hi]
    assert ix.read_indexes(ix.keys[-1] + 1.0, ix.keys[-1] + 2.0) == []


# ---------------------------------------------------------------- cNSM variant: kvm_norm_intervals_* against the restatement
def _norm_tuples(arr):
    return [(int(a["left"]), int(a["right"]), float(a["ex"]), float(a["ex2"]), float(a["exu"]), float(a["ex2u"]), int(a["bp"])) for a in arr]


def _norm_array(tuples):
    out = np.zeros(len(tuples), dtype=phase1.NORM_IV)
    for i, t in enumerate(tuples):
        out[i] = t
    return out


def random_norm_intervals(rng, k, span=100_000, width=300, disjoint=False, upper=True):
    sums = [(-12.5, 160.0), (0.0, 0.0), (3.0, 9.5), (3.0, 9.5000000000000018), (-0.0, 0.0), (40.0, 1700.0)]
    if disjoint:
        base = disjoint_sorted(rng, k, span)
        lr = [(l, r) for l, r, _ in base]
    else:
        lefts = rng.integers(1, span, size=k)
        lr = [(int(l), int(l) + int(rng.integers(0, width))) for l in lefts]
    out = []
    for l, r in lr:
        ex, ex2 = sums[int(rng.integers(0, len(sums)))] if rng.random() < 0.7 else (float(rng.normal() * 30), float(rng.random() * 3000))
        bp = int(rng.choice([0, 1, 1, 3, 6, -(1 << 31), (1 << 40) | 1]))
        exu, ex2u = (ex + float(rng.integers(0, 3)), ex2 + float(rng.integers(0, 40))) if upper else (0.0, 0.0)
        out.append((l, r, ex, ex2, exu, ex2u, bp))
    return out


@pytest.mark.parametrize("seed", range(8))
def test_norm_sort_merge_modes(seed):
    rng = np.random.default_rng(300 + seed)
    for k in (0, 1, 2, 23, 600, 4000):
        ivs = random_norm_intervals(rng, k, span=15_000 if seed % 2 else 300_000)
        arr = _norm_array(ivs)
        got, _, _ = phase1.norm_sort_merge(arr, 0)
        assert _norm_tuples(got) == po.norm_sort_but_not_merge(ivs)
        got, cd, co = phase1.norm_sort_merge(arr, 1)
        exp, ed, eo = po.norm_sort_but_not_merge(ivs, count=True)
        assert _norm_tuples(got) == exp and (cd, co) == (ed, eo)
        got, _, _ = phase1.norm_sort_merge(arr, 2)
        assert _norm_tuples(got) == po.norm_sort_and_merge(ivs)
        for a, b in zip(got, got[1:]):
            assert a["right"] + 1 < b["left"]


def test_norm_merge_needs_bit_equal_sums():
    """Neighbours (left - 1 == end) merge only when both sums are the same double (Double.compare == 0: 0.0 and -0.0 differ)."""
    a = (10, 19, 3.0, 9.5, 4.0, 11.0, 1)
    for b, merged in (((20, 29, 3.0, 9.5, 5.0, 10.0, 2), True), ((20, 29, 3.0, 9.5000000000000018, 4.0, 11.0, 2), False),
                      ((20, 29, 3.0000000000000004, 9.5, 4.0, 11.0, 2), False), ((19, 29, 7.0, 1.0, 8.0, 2.0, 2), True)):
        got, _, _ = phase1.norm_sort_merge(_norm_array([a, b]), 0)
        assert len(got) == (1 if merged else 2)
        if merged:
            assert _norm_tuples(got)[0] == (10, 29, min(a[2], b[2]), min(a[3], b[3]), min(a[4], b[4]), min(a[5], b[5]), 3)
    z = [(1, 5, 0.0, 0.0, 0.0, 0.0, 1), (6, 9, -0.0, 0.0, 0.0, 0.0, 1)]
    assert len(phase1.norm_sort_merge(_norm_array(z), 0)[0]) == 2


@pytest.mark.parametrize("seed", range(6))
def test_norm_intersect_and_first_segment(seed):
    rng = np.random.default_rng(400 + seed)
    for k1, k2 in ((0, 5), (5, 0), (40, 60), (1500, 900)):
        cs, csi = random_norm_intervals(rng, k1, disjoint=True), random_norm_intervals(rng, k2, disjoint=True)
        for pre, length, mq, sq, alpha, beta, dw in ((3, 512, 1.0, 2.0, 1.5, 5.0, 25), (12, 1024, -20.0, 0.7, 1.1, 1.0, -50),
                                                     (4, 100, 0.0, 10.0, 2.0, 100.0, 0), (2, 256, 5.0, float("nan"), 1.5, 5.0, 0)):
            for dtw in (False, True):
                got = phase1.norm_intersect(_norm_array(cs), _norm_array(csi), pre, length, mq, sq, alpha, beta, dw, dtw)
                assert _norm_tuples(got) == po.norm_intersect(cs, csi, pre, 25, length, mq, sq, alpha, beta, dw, dtw)
    pos = random_norm_intervals(rng, 300, span=50_000, disjoint=True)
    for order, length, n, dw in ((1, 512, 50_000, 0), (7, 1024, 49_000, 75), (3, 8192, 50_400, -25)):
        got = phase1.norm_first_segment(_norm_array(pos), order, length, n, dw)
        assert _norm_tuples(got) == po.norm_first_segment(pos, order, 25, length, n, dw)


def test_beta_partition_bits_follow_java_int_shift():
    """`partitions |= 1 << idx` shifts an int: index 31 sign-extends, 32 wraps to bit 0 (K/NormQueryEngine.java:692)."""
    assert phase1._java_int_shl1(0) == 1 and phase1._java_int_shl1(30) == 1 << 30
    assert phase1._java_int_shl1(31) == -(1 << 31) and phase1._java_int_shl1(32) == 1 and phase1._java_int_shl1(63) == -(1 << 31)


@pytest.fixture(scope="module")
def small_world():
    from kvmatch_b200 import datagen
    from oracle import kvm_oracle
    s = datagen.generate(120_000, seed=11)
    indexes = [phase1.IndexFile(kvm_oracle.index_file_image(s, w)[0]) for w in phase1.WU_LIST]
    return s, indexes


@pytest.mark.parametrize("off,length,eps,alpha,beta", [(30_000, 512, 3.0, 1.5, 5.0), (77_777, 1024, 6.0, 1.2, 5.0), (5_000, 256, 2.0, 2.0, 10.0),
                                                       (100_000, 400, 4.0, 1.5, 20.0)])
def test_cnsm_phase1_has_no_false_dismissals(small_world, off, length, eps, alpha, beta):
    """Phases 0 / 1 of the cNSM-ED engine over real index files, phase 2 by the oracle: the index-pruned answers equal the
    full scan's (KV-match_DP guarantees no false dismissals) and the index prunes."""
    from oracle import kvm_oracle
    s, indexes = small_world
    n = len(s)
    q = s[off - 1:off - 1 + length].copy()
    valid, last_segment, plan = phase1.phase1_norm(q, eps, alpha, beta, n, indexes)
    assert valid and 1 <= last_segment <= length // 25
    for (l1, r1), (l2, r2) in zip(valid, valid[1:]):
        assert r1 + 1 < l2
    shift = (last_segment - 1) * 25
    full = kvm_oracle.verify_cnsm_ed(s, q, eps, alpha, beta, [(1, n - length + 1)])
    pruned = kvm_oracle.verify_cnsm_ed(s, q, eps, alpha, beta, valid, shift)
    assert off in full.offsets.tolist()
    assert pruned.offsets.tolist() == full.offsets.tolist()
    assert sum(r - l + 1 for l, r in valid) < n


def test_cnsm_phase1_small_beta_quirk(small_world):
    """beta < 5 gives (int)(2 beta / 10) = 0 beta partitions: every row's bit set is empty and the second segment's
    intersection drops everything.  The reference's behaviour (ENABLE_BETA_PARTITION), kept."""
    s, indexes = small_world
    q = s[40_000:40_000 + 512].copy()
    valid, _, plan = phase1.phase1_norm(q, 3.0, 1.5, 2.0, len(s), indexes)
    assert len(plan) > 1 and valid == []


def test_query_envelope_padded_matches_clamped_windows():
    rng = np.random.default_rng(9)
    q = rng.normal(size=300)
    for rho in (0, 1, 7, 40):
        lo, up = phase1.query_envelope_padded(q, rho)
        for i in (0, 1, rho, 150, 299 - rho, 298, 299):
            a, b = max(0, i - rho), min(len(q) - 1, i + rho)
            assert lo[i] == q[a:b + 1].min() and up[i] == q[a:b + 1].max()


@pytest.mark.parametrize("off,length,eps,rho,alpha,beta", [(30_000, 512, 3.0, 25, 1.5, 5.0), (77_777, 256, 2.0, 12, 1.2, 5.0),
                                                           (5_000, 400, 4.0, 20, 2.0, 10.0)])
def test_cnsm_dtw_phase1_has_no_false_dismissals(small_world, off, length, eps, rho, alpha, beta):
    """The cNSM-DTW engine's phases 0 / 1 (segments carry the mean RANGE of the query's envelope, intervals carry lower and
    upper block sums), phase 2 by the oracle: index-pruned answers = full-scan answers."""
    from oracle import kvm_oracle
    s, indexes = small_world
    n = len(s)
    q = s[off - 1:off - 1 + length].copy()
    valid, last_segment, plan = phase1.phase1_norm_dtw(q, eps, rho, alpha, beta, n, indexes)
    assert valid and all(isinstance(seg, phase1.RangeQuerySegment) and seg.mean_min <= seg.mean_max for seg in plan)
    for (l1, r1), (l2, r2) in zip(valid, valid[1:]):
        assert r1 + 1 < l2
    shift = (last_segment - 1) * 25
    full = kvm_oracle.verify_cnsm_dtw(s, q, eps, rho, alpha, beta, [(1, n - length + 1)])
    pruned = kvm_oracle.verify_cnsm_dtw(s, q, eps, rho, alpha, beta, valid, shift)
    assert off in full.offsets.tolist()
    assert pruned.offsets.tolist() == full.offsets.tolist()
    # (the running sums restart at every interval: the same window's statistics differ in the last bits between the two lists)
    assert np.allclose(pruned.distances, full.distances, rtol=1e-7, atol=1e-6)
    assert sum(r - l + 1 for l, r in valid) < n
    # the ED engine's candidates for the same query are a subset of the DTW engine's (rho = 0 would make them equal)
    ed_valid, ed_last, _ = phase1.phase1_norm(q, eps, alpha, beta, n, indexes)
    assert sum(r - l + 1 for l, r in ed_valid) <= sum(r - l + 1 for l, r in valid) or ed_last != last_segment


@pytest.mark.parametrize("off,length,eps,rho", [(30_000, 512, 12.0, 25), (77_777, 256, 6.0, 12), (5_000, 1000, 30.0, 50)])
def test_rsm_dtw_phase1_has_no_false_dismissals(small_world, off, length, eps, rho):
    from oracle import kvm_oracle
    s, indexes = small_world
    n = len(s)
    q = s[off - 1:off - 1 + length].copy()
    valid, last_segment, plan = phase1.phase1_dtw(q, eps, rho, n, indexes)
    assert valid
    shift = (last_segment - 1) * 25
    full = kvm_oracle.verify_dtw(s, q, eps, rho, [(1, n - length + 1)])
    pruned = kvm_oracle.verify_dtw(s, q, eps, rho, valid, shift)
    assert off in full.offsets.tolist()
    assert pruned.offsets.tolist() == full.offsets.tolist() and pruned.distances.tolist() == full.distances.tolist()
    n_cand = sum(r - l + 1 for l, r in valid)
    assert n_cand < 0.5 * n


# ---------------------------------------------------------------- per-shard index layout
@pytest.fixture(scope="module")
def sharded_world(small_world):
    from oracle import kvm_oracle
    s, single = small_world
    sharded = []
    for w in phase1.WU_LIST:
        k, f, l = kvm_oracle.window_mean_runs(s, w)
        pieces = phase1.split_runs(k, f, l, phase1.shard_ranges(int(l[-1]), 3))
        sharded.append(phase1.ShardedIndexFile([_lib.index_image_from_runs(*p)[0] for p in pieces]))
    return s, single, sharded


def test_sharded_index_holds_the_same_positions(sharded_world):
    s, single, sharded = sharded_world
    for one, many in zip(single, sharded):
        assert len(many.parts()) == 3
        cover_one = sorted(p for i in range(one.n_rows) for p in one.row(i)[1])
        cover_many = sorted(p for part in many.parts() for i in range(part.n_rows) for p in part.row(i)[1])
        total = sum(r - l + 1 for l, r in cover_one)
        assert sum(r - l + 1 for l, r in cover_many) == total == many.stat[-1][2] == one.stat[-1][2]
        for (l1, r1), (l2, r2) in zip(cover_many, cover_many[1:]):
            assert r1 < l2                                    # no window start in two files
        # every part covers one contiguous range of window starts
        spans = [(min(p[0] for i in range(part.n_rows) for p in part.row(i)[1]), max(p[1] for i in range(part.n_rows) for p in part.row(i)[1]))
                 for part in many.parts()]
        assert spans[0][0] == 1 and all(a[1] + 1 == b[0] for a, b in zip(spans, spans[1:])) and spans[-1][1] == total
        # the aggregated table is cumulative in both counts
        assert all(a[1] <= b[1] and a[2] <= b[2] for a, b in zip(many.stat, many.stat[1:]))


@pytest.mark.parametrize("off,length,eps", [(30_000, 512, 3.0), (77_777, 1024, 6.0), (41_000, 256, 2.0)])
def test_sharded_index_queries_equal_full_scan(sharded_world, off, length, eps):
    """All four engines' phase 1 over the per-shard layout: no false dismissals, and never fewer candidates than what the
    true answers need (a query cut across a shard boundary at 40 000 / 80 000 included)."""
    from oracle import kvm_oracle
    s, single, sharded = sharded_world
    n = len(s)
    q = s[off - 1:off - 1 + length].copy()
    v, last, _ = phase1.phase1(q, eps * 3, n, sharded)
    full = kvm_oracle.verify_ed(s, q, eps * 3, [(1, n - length + 1)])
    got = kvm_oracle.verify_ed(s, q, eps * 3, v, (last - 1) * 25)
    assert got.offsets.tolist() == full.offsets.tolist() and got.distances.tolist() == full.distances.tolist() and off in got.offsets.tolist()
    v, last, _ = phase1.phase1_norm(q, eps, 1.5, 5.0, n, sharded)
    full = kvm_oracle.verify_cnsm_ed(s, q, eps, 1.5, 5.0, [(1, n - length + 1)])
    assert kvm_oracle.verify_cnsm_ed(s, q, eps, 1.5, 5.0, v, (last - 1) * 25).offsets.tolist() == full.offsets.tolist()
    rho = length // 20
    v, last, _ = phase1.phase1_dtw(q, eps * 3, rho, n, sharded)
    full = kvm_oracle.verify_dtw(s, q, eps * 3, rho, [(1, n - length + 1)])
    assert kvm_oracle.verify_dtw(s, q, eps * 3, rho, v, (last - 1) * 25).offsets.tolist() == full.offsets.tolist()
    v, last, _ = phase1.phase1_norm_dtw(q, eps, rho, 1.5, 5.0, n, sharded)
    full = kvm_oracle.verify_cnsm_dtw(s, q, eps, rho, 1.5, 5.0, [(1, n - length + 1)])
    assert kvm_oracle.verify_cnsm_dtw(s, q, eps, rho, 1.5, 5.0, v, (last - 1) * 25).offsets.tolist() == full.offsets.tolist()


def test_split_runs_cuts_straddling_runs():
    keys = np.array([0.5, 1.0, 0.5, 2.0])
    first = np.array([1, 11, 31, 41], dtype=np.int32)
    last = np.array([10, 30, 40, 50], dtype=np.int32)
    a, b, c = phase1.split_runs(keys, first, last, [(1, 20), (21, 35), (36, 50)])
    assert (a[0].tolist(), a[1].tolist(), a[2].tolist()) == ([0.5, 1.0], [1, 11], [10, 20])
    assert (b[0].tolist(), b[1].tolist(), b[2].tolist()) == ([1.0, 0.5], [21, 31], [30, 35])
    assert (c[0].tolist(), c[1].tolist(), c[2].tolist()) == ([0.5, 2.0], [36, 41], [40, 50])
    assert phase1.shard_ranges(50, 3) == [(1, 17), (18, 34), (35, 50)] and phase1.shard_ranges(2, 3) == [(1, 1), (2, 2)]


# ---------------------------------------------------------------- adversarial small cases (hypothesis): ties on `left`,
# adjacency, nested intervals, signed zeros and NaN sums, bit 31 / bit 63 partition sets
from hypothesis import given, settings, strategies as st

_sum = st.sampled_from([0.0, -0.0, 1.5, 1.5000000000000002, -7.25, 1e300, float("nan")])
_bp = st.sampled_from([0, 1, 2, 3, -(1 << 31), (1 << 62), -1])
_norm_iv = st.tuples(st.integers(1, 60), st.integers(0, 12), _sum, _sum, _sum, _sum, _bp).map(
    lambda t: (t[0], t[0] + t[1], t[2], t[3], t[4], t[5], t[6]))


def _same(a, b):
    """Tuple lists equal with NaN == NaN and -0.0 != 0.0 (what Double.compare sees)."""
    import struct
    key = lambda lst: [tuple(struct.pack(">d", x) if isinstance(x, float) else x for x in t) for t in lst]
    norm = lambda lst: [tuple((float("nan") if isinstance(x, float) and x != x else x) for x in t) for t in lst]
    ka, kb = key(norm(a)), key(norm(b))
    return ka == kb


@settings(max_examples=300, deadline=None)
@given(st.lists(_norm_iv, max_size=14), st.integers(0, 2))
def test_norm_sort_merge_small_adversarial(ivs, mode):
    got, cd, co = phase1.norm_sort_merge(_norm_array(ivs), mode)
    if mode == 2:
        exp, ed, eo = po.norm_sort_and_merge(ivs), None, None
    else:
        exp, ed, eo = po.norm_sort_but_not_merge(ivs, count=True)
    assert _same(_norm_tuples(got), exp)
    if mode == 1:
        assert (cd, co) == (ed, eo)


@settings(max_examples=300, deadline=None)
@given(st.lists(_norm_iv, max_size=10), st.lists(_norm_iv, max_size=10), st.integers(1, 6), st.sampled_from([-3.0, 0.0, 2.0]),
       st.sampled_from([0.5, 2.0, float("nan")]), st.sampled_from([1.0, 1.5]), st.sampled_from([0.0, 5.0]), st.integers(-50, 50), st.booleans())
def test_norm_intersect_small_adversarial(cs, csi, pre, mean_q, std_q, alpha, beta, dw, dtw):
    # the engine hands over lists that are sorted by left and internally disjoint: make them so
    def disjoint(lst):
        out, end = [], 0
        for t in sorted(lst, key=lambda t: t[0]):
            if t[0] > end:
                out.append(t)
                end = t[1]
        return out
    cs, csi = disjoint(cs), disjoint(csi)
    got = phase1.norm_intersect(_norm_array(cs), _norm_array(csi), pre, 400, mean_q, std_q, alpha, beta, dw, dtw)
    assert _same(_norm_tuples(got), po.norm_intersect(cs, csi, pre, 25, 400, mean_q, std_q, alpha, beta, dw, dtw))


_rsm_iv = st.tuples(st.integers(1, 60), st.integers(0, 12), st.sampled_from([0.0, 0.25, 0.999, 1.0, 1.25, 7.5, 100.0])).map(
    lambda t: (t[0], t[0] + t[1], t[2]))


@settings(max_examples=300, deadline=None)
@given(st.lists(_rsm_iv, max_size=14), st.integers(0, 2))
def test_rsm_sort_merge_small_adversarial(ivs, mode):
    """Ties on `left`, adjacency with |eps difference| around the merge threshold 1 (K/QueryEngine.java:607), nested intervals."""
    got, cd, co = phase1.sort_merge(ivs, mode)
    if mode == 2:
        assert got == po.sort_and_merge(ivs)
    else:
        exp, ed, eo = po.sort_but_not_merge(ivs, count=True)
        assert got == exp
        if mode == 1:
            assert (cd, co) == (ed, eo)


@settings(max_examples=300, deadline=None)
@given(st.lists(_rsm_iv, max_size=10), st.lists(_rsm_iv, max_size=10), st.sampled_from([0.5, 1.25, 8.0, 200.0]), st.integers(-50, 50))
def test_rsm_intersect_small_adversarial(cs, csi, eps2, dw):
    def disjoint(lst):
        out, end = [], 0
        for t in sorted(lst, key=lambda t: t[0]):
            if t[0] > end:
                out.append(t)
                end = t[1]
        return out
    cs, csi = disjoint(cs), disjoint(csi)
    got, gm = phase1.intersect(cs, csi, eps2, dw)
    exp, em = po.intersect(cs, csi, eps2, dw)
    assert got == exp and gm == em
