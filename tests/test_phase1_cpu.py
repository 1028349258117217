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
