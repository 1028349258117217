"""CPU tests that pin oracle/ as far as the reference allows (it ships no tests or golden vectors):
the reference's self-checks, a second independent restatement (tests/refimpl.py), and invariants."""
import math
import os
import struct

import numpy as np
import pytest

from kvmatch_b200 import datagen
from tests import refimpl

RNG = np.random.default_rng(7)


def as_pairs(res):
    return list(zip(res.offsets.tolist(), res.distances.tolist()))


# ---- MeanIntervalUtils.toRound: doc-comment examples (K/utils/MeanIntervalUtils.java:44-47, x10 scaled) ----
@pytest.mark.parametrize("v,expect", [(0.19, 0.15), (0.14, 0.10), (0.15, 0.15), (-0.19, -0.20), (-0.14, -0.15),
                                      (-0.15, -0.15), (0.0, 0.0), (12.3456, 12.30), (-7.77, -7.80)])
def test_to_round_examples(oracle, v, expect):
    got = oracle.to_round(v)
    assert got == pytest.approx(expect, abs=1e-12)
    assert got == refimpl.to_round(v)  # bit-identical to the second restatement


def test_to_round_random_bitexact(oracle):
    for v in RNG.uniform(-600, 600, 5000):
        assert oracle.to_round(float(v)) == refimpl.to_round(float(v))


# ---- Lemire envelope == clamped sliding min/max (K/utils/DtwUtils.java:50-91 + CircularArray) ----
@pytest.mark.parametrize("trial", range(40))
def test_lemire_equals_clamped_minmax(oracle, trial):
    rng = np.random.default_rng(100 + trial)
    r = int(rng.integers(0, 12))
    n = int(rng.integers(r + 1, 80))
    t = rng.integers(-4, 5, n).astype(np.float64) if trial % 2 else rng.normal(size=n)
    lo, up = oracle.lower_upper_lemire(t, r)
    elo, eup = refimpl.envelope(t.tolist(), r)
    assert lo.tolist() == elo and up.tolist() == eup


def test_lemire_short_input_is_a_reference_throw(oracle):
    with pytest.raises(oracle.ReferenceThrows):
        oracle.lower_upper_lemire(np.zeros(3), 5)


# ---- DTW invariants ----
def test_dtw_rho0_is_squared_ed(oracle):
    a = RNG.normal(size=64)
    b = RNG.normal(size=64)
    s = 0.0
    for x, y in zip(a, b):
        s += (x - y) * (x - y)
    assert oracle.dtw(a, b, np.zeros(64), 0, 1e300) == s


@pytest.mark.parametrize("trial", range(20))
def test_dtw_matches_full_matrix_and_bounds(oracle, trial):
    rng = np.random.default_rng(200 + trial)
    m = int(rng.integers(8, 60))
    r = int(rng.integers(0, m // 2 + 1))
    a = np.cumsum(rng.normal(size=m))
    b = np.cumsum(rng.normal(size=m))
    full = refimpl.dtw_full(a.tolist(), b.tolist(), r)
    assert oracle.dtw(a, b, np.zeros(m), r, 1e300) == full
    # lower bounds (mean 0, std 1): Kim and both Keoghs never exceed the band DTW
    t2 = np.concatenate([a, a])
    order = np.arange(m, dtype=np.int32)
    assert oracle.lb_kim(t2, b, 0, m, 0.0, 1.0, 1e300) <= full * (1 + 1e-12) + 1e-12
    lo, up = oracle.lower_upper_lemire(b, r)
    lbk, cb1 = oracle.lb_keogh(order, t2, up, lo, 0, m, 0.0, 1.0, 1e300)
    assert lbk <= full * (1 + 1e-12) + 1e-12
    dlo, dup = oracle.lower_upper_lemire(a, r)
    lbk2, cb2 = oracle.lb_keogh_data(order, b, 0, dlo, dup, m, 0.0, 1.0, 1e300)
    assert lbk2 <= full * (1 + 1e-12) + 1e-12
    # early abandoning with a true cumulative bound never changes an accepted distance
    cb = np.cumsum(cb1[::-1])[::-1].copy()
    bsf = full * 1.5 + 1.0
    assert oracle.dtw(a, b, cb, r, bsf) == full


# ---- the four phase-2 loops vs the second restatement ----
@pytest.fixture(scope="module")
def small_series():
    return datagen.generate(6000, seed=11)


def some_intervals(n, m, rng, k=6, span=40):
    lefts = np.sort(rng.choice(np.arange(1, n - m - span), size=k, replace=False))
    out = []
    end = 0
    for l in lefts:
        l = max(int(l), end + 2)
        r = l + int(rng.integers(0, span))
        out.append((l, r))
        end = r
    return out


@pytest.mark.parametrize("m,eps,shift", [(32, 3.0, 0), (64, 12.0, 25), (100, 40.0, 50)])
def test_verify_ed_matches_second_restatement(oracle, small_series, m, eps, shift):
    rng = np.random.default_rng(m)
    s = small_series
    off = 1234
    q = s[off - 1:off - 1 + m].copy()
    iv = some_intervals(len(s), m, rng) + [(off + shift - 3, off + shift + 3)]
    iv.sort()
    got = oracle.verify_ed(s, q, eps, iv, shift)
    exp = refimpl.verify_ed(s.tolist(), q.tolist(), eps, iv, shift)
    assert as_pairs(got) == exp
    assert (off, 0.0) in exp  # self match, distance exactly 0.0 (README.md:78-84)
    assert got.cnt_candidate == sum(r - l + 1 for l, r in iv)


@pytest.mark.parametrize("m,eps,alpha,beta", [(32, 2.0, 1.5, 5.0), (64, 4.0, 2.0, 1.0), (128, 8.0, 1.1, 10.0)])
def test_verify_cnsm_ed_matches_second_restatement(oracle, small_series, m, eps, alpha, beta):
    s = small_series
    off = 2500
    q = s[off - 1:off - 1 + m].copy()
    iv = [(1, 700), (900, 905), (2400, 2600), (5000, len(s))]
    got = oracle.verify_cnsm_ed(s, q, eps, alpha, beta, iv, 0)
    exp = refimpl.verify_cnsm_ed(s.tolist(), q.tolist(), eps, alpha, beta, iv, 0)
    assert as_pairs(got) == exp
    assert off in got.offsets.tolist()
    # the self match sits mid-chain, so its running statistics differ from the query's by rounding
    assert got.distances[got.offsets.tolist().index(off)] < 1e-5


@pytest.mark.parametrize("m,rho,eps", [(24, 2, 3.0), (40, 4, 8.0), (64, 3, 20.0)])
def test_verify_dtw_matches_bruteforce(oracle, small_series, m, rho, eps):
    s = small_series
    off = 3100
    q = s[off - 1:off - 1 + m].copy()
    iv = [(1, 150), (3050, 3150), (5800, len(s))]
    got = oracle.verify_dtw(s, q, eps, rho, iv, 0)
    exp = refimpl.verify_dtw(s.tolist(), q.tolist(), eps, rho, iv, 0)
    assert as_pairs(got) == exp
    assert (off, 0.0) in exp
    assert got.n_dtw <= got.n_keogh_pass <= got.n_kim_pass <= got.n_verified


@pytest.mark.parametrize("m,rho,eps,alpha,beta", [(24, 2, 1.5, 1.5, 5.0), (48, 4, 3.0, 2.0, 2.0)])
def test_verify_cnsm_dtw_matches_bruteforce(oracle, small_series, m, rho, eps, alpha, beta):
    s = small_series
    off = 700
    q = s[off - 1:off - 1 + m].copy()
    iv = [(1, 200), (650, 760), (4000, 4100)]
    got = oracle.verify_cnsm_dtw(s, q, eps, rho, alpha, beta, iv, 0)
    exp = refimpl.verify_cnsm_dtw(s.tolist(), q.tolist(), eps, rho, alpha, beta, iv, 0)
    assert as_pairs(got) == exp
    assert off in got.offsets.tolist()


def test_interval_left_of_series_is_a_reference_throw(oracle, small_series):
    q = small_series[:32].copy()
    with pytest.raises(oracle.ReferenceThrows):
        oracle.verify_ed(small_series, q, 1.0, [(1, 2)], shift=100)  # end < begin -> readTimeSeries throws


# ---- full-scan executors agree with the engines on the matching chain structure ----
def test_ucr_ed_equals_engine_on_one_unbroken_chain(oracle):
    s = datagen.generate(20000, seed=5)  # multiple of 125: no phantom samples
    m = 64
    q = s[5000:5000 + m].copy()
    a = oracle.ucr_ed(s, q, 3.0, 1.5, 5.0)
    b = oracle.verify_cnsm_ed(s, q, 3.0, 1.5, 5.0, [(1, len(s) - m + 1)], 0)
    assert as_pairs(a) == as_pairs(b) and a.count > 0


def test_ucr_dtw_equals_engine_on_epoch_chains(oracle):
    n, m, rho = 250000, 64, 3
    s = datagen.generate(n, seed=9)
    q = s[120000:120000 + m].copy()
    a = oracle.ucr_dtw(s, q, 2.0, rho, 1.5, 5.0)
    iv = datagen.chain_intervals(n, m, 100000 - m + 1)
    b = oracle.verify_cnsm_dtw(s, q, 2.0, rho, 1.5, 5.0, iv, 0)
    assert (a.offsets + 1).tolist() == b.offsets.tolist()  # UcrDtw reports 0-based offsets
    assert a.distances.tolist() == b.distances.tolist() and a.count > 0


# ---- IndexBuilder step 1 ----
@pytest.mark.parametrize("w", [25, 50, 400])
def test_window_mean_runs_matches_second_restatement(oracle, w):
    n = 230000  # > 2 epochs, multiple of 125
    s = datagen.generate(n, seed=3)
    keys, first, last = oracle.window_mean_runs(s, w)
    exp = refimpl.window_mean_runs(s.tolist(), w)
    assert len(keys) == len(exp)
    assert keys.tolist() == [e[0] for e in exp]
    assert first.tolist() == [e[1] for e in exp] and last.tolist() == [e[2] for e in exp]
    assert first[0] == 1 and last[-1] == n - w + 1
    assert np.all(first[1:] == last[:-1] + 1) and np.all(last - first <= 254)


def test_window_mean_runs_phantom_zero_samples(oracle):
    # n not a multiple of 125: the block iterator pads the last 1000-byte block with zeros
    # (K/operator/file/TimeSeriesNodeIterator.java:55-59) and nextData() only counts within-node
    # advances against n (K/IndexBuilder.java:152-156): 8 full nodes + 39 samples of the 9th are fed,
    # i.e. 1030 real samples + 9 phantom zeros, so the last window starts at 1039 - 25 + 1.
    s = datagen.generate(1030, seed=4)
    keys, first, last = oracle.window_mean_runs(s, 25)
    assert last[-1] == 1015
    padded = np.concatenate([s, np.zeros(95)])
    k2, f2, l2 = oracle.window_mean_runs(padded, 25, n=1030)
    assert keys.tolist() == k2.tolist() and last.tolist() == l2.tolist()


# ---- data file codec (big-endian doubles, K/DataGenerator.java:102-113) ----
def test_series_file_roundtrip(oracle, tmp_path):
    s = datagen.generate(3000, seed=2)
    p = str(tmp_path / "data-3000")
    oracle.write_series_be(p, s)
    raw = open(p, "rb").read()
    assert len(raw) == 8 * 3000 and struct.unpack(">d", raw[:8])[0] == s[0]
    assert oracle.read_series_be(p, 3000).tolist() == s.tolist()


# ---- BASELINE config 1 (README demo): n=1e6, offset 123456, length 8192, eps 10 ----
def test_config1_self_match(oracle):
    n, m, off = 1_000_000, 8192, 123456
    s = datagen.generate(n)
    q = s[off - 1:off - 1 + m].copy()
    res = oracle.verify_ed(s, q, 10.0, [(1, n - m + 1)], 0)
    best = int(np.argmin(res.distances))
    assert res.offsets[best] == off and res.distances[best] == 0.0  # "Best: 123456, distance: 0.0"
