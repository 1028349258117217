"""GPU tests of the streaming cNSM statistics pass (stream_kernels.cuh): guard bands, exact re-walk, fallbacks.

The stream decides most windows from its own window sums and re-walks the reference's chain exactly only where a
decision could flip or a reported value is needed.  These tests attack the guard bands: data and queries chosen so
that many windows sit on a gate threshold, amplitudes span orders of magnitude, |mean|/std is huge, or the variance
is tiny — offsets, distances and the gate count must still equal the oracle's, and the three execution modes
(stream, stream with every banded window flagged, relay walker) must agree bit for bit.
"""
import numpy as np
import pytest

from kvmatch_b200 import _lib, datagen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import kvmatch_b200
    g = kvmatch_b200.GpuSeries(0)
    yield g
    g.set_option(_lib.KVM_OPT_CNSM_PATH, _lib.KVM_CNSM_STREAM)
    g.set_option(_lib.KVM_OPT_STREAM_FLAG_ALL, 0)
    g.close()


def same(a, b):
    return (a.offsets.tolist() == b.offsets.tolist() and a.distances.tolist() == b.distances.tolist()
            and a.n_gate_pass == b.n_gate_pass and a.n_verified == b.n_verified)


def three_modes(gpu, call):
    """Run `call()` in stream mode, flag-all mode and relay mode; return the three results."""
    gpu.set_option(_lib.KVM_OPT_CNSM_PATH, _lib.KVM_CNSM_STREAM)
    gpu.set_option(_lib.KVM_OPT_STREAM_FLAG_ALL, 0)
    a = call()
    gpu.set_option(_lib.KVM_OPT_STREAM_FLAG_ALL, 1)
    b = call()
    gpu.set_option(_lib.KVM_OPT_STREAM_FLAG_ALL, 0)
    gpu.set_option(_lib.KVM_OPT_CNSM_PATH, _lib.KVM_CNSM_RELAY)
    c = call()
    gpu.set_option(_lib.KVM_OPT_CNSM_PATH, _lib.KVM_CNSM_STREAM)
    return a, b, c


@pytest.mark.parametrize("m,chunk,eps", [(64, 500, 2.0), (256, 4096, 4.0), (1024, 2048, 5.0), (1024, 98977, 8.0),
                                         (2048, 20000, 10.0), (4096, 30000, 20.0)])
def test_modes_agree_and_match_oracle(gpu, oracle, m, chunk, eps):
    n = 600_000
    s = datagen.generate(n, seed=100 + m)
    gpu.load(s)
    off = 345_678
    q = s[off - 1:off - 1 + m].copy()
    iv = datagen.chain_intervals(n, m, chunk)
    a, b, c = three_modes(gpu, lambda: gpu.verify_cnsm_ed(q, eps, 1.5, 5.0, iv))
    exp = oracle.verify_cnsm_ed(s, q, eps, 1.5, 5.0, iv)
    assert same(a, exp) and same(b, exp) and same(c, exp)
    assert off in a.offsets.tolist()
    assert a.n_rewalked <= b.n_rewalked and a.n_rewalked >= a.count
    assert b.n_rewalked >= exp.n_gate_pass  # flag-all: every gate pass went through the exact stages


def test_threshold_hugging_windows(gpu, oracle):
    """alpha = 1 + tiny and beta = tiny: the band is thinner than anything but an exact copy, and scaled copies of the
    query sit exactly on / next to the thresholds."""
    rng = np.random.default_rng(5)
    m = 200
    base = rng.normal(size=m)
    parts = []
    for k in range(400):
        scale = 1.0 + (k - 200) * 1e-9          # std ratio within 2e-7 of 1
        shift = (k % 7 - 3) * 1e-9               # mean offset within 3e-9
        parts.append(base * scale + shift)
        parts.append(rng.normal(size=37))
    s = np.concatenate(parts)
    gpu.load(s)
    iv = datagen.chain_intervals(len(s), m, 1500)
    for alpha, beta in [(1.0 + 5e-8, 2e-9), (1.0 + 1e-9, 1e-9), (1.0, 0.0), (1.0000001, 1e-12)]:
        a, b, c = three_modes(gpu, lambda: gpu.verify_cnsm_ed(base, 0.5, alpha, beta, iv))
        exp = oracle.verify_cnsm_ed(s, base, 0.5, alpha, beta, iv)
        assert same(a, exp) and same(b, exp) and same(c, exp), (alpha, beta)
    assert exp.n_gate_pass >= 0


@pytest.mark.parametrize("offset,scale", [(1e6, 1.0), (-3e7, 0.01), (12345.678, 1e-3), (0.0, 1e4)])
def test_large_offset_and_tiny_variance(gpu, oracle, offset, scale):
    """|mean| / std up to ~1e10: cancellation in ex2/m - mean^2 (the reference's own arithmetic is what counts)."""
    n, m = 200_000, 512
    s = datagen.generate(n, seed=31) * scale + offset
    gpu.load(s)
    off = 77_001
    q = s[off - 1:off - 1 + m].copy()
    iv = datagen.chain_intervals(n, m, 3000)
    a, b, c = three_modes(gpu, lambda: gpu.verify_cnsm_ed(q, 3.0, 1.5, 5.0 * scale, iv))
    exp = oracle.verify_cnsm_ed(s, q, 3.0, 1.5, 5.0 * scale, iv)
    assert same(a, exp) and same(b, exp) and same(c, exp)


def test_amplitude_spans_orders_of_magnitude(gpu, oracle):
    """Quiet stretches (1e-3) next to bursts (1e5): the guard is evaluated per tile from the block-maximum table."""
    rng = np.random.default_rng(11)
    n, m = 400_000, 256
    s = datagen.generate(n, seed=41)
    amp = np.repeat(10.0 ** rng.integers(-3, 6, size=n // 5000), 5000)
    s = s * amp
    gpu.load(s)
    iv = datagen.chain_intervals(n, m, 2048)
    for off in (12_345, 203_000, 377_777):
        q = s[off - 1:off - 1 + m].copy()
        beta = 5.0 * float(amp[off])
        a, b, c = three_modes(gpu, lambda: gpu.verify_cnsm_ed(q, 4.0, 1.5, beta, iv))
        exp = oracle.verify_cnsm_ed(s, q, 4.0, 1.5, beta, iv)
        assert same(a, exp) and same(b, exp) and same(c, exp), off
        assert off in a.offsets.tolist()


def test_dense_answers_everywhere(gpu, oracle):
    """Periodic data: every window is an answer (queue overflow, flagged chains = all chains, list regrowth)."""
    period = np.sin(np.arange(50) * 2 * np.pi / 50)
    s = np.tile(period, 8000)
    m = 300
    gpu.load(s)
    q = s[:m].copy()
    iv = datagen.chain_intervals(len(s), m, 1000)
    a, b, c = three_modes(gpu, lambda: gpu.verify_cnsm_ed(q, 30.0, 1.5, 5.0, iv))
    exp = oracle.verify_cnsm_ed(s, q, 30.0, 1.5, 5.0, iv)
    assert same(a, exp) and same(b, exp) and same(c, exp)
    assert a.count > 100_000 and a.n_chains_rewalked == len(iv)


def test_constant_and_zero_stretches(gpu, oracle):
    """std == 0, negative variance from rounding (NaN std), all-zero tiles."""
    s = np.concatenate([np.zeros(5000), np.full(5000, 7.25), datagen.generate(20_000, seed=3), np.zeros(3000),
                        np.full(4000, -1e-3), datagen.generate(10_000, seed=4)])
    gpu.load(s)
    m = 128
    iv = datagen.chain_intervals(len(s), m, 700)
    for off in (12_000, 38_500):
        q = s[off - 1:off - 1 + m].copy()
        a, b, c = three_modes(gpu, lambda: gpu.verify_cnsm_ed(q, 6.0, 3.0, 50.0, iv))
        exp = oracle.verify_cnsm_ed(s, q, 6.0, 3.0, 50.0, iv)
        assert same(a, exp) and same(b, exp) and same(c, exp), off


def test_single_long_interval_and_ragged_lists(gpu, oracle):
    """One 1e6-candidate interval (the re-walk of a long chain), then a ragged list mixing adjacent chains, gaps,
    one-candidate intervals and an interval clamped at n."""
    n, m = 1_100_000, 512
    s = datagen.generate(n, seed=55)
    gpu.load(s)
    off = 987_654
    q = s[off - 1:off - 1 + m].copy()
    iv = [(50_000, 1_050_000)]
    a = gpu.verify_cnsm_ed(q, 5.0, 1.5, 5.0, iv)
    exp = oracle.verify_cnsm_ed(s, q, 5.0, 1.5, 5.0, iv)
    assert same(a, exp) and off in a.offsets.tolist()
    rng = np.random.default_rng(9)
    iv2, pos = [], 1
    while pos < n - m:
        length = int(rng.choice([1, 2, 30, 333, 5000, 40_000]))
        iv2.append((pos, min(pos + length - 1, n)))
        pos += length + int(rng.choice([0, 0, 1, 17, 3000]))   # 0: adjacent chains (one stream segment)
    iv2.append((n - 100, n + 500))  # clamped at n: fewer than m samples -> no window
    a, b, c = three_modes(gpu, lambda: gpu.verify_cnsm_ed(q, 5.0, 1.5, 5.0, iv2))
    exp = oracle.verify_cnsm_ed(s, q, 5.0, 1.5, 5.0, iv2)
    assert same(a, exp) and same(b, exp) and same(c, exp)


@pytest.mark.parametrize("m,rho,eps,chunk", [(128, 6, 3.0, 2048), (1024, 51, 8.0, 6000), (2048, 102, 12.0, 30000)])
def test_cnsm_dtw_modes_agree(gpu, oracle, m, rho, eps, chunk):
    n = 300_000
    s = datagen.generate(n, seed=23)
    gpu.load(s)
    off = 201_001
    q = 1.1 * s[off - 1:off - 1 + m] + 0.5
    iv = datagen.chain_intervals(n, m, chunk)
    a, b, c = three_modes(gpu, lambda: gpu.verify_cnsm_dtw(q, eps, rho, 1.5, 5.0, iv))
    exp = oracle.verify_cnsm_dtw(s, q, eps, rho, 1.5, 5.0, iv)
    assert same(a, exp) and same(b, exp) and same(c, exp)
    assert off in a.offsets.tolist()


def test_two_contexts_interleaved(oracle):
    """Two ctxs on one device, calls interleaved: per-ctx flag state (bitmap, chain table) must not leak."""
    import kvmatch_b200
    g1, g2 = kvmatch_b200.GpuSeries(0), kvmatch_b200.GpuSeries(0)
    s1, s2 = datagen.generate(120_000, seed=1), datagen.generate(90_000, seed=2)
    g1.load(s1)
    g2.load(s2)
    m = 200
    iv1, iv2 = datagen.chain_intervals(len(s1), m, 1000), datagen.chain_intervals(len(s2), m, 7000)
    for k in range(3):
        q1, q2 = s1[1000 * k + 5:1000 * k + 5 + m].copy(), s2[2000 * k + 9:2000 * k + 9 + m].copy()
        r1 = g1.verify_cnsm_ed(q1, 4.0, 1.5, 5.0, iv1)
        r2 = g2.verify_cnsm_ed(q2, 4.0, 1.5, 5.0, iv2)
        assert same(r1, oracle.verify_cnsm_ed(s1, q1, 4.0, 1.5, 5.0, iv1))
        assert same(r2, oracle.verify_cnsm_ed(s2, q2, 4.0, 1.5, 5.0, iv2))
    g1.close()
    g2.close()


def test_deferred_regular_grid_check_catches_an_irregular_list(oracle):
    """A long interval list that looks like a regular chain grid from its first and last entries is planned as one
    while the full check runs behind the kernel launches (plan cache off, K >= 4096).  A list with one hole in the
    middle must be caught by that check and re-planned: same answers as the oracle, and as the truly regular list
    except inside the hole."""
    import os
    import kvmatch_b200
    from kvmatch_b200 import _lib
    n, m, chunk = 3_000_000, 128, 512
    s = datagen.generate(n, seed=4711)
    g = kvmatch_b200.GpuSeries(0)
    g.load(s)
    g.set_option(_lib.KVM_OPT_PLAN_CACHE, 0)
    off = 1_500_123
    q = s[off - 1:off - 1 + m].copy()
    iv = np.asarray(datagen.chain_intervals(n, m, chunk), dtype=np.int32).reshape(-1, 2)
    assert len(iv) >= 4096
    full = g.verify_cnsm_ed(q, 3.0, 1.5, 5.0, iv)
    e = oracle.verify_cnsm_ed(s, q, 3.0, 1.5, 5.0, iv)
    assert full.offsets.tolist() == e.offsets.tolist() and full.distances.tolist() == e.distances.tolist()
    assert off in full.offsets.tolist()
    holed = iv.copy()
    k = int(np.searchsorted(iv[:, 0], off, side="right") - 1)
    holed[k, 1] -= 7                                   # interval k loses its last 7 starts: both ends still look regular
    got = g.verify_cnsm_ed(q, 3.0, 1.5, 5.0, holed)
    exp = oracle.verify_cnsm_ed(s, q, 3.0, 1.5, 5.0, holed)
    assert got.offsets.tolist() == exp.offsets.tolist() and got.distances.tolist() == exp.distances.tolist()
    assert got.n_verified == full.n_verified - 7 and got.n_gate_pass == exp.n_gate_pass
    again = g.verify_cnsm_ed(q, 3.0, 1.5, 5.0, iv)    # and the state left behind by the discarded attempt does no harm
    assert again.offsets.tolist() == full.offsets.tolist() and again.n_gate_pass == full.n_gate_pass
    g.close()
