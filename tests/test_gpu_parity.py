"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar: answer offsets and #answers bit-exact; distances bit-exact as well (the exact stages use the
reference's operation order with unfused binary64 ops), which is stricter than the 1e-9 relative
north_star asks for.  /root/reference is never read here.
"""
import numpy as np
import pytest

from kvmatch_b200 import datagen

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import kvmatch_b200
    g = kvmatch_b200.GpuSeries(0)
    yield g
    g.close()


@pytest.fixture(scope="module")
def series_1m():
    return datagen.generate(1_000_000)


def assert_same(got, exp, exact_dist=True):
    assert got.offsets.tolist() == exp.offsets.tolist()
    if exact_dist:
        assert got.distances.tolist() == exp.distances.tolist()
    else:
        np.testing.assert_allclose(got.distances, exp.distances, rtol=1e-9, atol=0)
    assert got.cnt_candidate == exp.cnt_candidate
    assert got.n_verified == exp.n_verified
    assert got.s_total == exp.s_total


def pruned_intervals(n, m, rng, k=200, span=60, around=()):
    lefts = np.sort(rng.choice(np.arange(1, n - m - span), size=k, replace=False)).tolist()
    lefts = sorted(set(lefts) | set(around))
    out, end = [], 0
    for l in lefts:
        l = max(int(l), end + 2)
        r = l + int(rng.integers(0, span))
        out.append((l, r))
        end = r
    return out


# ---------------------------------------------------------------- RSM-ED (a1)
def test_ed_config1_full_scan(gpu, oracle, series_1m):
    """BASELINE config 1: n=1e6, offset 123456, length 8192, eps 10 -> Best: 123456, distance 0.0."""
    s = series_1m
    n, m, off = len(s), 8192, 123456
    q = s[off - 1:off - 1 + m].copy()
    gpu.load(s)
    iv = [(1, n - m + 1)]
    got = gpu.verify_ed(q, 10.0, iv)
    exp = oracle.verify_ed(s, q, 10.0, iv)
    assert_same(got, exp)
    assert (off, 0.0) in list(zip(got.offsets.tolist(), got.distances.tolist()))


@pytest.mark.parametrize("m,eps,shift", [(25, 2.0, 0), (128, 30.0, 25), (1000, 400.0, 75), (8192, 3000.0, 0)])
def test_ed_pruned_intervals(gpu, oracle, series_1m, m, eps, shift):
    s = series_1m
    rng = np.random.default_rng(m)
    gpu.load(s)
    off = 400_000
    q = s[off - 1:off - 1 + m] + rng.normal(scale=0.01, size=m)
    iv = pruned_intervals(len(s), m, rng, around=(off + shift - 5,))
    assert_same(gpu.verify_ed(q, eps, iv, shift), oracle.verify_ed(s, q, eps, iv, shift))


def test_ed_many_answers_and_clamping(gpu, oracle):
    s = np.tile(np.array([0.0, 1.0, 0.5, -1.0]), 5000)  # periodic: thousands of exact matches
    gpu.load(s)
    q = s[:64].copy()
    iv = [(1, 300), (9000, len(s))]  # right end far beyond n-m+1: clamped like the reference
    got = gpu.verify_ed(q, 0.0, iv, 50)
    exp = oracle.verify_ed(s, q, 0.0, iv, 50)
    assert_same(got, exp)
    assert got.count > 1000


# ---------------------------------------------------------------- cNSM-ED (a2)
@pytest.mark.parametrize("m,eps,alpha,beta,chunk", [(128, 4.0, 1.5, 5.0, 4096), (1024, 10.0, 1.5, 5.0, 100000 - 1023),
                                                    (1024, 5.0, 2.0, 1.0, 20000), (256, 1.0, 1.1, 10.0, 999)])
def test_cnsm_ed_full_scan(gpu, oracle, series_1m, m, eps, alpha, beta, chunk):
    s = series_1m
    gpu.load(s)
    off = 654_321
    q = s[off - 1:off - 1 + m].copy()
    iv = datagen.chain_intervals(len(s), m, chunk)
    got = gpu.verify_cnsm_ed(q, eps, alpha, beta, iv)
    exp = oracle.verify_cnsm_ed(s, q, eps, alpha, beta, iv)
    assert_same(got, exp)
    assert got.n_gate_pass == exp.n_gate_pass
    assert off in got.offsets.tolist()


def test_cnsm_ed_pruned_intervals_with_shift(gpu, oracle, series_1m):
    s = series_1m
    m = 512
    rng = np.random.default_rng(3)
    gpu.load(s)
    off = 222_222
    q = 1.3 * s[off - 1:off - 1 + m] + 2.0  # scaled + shifted copy: z-normalised distance ~ 0
    iv = pruned_intervals(len(s), m, rng, k=500, span=300, around=(off + 50 - 7,))
    got = gpu.verify_cnsm_ed(q, 3.0, 1.5, 100.0, iv, 50)
    exp = oracle.verify_cnsm_ed(s, q, 3.0, 1.5, 100.0, iv, 50)
    assert_same(got, exp)
    assert got.n_gate_pass == exp.n_gate_pass and off in got.offsets.tolist()


@pytest.mark.parametrize("m", [25, 127, 1023, 1025])
def test_cnsm_ed_odd_lengths_and_ragged_chains(gpu, oracle, series_1m, m):
    """Odd m (outgoing rows aligned with the incoming ones), chains of very different lengths in one warp,
    chains shorter than m (no window), chains starting at odd/even offsets and ending at n."""
    s = series_1m
    rng = np.random.default_rng(m)
    gpu.load(s)
    off = 500_001
    q = s[off - 1:off - 1 + m].copy()
    iv = [(1, 3), (10, 10 + 5 * m), (40_000, 40_001), (100_001, 190_000), (300_000, 300_000 + m // 2),
          (499_990, 500_020), (700_003, 700_004 + 37), (len(s) - m - 5000, len(s))]
    iv += [(int(a), int(a) + int(rng.integers(0, 2000))) for a in np.arange(800_000, 900_000, 3001)]
    iv.sort()
    got = gpu.verify_cnsm_ed(q, 6.0, 2.0, 50.0, iv)
    exp = oracle.verify_cnsm_ed(s, q, 6.0, 2.0, 50.0, iv)
    assert_same(got, exp)
    assert got.n_gate_pass == exp.n_gate_pass and off in got.offsets.tolist()


def test_cnsm_ed_many_regions(gpu, oracle, series_1m):
    """More walker CTAs than can be resident at once (> 2 per SM)."""
    s = series_1m
    m = 64
    gpu.load(s)
    q = s[250_000:250_000 + m].copy()
    iv = datagen.chain_intervals(len(s), m, 60)  # 16666 chains -> 521 regions
    got = gpu.verify_cnsm_ed(q, 2.0, 1.5, 5.0, iv)
    exp = oracle.verify_cnsm_ed(s, q, 2.0, 1.5, 5.0, iv)
    assert_same(got, exp)
    assert got.n_gate_pass == exp.n_gate_pass


def test_cnsm_ed_degenerate_query_and_constant_data(gpu, oracle):
    s = np.concatenate([np.full(3000, 2.5), datagen.generate(5000, seed=8)])
    gpu.load(s)
    iv = [(1, len(s) - 63)]
    # constant query: stdQ == 0 -> every gate comparison is false in the reference
    assert gpu.verify_cnsm_ed(np.full(64, 1.0), 5.0, 1.5, 5.0, iv).count == 0
    # constant data windows: std == 0 (or NaN from negative variance) -> silently dropped
    q = s[4000:4064].copy()
    assert_same(gpu.verify_cnsm_ed(q, 5.0, 1.5, 5.0, iv), oracle.verify_cnsm_ed(s, q, 5.0, 1.5, 5.0, iv))


# ---------------------------------------------------------------- RSM-DTW (a3) / cNSM-DTW (a4)
@pytest.mark.parametrize("m,rho,eps", [(64, 3, 6.0), (128, 6, 12.0), (512, 25, 40.0), (600, 77, 80.0)])
def test_dtw_scan(gpu, oracle, m, rho, eps):
    s = datagen.generate(200_000, seed=21)
    gpu.load(s)
    off = 150_000
    rng = np.random.default_rng(rho)
    q = s[off - 1:off - 1 + m] + rng.normal(scale=0.05, size=m)
    iv = [(1, 60_000), (149_000, 151_000), (199_000, len(s) - m + 1)]
    got = gpu.verify_dtw(q, eps, rho, iv)
    exp = oracle.verify_dtw(s, q, eps, rho, iv)
    assert_same(got, exp)
    assert off in got.offsets.tolist()
    # all three bounds of the reference's cascade run on the GPU too (its data envelope, taken inside the window, is
    # tighter than the reference's buffer-wide one): about as many DTWs as the reference, never fewer than the answers
    assert got.count <= got.n_lb_pass <= 1.5 * exp.n_dtw + 64


@pytest.mark.parametrize("m,rho,eps,chunk", [(128, 6, 3.0, 5000), (512, 25, 6.0, 100000 - 511),
                                             (2048, 102, 12.0, 30000)])
def test_cnsm_dtw_scan(gpu, oracle, m, rho, eps, chunk):
    s = datagen.generate(300_000, seed=22)
    gpu.load(s)
    off = 123_456
    q = 0.8 * s[off - 1:off - 1 + m] - 1.0
    iv = datagen.chain_intervals(len(s), m, chunk)
    got = gpu.verify_cnsm_dtw(q, eps, rho, 1.5, 5.0, iv)
    exp = oracle.verify_cnsm_dtw(s, q, eps, rho, 1.5, 5.0, iv)
    assert_same(got, exp)
    assert got.n_gate_pass == exp.n_gate_pass and off in got.offsets.tolist()


def test_dtw_rho0_equals_ed(gpu, series_1m):
    s = series_1m[:100_000]
    gpu.load(s)
    q = s[5000:5128].copy()
    iv = [(1, len(s) - 127)]
    a = gpu.verify_dtw(q, 25.0, 0, iv)
    b = gpu.verify_ed(q, 25.0, iv)
    assert a.offsets.tolist() == b.offsets.tolist() and a.distances.tolist() == b.distances.tolist()


# ---------------------------------------------------------------- IndexBuilder step 1 (a10/a11)
@pytest.mark.parametrize("w", [25, 50, 100, 200, 400])
def test_window_mean_runs(gpu, oracle, series_1m, w):
    gpu.load(series_1m)
    keys, first, last, _, _ = gpu.window_mean_runs(w)
    ek, ef, el = oracle.window_mean_runs(series_1m, w)
    assert first.tolist() == ef.tolist() and last.tolist() == el.tolist()
    assert keys.view(np.int64).tolist() == ek.view(np.int64).tolist()  # Double.equals: bitwise


def test_window_mean_runs_phantom_padding_and_long_runs(gpu, oracle):
    s = np.concatenate([np.full(2000, 3.0), datagen.generate(1030, seed=4)])  # constant prefix: runs split at 255
    gpu.load(s)
    keys, first, last, _, _ = gpu.window_mean_runs(25)
    ek, ef, el = oracle.window_mean_runs(s, 25)
    assert first.tolist() == ef.tolist() and last.tolist() == el.tolist() and keys.tolist() == ek.tolist()
    assert (last - first).max() == 254


# ---------------------------------------------------------------- boundary behaviour
def test_file_loader_big_endian(gpu, oracle, tmp_path):
    s = datagen.generate(50_000, seed=6)
    p = str(tmp_path / "data-50000")
    oracle.write_series_be(p, s)
    gpu.load_file(p, len(s))
    q = s[1000:1256].copy()
    iv = [(1, len(s) - 255)]
    assert_same(gpu.verify_ed(q, 5.0, iv), oracle.verify_ed(s, q, 5.0, iv))
    # a shard with halo: samples [20001, 40000] of the same file
    gpu.load_file(p, len(s), first=20001, count=20000)
    iv2 = [(20001, 39000)]
    assert_same(gpu.verify_cnsm_ed(q, 50.0, 2.0, 50.0, iv2), oracle.verify_cnsm_ed(s, q, 50.0, 2.0, 50.0, iv2))


def test_error_codes(gpu, series_1m):
    import kvmatch_b200
    fresh = kvmatch_b200.GpuSeries(0)
    with pytest.raises(kvmatch_b200.KvmError) as e:
        fresh.verify_ed(np.zeros(32), 1.0, [(1, 2)])
    assert e.value.code == -6  # KVM_E_STATE
    fresh.load(series_1m[:1000])
    with pytest.raises(kvmatch_b200.KvmError) as e:
        fresh.verify_ed(np.zeros(32), 1.0, [(1, 2)], shift=100)  # the reference throws IllegalArgumentException
    assert e.value.code == -7
    fresh.load(series_1m[:1000], n=5000, first=2001)
    with pytest.raises(kvmatch_b200.KvmError) as e:
        fresh.verify_ed(np.zeros(32), 1.0, [(1, 200)])  # outside this shard
    assert e.value.code == -7
    fresh.close()


def test_engine_mirror_output(gpu, series_1m, caplog):
    import logging

    import kvmatch_b200
    gpu.load(series_1m)
    q = series_1m[123455:123455 + 8192].copy()
    stats = [kvmatch_b200.StatisticInfo() for _ in range(6)]
    eng = kvmatch_b200.QueryEngine(gpu)
    with caplog.at_level(logging.INFO, logger="kvmatch_b200"):
        assert eng.query(stats, q, 10.0) is True
    assert eng.answers[0] == (123456, 0.0)
    assert "Best: 123456, distance: 0.0" in caplog.text and "#answers: %d" % len(eng.answers) in caplog.text
    assert stats[4].values == [float(len(eng.answers))]


# ---------------------------------------------------------------- full-size properties (no oracle)
def test_cnsm_ed_1e7_properties(gpu):
    n, m = 10_000_000, 1024
    s = datagen.generate(n, seed=77)
    gpu.load(s)
    off = 7_654_321
    q = s[off - 1:off - 1 + m].copy()
    iv = datagen.chain_intervals(n, m, 16384)
    r5 = gpu.verify_cnsm_ed(q, 5.0, 1.5, 5.0, iv)
    r10 = gpu.verify_cnsm_ed(q, 10.0, 1.5, 5.0, iv)
    assert off in r5.offsets.tolist()
    assert np.all(np.diff(r5.offsets) > 0) and np.all(r5.distances <= 5.0)
    assert set(r5.offsets.tolist()) <= set(r10.offsets.tolist())  # monotone in epsilon
    assert r5.n_gate_pass == r10.n_gate_pass and r5.n_verified == n - m + 1
    # idempotent
    again = gpu.verify_cnsm_ed(q, 5.0, 1.5, 5.0, iv)
    assert again.offsets.tolist() == r5.offsets.tolist() and again.distances.tolist() == r5.distances.tolist()


def test_cnsm_plan_reuse_across_calls(oracle):
    """The cNSM engines keep the previous call's interval plan on the device: the same intervals under other
    queries, other intervals, calls of the other engines in between and a series reload must all match the oracle."""
    import kvmatch_b200
    g = kvmatch_b200.GpuSeries(0)
    n, m = 150_000, 256
    s = datagen.generate(n, seed=77)
    g.load(s)
    iv_a = datagen.chain_intervals(n, m, 4096)
    iv_b = datagen.chain_intervals(n, m, 3000, 2000, 140_000)
    qs = [s[off:off + m].copy() for off in (5_000, 71_234, 120_001)]
    plan = [(qs[0], iv_a), (qs[1], iv_a), (qs[2], iv_b), (qs[0], iv_b), ("ed", None), (qs[1], iv_b), (qs[2], iv_a)]
    for q, iv in plan:
        if isinstance(q, str):
            assert_same(g.verify_ed(qs[0], 3.0, iv_a), oracle.verify_ed(s, qs[0], 3.0, iv_a))
            continue
        assert_same(g.verify_cnsm_ed(q, 4.0, 1.5, 3.0, iv), oracle.verify_cnsm_ed(s, q, 4.0, 1.5, 3.0, iv))
    s2 = datagen.generate(n, seed=78)  # a new series invalidates the plan even for a byte-identical interval list
    g.load(s2)
    q = s2[9_000:9_000 + m].copy()
    assert_same(g.verify_cnsm_ed(q, 4.0, 1.5, 3.0, iv_a), oracle.verify_cnsm_ed(s2, q, 4.0, 1.5, 3.0, iv_a))
    g.close()


@pytest.mark.parametrize("m,rho,eps", [(128, 6, 3.0), (1000, 50, 9.0)])
def test_scan_ucr_dtw_matches_reference_executor(gpu, oracle, m, rho, eps):
    """f4: kvm_scan_ucr_dtw against the oracle's restatement of UcrDtwQueryExecutor (EPOCH buffers with m-1 overlap,
    statistics restart per buffer, 0-based offsets), three buffers long."""
    n = 250_000  # multiple of 125: no phantom samples in the reference's block feed
    s = datagen.generate(n, seed=4242)
    gpu.load(s)
    q = s[123_000:123_000 + m].copy() + 0.01 * np.cos(np.arange(m))
    got = gpu.scan_ucr_dtw(q, eps, rho, 1.5, 5.0)
    exp = oracle.ucr_dtw(s, q, eps, rho, 1.5, 5.0)
    assert got.offsets.tolist() == exp.offsets.tolist()
    assert got.distances.tolist() == exp.distances.tolist()
    assert got.count > 0 and got.offsets.min() >= 0
    assert got.n_verified == n - m + 1


@pytest.mark.parametrize("n", [250_007, 100_124, 199_999])
def test_scan_ucr_dtw_phantom_samples(gpu, oracle, n):
    """n % 125 != 0: the executor's block iterator zero-pads the last 1000-byte block and also verifies the windows
    that run into the padding."""
    s = datagen.generate(n, seed=n)
    s[-300:] *= 0.01  # a quiet tail: windows over the zero padding can pass the gate
    gpu.load(s)
    m = 128
    q = np.concatenate([s[-100:], np.zeros(m - 100)]) + 0.001 * np.cos(np.arange(m))
    got = gpu.scan_ucr_dtw(q, 3.0, 6, 3.0, 50.0)
    exp = oracle.ucr_dtw(s, q, 3.0, 6, 3.0, 50.0)
    assert got.offsets.tolist() == exp.offsets.tolist()
    assert got.distances.tolist() == exp.distances.tolist()
    fed = cnt = 0  # samples the block iterator feeds: 125-sample nodes, only within-node advances count against n
    for _ in range((n + 124) // 125):
        within = min(124, max(0, n - cnt))
        fed += 1 + within
        cnt += within
        if within < 124:
            break
    assert got.n_verified == fed - m + 1 and fed > n    # the windows over the zero padding were scanned


@pytest.mark.parametrize("n,m,eps", [(250_000, 128, 3.0), (1_000_000, 1024, 8.0), (300_007, 256, 4.0)])
def test_scan_ucr_ed_matches_reference_executor(gpu, oracle, n, m, eps):
    """f4: kvm_scan_ucr_ed against the oracle's restatement of UcrEdQueryExecutor — ONE statistics chain over the whole
    series (never reset), 1-based offsets, the block iterator's zero padding when n % 125 != 0."""
    s = datagen.generate(n, seed=n + m)
    gpu.load(s)
    off = (2 * n) // 3
    q = s[off:off + m].copy() + 0.01 * np.cos(np.arange(m))
    got = gpu.scan_ucr_ed(q, eps, 1.5, 5.0)
    exp = oracle.ucr_ed(s, q, eps, 1.5, 5.0)
    assert got.offsets.tolist() == exp.offsets.tolist()
    assert got.distances.tolist() == exp.distances.tolist()
    assert got.n_gate_pass == exp.n_gate_pass
    assert got.count > 0 and off + 1 in got.offsets.tolist()
    assert got.n_verified == exp.n_verified


def test_scan_ucr_dtw_needs_whole_series(gpu):
    import kvmatch_b200
    s = datagen.generate(50_000, seed=5)
    gpu.load(s[1000:], n=50_000, first=1001)
    with pytest.raises(kvmatch_b200.KvmError) as e:
        gpu.scan_ucr_dtw(s[:64].copy(), 1.0, 3, 1.5, 5.0)
    assert e.value.code == kvmatch_b200._lib.KVM_E_STATE


@pytest.mark.parametrize("w", [25, 400])
def test_build_index_file_matches_oracle_image(gpu, oracle, series_1m, tmp_path, w):
    """f2/f3: GPU window-mean pass + host step 2 + file codec against the oracle's image of files/index-N-w."""
    s = series_1m[:300_000]
    gpu.load(s)
    path = tmp_path / f"index-{len(s)}-{w}"
    info = gpu.build_index_file(w, str(path))
    exp, rows1, rows = oracle.index_file_image(s, w)
    assert path.read_bytes() == exp
    assert (info.n_rows_step1, info.n_rows, info.file_bytes) == (rows1, rows, len(exp))
    assert info.n_offsets == len(s) - w + 1


def test_full_size_1e8_against_oracle_on_chain_subsets(oracle):
    """BASELINE configs[1] at full size (n = 1e8, m = 1024): size-independent properties on the whole scan, and exact
    parity with the oracle on every chain that holds an answer plus a random sample of other chains (a chain is an
    independent unit of the path, so the oracle can be run on it alone)."""
    import kvmatch_b200
    n, m, chunk = 100_000_000, 1024, 6144
    s = datagen.generate(n)
    g = kvmatch_b200.GpuSeries(0)
    g.load(s)
    iv = np.asarray(datagen.chain_intervals(n, m, chunk), dtype=np.int64).reshape(-1, 2)
    rng = np.random.default_rng(5)
    for off in (3_941_174, 61_550_030):
        q = s[off - 1:off - 1 + m].copy()
        r5 = g.verify_cnsm_ed(q, 5.0, 1.5, 5.0, iv)
        r1 = g.verify_cnsm_ed(q, 1.0, 1.5, 5.0, iv)
        assert r5.n_verified == n - m + 1 and r5.cnt_candidate == n - m + 1
        assert np.all(np.diff(r5.offsets) > 0) and np.all(r5.distances <= 5.0)
        i = r5.offsets.tolist().index(off)
        assert r5.distances[i] < 1e-9   # the planted self-match (not exactly 0: chain sums vs the query's fresh sums)
        assert set(r1.offsets.tolist()) <= set(r5.offsets.tolist())      # monotone in epsilon
        hit = np.unique(np.searchsorted(iv[:, 0], r5.offsets, side="right") - 1)
        others = rng.choice(len(iv), size=24, replace=False)
        sub = iv[np.unique(np.concatenate([hit, others]))]
        exp = oracle.verify_cnsm_ed(s, q, 5.0, 1.5, 5.0, sub)
        got = g.verify_cnsm_ed(q, 5.0, 1.5, 5.0, sub)
        assert_same(got, exp)
        assert got.n_gate_pass == exp.n_gate_pass
        assert got.offsets.tolist() == [o for o in r5.offsets.tolist() if np.any((sub[:, 0] <= o) & (o <= sub[:, 1]))]
    # RSM-ED and cNSM-DTW on the same series, same scheme
    q = s[off - 1:off - 1 + m].copy()
    full = g.verify_ed(q, 10.0, [(1, n - m + 1)])
    assert off in full.offsets.tolist()
    sub = iv[np.unique(np.concatenate([np.searchsorted(iv[:, 0], full.offsets, side="right") - 1, rng.choice(len(iv), 16)]))]
    assert_same(g.verify_ed(q, 10.0, sub), oracle.verify_ed(s, q, 10.0, sub))
    m2, rho = 512, 25
    q2 = s[off - 1:off - 1 + m2].copy()
    iv2 = np.asarray(datagen.chain_intervals(n, m2, 100_000 - m2 + 1), dtype=np.int64).reshape(-1, 2)
    d = g.verify_cnsm_dtw(q2, 1.0, rho, 1.5, 5.0, iv2)
    assert off in d.offsets.tolist() and d.n_verified == n - m2 + 1
    sub2 = iv2[np.unique(np.concatenate([np.searchsorted(iv2[:, 0], d.offsets, side="right") - 1, rng.choice(len(iv2), 3)]))]
    assert_same(g.verify_cnsm_dtw(q2, 1.0, rho, 1.5, 5.0, sub2), oracle.verify_cnsm_dtw(s, q2, 1.0, rho, 1.5, 5.0, sub2))
    # BASELINE configs[2]: RSM-DTW, m = 512, rho = 25 (5 %), the reference's raw-data epsilon grid; the engine's data
    # envelope covers one interval read, so the scan is cut into EPOCH-sized intervals as the reference's executor does
    for eps in (50.0, 75.0):
        r = g.verify_dtw(q2, eps, rho, iv2)
        assert off in r.offsets.tolist() and r.n_verified == n - m2 + 1
        assert np.all(np.diff(r.offsets) > 0) and np.all(r.distances <= eps)
        sub3 = iv2[np.unique(np.concatenate([np.searchsorted(iv2[:, 0], r.offsets, side="right") - 1, rng.choice(len(iv2), 3)]))][:8]
        e3 = oracle.verify_dtw(s, q2, eps, rho, sub3)
        g3 = g.verify_dtw(q2, eps, rho, sub3)
        assert_same(g3, e3)
        assert g3.n_lb_pass <= 1.5 * e3.n_dtw + 64     # the cascade prunes like the reference's
    # one merged interval of 1e6 candidates (what an index-pruned phase 1 can hand over unchunked): a single
    # statistics chain, K/NormQueryEngine.java:487
    lo = max(1, off - 400_000)
    one = [(lo, lo + 1_000_000 - 1)]
    assert_same(g.verify_cnsm_ed(q, 5.0, 1.5, 5.0, one), oracle.verify_cnsm_ed(s, q, 5.0, 1.5, 5.0, one))
    g.close()


def test_full_size_1e9_against_oracle_on_chain_subsets(oracle):
    """The metric's size (n = 1e9, 8 GB resident): whole-scan properties, and exact parity with the oracle on every chain
    that holds an answer plus random other chains, for cNSM-ED (m = 1024) and cNSM-DTW (BASELINE configs[3]: m = 2048,
    rho = 102).  The oracle only ever touches the chains it is given."""
    import kvmatch_b200
    n, chunk = 1_000_000_000, 2048
    s = datagen.generate_range(n, 0, n, datagen.DEFAULT_SEED)
    g = kvmatch_b200.GpuSeries(0)
    g.load(s)
    rng = np.random.default_rng(9)
    off = 470_341_747
    for m, rho, eps in ((1024, None, 5.0), (2048, 102, 1.0)):
        iv = np.asarray(datagen.chain_intervals(n, m, chunk), dtype=np.int64).reshape(-1, 2)
        q = s[off - 1:off - 1 + m].copy()
        if rho is None:
            full = g.verify_cnsm_ed(q, eps, 1.5, 5.0, iv)
        else:
            full = g.verify_cnsm_dtw(q, eps, rho, 1.5, 5.0, iv)
        assert full.n_verified == n - m + 1
        assert off in full.offsets.tolist() and np.all(np.diff(full.offsets) > 0) and np.all(full.distances <= eps)
        hit = np.unique(np.searchsorted(iv[:, 0], full.offsets, side="right") - 1)
        sub = iv[np.unique(np.concatenate([hit[:40], rng.choice(len(iv), size=24, replace=False)]))]
        if rho is None:
            got, exp = g.verify_cnsm_ed(q, eps, 1.5, 5.0, sub), oracle.verify_cnsm_ed(s, q, eps, 1.5, 5.0, sub)
        else:
            got, exp = g.verify_cnsm_dtw(q, eps, rho, 1.5, 5.0, sub), oracle.verify_cnsm_dtw(s, q, eps, rho, 1.5, 5.0, sub)
        assert_same(got, exp)
        assert got.n_gate_pass == exp.n_gate_pass
        assert got.offsets.tolist() == [o for o in full.offsets.tolist() if np.any((sub[:, 0] <= o) & (o <= sub[:, 1]))]
    g.close()


@pytest.mark.parametrize("m,chunk", [(128, 4096), (1024, 6144), (257, 5000)])
def test_cnsm_ed_query_set_equals_single_queries(gpu, oracle, series_1m, m, chunk):
    """kvm_verify_cnsm_ed_batch: one statistics pass for a set of queries must return, per query, exactly what the
    single-query entry (itself checked against the oracle) returns — sparse, dense and degenerate queries mixed."""
    s = series_1m
    n = len(s)
    gpu.load(s)
    iv = datagen.chain_intervals(n, m, chunk)
    offs = [1234, 250_000, 500_017, 777_777, 999_000 - m]
    qs = np.stack([s[o:o + m] for o in offs] + [np.full(m, 3.25)])      # the last one is degenerate (std = 0)
    qs[1] = qs[1] + 0.01 * np.sin(np.arange(m))                          # a near match instead of an exact one
    got = gpu.verify_cnsm_ed_batch(qs, 5.0, 1.5, 5.0, iv)
    assert len(got) == len(qs)
    for q, r in zip(qs, got):
        one = gpu.verify_cnsm_ed(q, 5.0, 1.5, 5.0, iv)
        assert_same(r, one)
        assert r.n_gate_pass == one.n_gate_pass and r.n_exact == one.n_exact
    exp = oracle.verify_cnsm_ed(s, qs[0], 5.0, 1.5, 5.0, iv)
    assert_same(got[0], exp)
    assert got[-1].count == 0


def test_every_window_is_an_answer_buffers_regrow(oracle):
    """Periodic data, generous thresholds: all 3e5 windows are answers, so the candidate and answer buffers overflow
    their first allocation and every engine (and the query-set entry) must re-run and still match the oracle."""
    import kvmatch_b200
    n, m = 300_000, 64
    t = np.arange(n)
    s = np.sin(2 * np.pi * t / 64.0) * 3 + 0.001 * np.cos(t * 0.37)
    g = kvmatch_b200.GpuSeries(0)
    g.load(s)
    iv = datagen.chain_intervals(n, m, 5000)
    q = s[1000:1000 + m].copy()
    for got, exp in [
        (g.verify_cnsm_ed(q, 50.0, 2.0, 100.0, iv), oracle.verify_cnsm_ed(s, q, 50.0, 2.0, 100.0, iv)),
        (g.verify_ed(q, 500.0, iv), oracle.verify_ed(s, q, 500.0, iv)),
        (g.verify_cnsm_dtw(q, 50.0, 3, 2.0, 100.0, iv), oracle.verify_cnsm_dtw(s, q, 50.0, 3, 2.0, 100.0, iv)),
        (g.verify_dtw(q, 500.0, 3, iv), oracle.verify_dtw(s, q, 500.0, 3, iv)),
    ]:
        assert got.count == n - m + 1
        assert_same(got, exp)
    qs = np.stack([q, s[5:5 + m], s[77:77 + m]])
    for r, qq in zip(g.verify_cnsm_ed_batch(qs, 50.0, 2.0, 100.0, iv), qs):
        assert_same(r, oracle.verify_cnsm_ed(s, qq, 50.0, 2.0, 100.0, iv))
    small = g.verify_cnsm_ed(q, 0.01, 1.1, 0.1, iv)                     # and a selective call on the grown buffers
    assert_same(small, oracle.verify_cnsm_ed(s, q, 0.01, 1.1, 0.1, iv))
    g.close()


def test_rsm_engines_coalesce_adjacent_intervals(gpu, oracle):
    """The raw-series engines merge intervals whose window starts are adjacent into one run of candidates (and recognise
    a regular grid without per-interval planning): answers, #candidates, #verified and s_total must not change — mixed
    adjacent / separate / clamped intervals, with a shift, against the oracle."""
    n, m = 400_000, 256
    s = datagen.generate(n, seed=909)
    gpu.load(s)
    off = 123_457
    q = s[off - 1:off - 1 + m].copy()
    shift = 50
    iv = [(60, 5_000), (5_001, 9_000), (9_001, 9_001), (20_000, 30_000), (30_001, 30_500), (off + shift - 40, off + shift + 40),
          (off + shift + 41, off + shift + 2_000), (n - m - 3_000 + shift, n - m + 1 + shift), (n - m + 2 + shift, n + shift)]
    assert_same(gpu.verify_ed(q, 12.0, iv, shift), oracle.verify_ed(s, q, 12.0, iv, shift))
    g2, e2 = gpu.verify_dtw(q, 12.0, 12, iv, shift), oracle.verify_dtw(s, q, 12.0, 12, iv, shift)
    assert_same(g2, e2)
    assert off in g2.offsets.tolist()
    grid = datagen.chain_intervals(n, m, 1000)                       # a regular grid: one run
    assert_same(gpu.verify_ed(q, 12.0, grid), oracle.verify_ed(s, q, 12.0, grid))
    ragged = [tuple(x) for x in np.asarray(grid).reshape(-1, 2)]
    ragged[7] = (ragged[7][0], ragged[7][1] - 1)                     # one gap: not regular any more
    assert_same(gpu.verify_ed(q, 12.0, ragged), oracle.verify_ed(s, q, 12.0, ragged))


def test_rsm_ed_deferred_grid_check_catches_a_holed_list(gpu, oracle):
    """RSM-ED takes a long list that looks regular from its two ends as one run and checks it in full behind the kernel
    launch; a hole in the middle must be caught and the list planned interval by interval."""
    n, m, chunk = 3_000_000, 128, 512
    s = datagen.generate(n, seed=4712)
    gpu.load(s)
    off = 2_000_321
    q = s[off - 1:off - 1 + m].copy()
    iv = np.asarray(datagen.chain_intervals(n, m, chunk), dtype=np.int32).reshape(-1, 2)
    assert len(iv) >= 4096
    assert_same(gpu.verify_ed(q, 15.0, iv), oracle.verify_ed(s, q, 15.0, iv))
    holed = iv.copy()
    k = int(np.searchsorted(iv[:, 0], off, side="right") - 1)
    holed[k] = (off + 1, holed[k, 1])                  # the interval that holds the planted match now starts behind it
    got, exp = gpu.verify_ed(q, 15.0, holed), oracle.verify_ed(s, q, 15.0, holed)
    assert_same(got, exp)
    assert off not in got.offsets.tolist()
