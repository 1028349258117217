"""GPU tests of the fused window-mean pass (wmean_kernels.cuh, kvm_window_mean_runs_all): every width of Sigma in one
pass over the series, bitwise equal to the oracle's restatement of IndexBuilder step 1 (K/IndexBuilder.java:194-301)."""
import numpy as np
import pytest

from kvmatch_b200 import datagen

pytestmark = pytest.mark.gpu

SIGMA = (25, 50, 100, 200, 400)


@pytest.fixture(scope="module")
def gpu():
    import kvmatch_b200
    g = kvmatch_b200.GpuSeries(0)
    yield g
    g.close()


def check(gpu, oracle, s, widths=SIGMA):
    gpu.load(s)
    res = gpu.window_mean_runs_all(widths)
    assert res.widths == list(widths)
    for w, (keys, first, last) in zip(widths, res.runs):
        ek, ef, el = oracle.window_mean_runs(s, w)
        assert first.tolist() == ef.tolist(), w
        assert last.tolist() == el.tolist(), w
        assert keys.view(np.int64).tolist() == ek.view(np.int64).tolist(), w  # Double.equals: bitwise
    return res


def test_all_widths_1m(gpu, oracle):
    res = check(gpu, oracle, datagen.generate(1_000_000))
    assert res.n_launches >= 1 and res.n_runs > 0


@pytest.mark.parametrize("n", [999_999, 1_000_124, 200_001, 100_399, 100_400, 2_000])
def test_phantom_padding_and_epoch_edges(gpu, oracle, n):
    """n % 125 != 0 (zero-padded last block), series barely longer than one epoch, shorter than the widest window."""
    check(gpu, oracle, datagen.generate(n, seed=n))


def test_single_widths_and_subsets(gpu, oracle):
    s = datagen.generate(300_000, seed=12)
    check(gpu, oracle, s, (50,))
    check(gpu, oracle, s, (400, 25))
    check(gpu, oracle, s, (33, 66, 1000))


def test_constant_and_long_runs(gpu, oracle):
    s = np.concatenate([np.full(5000, 3.0), datagen.generate(20_000, seed=4), np.zeros(4000), np.full(3000, -0.05)])
    res = check(gpu, oracle, s)
    assert max(int((l - f).max()) for _, f, l in res.runs) == 254


def test_means_on_bucket_boundaries(gpu, oracle):
    """Window means that sit exactly on / a few ulps from multiples of 0.05: every such window is ambiguous for the
    stream and must come out of the exact re-walk."""
    rng = np.random.default_rng(2)
    base = np.round(rng.normal(size=40_000) * 2) * 0.05          # multiples of 0.05: sums of w of them hit boundaries
    s = np.concatenate([base, base * 1e3 + 1e-9 * rng.normal(size=len(base)), rng.normal(size=30_000)])
    res = check(gpu, oracle, s)
    assert res.n_chains_rewalked > 0


@pytest.mark.parametrize("offset,scale", [(1e6, 1.0), (-4e7, 1e-2), (0.0, 1e5)])
def test_large_offsets(gpu, oracle, offset, scale):
    check(gpu, oracle, datagen.generate(250_000, seed=77) * scale + offset)


def test_agrees_with_single_width_entry(gpu, oracle):
    s = datagen.generate(400_000, seed=5)
    gpu.load(s)
    res = gpu.window_mean_runs_all(SIGMA)
    for w, (keys, first, last) in zip(SIGMA, res.runs):
        k1, f1, l1, _, _ = gpu.window_mean_runs(w)
        assert first.tolist() == f1.tolist() and last.tolist() == l1.tolist()
        assert keys.view(np.int64).tolist() == k1.view(np.int64).tolist()
