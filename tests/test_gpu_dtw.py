"""GPU tests of the DTW lower-bound cascade and band kernel beyond tests/test_gpu_parity.py: the device envelope (a5),
LB_Keogh on the data envelope (a8), the wide-band instantiations (rho up to 409), long queries."""
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


def same(a, b):
    return a.offsets.tolist() == b.offsets.tolist() and a.distances.tolist() == b.distances.tolist()


@pytest.mark.parametrize("r,first,length", [(0, 1, 5000), (1, 1, 5000), (25, 1, 300_000), (102, 12_345, 70_001),
                                            (409, 200_000, 100_000), (512, 299_000, 1_000), (30, 5, 17)])
def test_device_envelope_equals_lemire(gpu, oracle, r, first, length):
    """a5: kvm_envelope == DtwUtils.lowerUpperLemire on the same buffer, bit for bit (clamped at the buffer ends)."""
    s = datagen.generate(300_000, seed=9)
    gpu.load(s)
    lo, up = gpu.envelope(r, first, length)
    buf = s[first - 1:first - 1 + length]
    if length > r:  # the reference throws on buffers shorter than r + 1 (documented deviation)
        el, eu = oracle.lower_upper_lemire(buf, r)
        assert lo.view(np.int64).tolist() == np.asarray(el).view(np.int64).tolist()
        assert up.view(np.int64).tolist() == np.asarray(eu).view(np.int64).tolist()
    # independent definition
    idx = np.arange(length)
    a, b = np.maximum(idx - r, 0), np.minimum(idx + r, length - 1)
    step = max(1, length // 500)
    for i in range(0, length, step):
        assert lo[i] == buf[a[i]:b[i] + 1].min() and up[i] == buf[a[i]:b[i] + 1].max()


@pytest.mark.parametrize("m,rho,eps", [(512, 25, 50.0), (512, 25, 100.0), (256, 12, 30.0)])
def test_rsm_dtw_cascade_prunes_like_the_reference(gpu, oracle, m, rho, eps):
    """a8: with LB_Keogh on the data envelope the GPU runs about as many DTWs as the reference's cascade (cfg 3
    shape: raw series, 5 % band); answers bit-exact."""
    n = 400_000
    s = datagen.generate(n, seed=33)
    gpu.load(s)
    off = 250_001
    rng = np.random.default_rng(m)
    q = s[off - 1:off - 1 + m] + rng.normal(scale=0.05, size=m)
    iv = datagen.chain_intervals(n, m, 50_000)
    got = gpu.verify_dtw(q, eps, rho, iv)
    exp = oracle.verify_dtw(s, q, eps, rho, iv)
    assert same(got, exp) and off in got.offsets.tolist()
    assert got.n_lb_pass >= got.count
    assert got.n_lb_pass <= 1.5 * exp.n_dtw + 64, (got.n_lb_pass, exp.n_dtw)


@pytest.mark.parametrize("m,rho,eps", [(2048, 102, 1.0), (2048, 102, 5.0), (1024, 51, 3.0)])
def test_cnsm_dtw_cascade_prunes_like_the_reference(gpu, oracle, m, rho, eps):
    """a8 on the cfg 4 shape (z-normalised, m = 2048, 5 % band)."""
    n = 300_000
    s = datagen.generate(n, seed=34)
    gpu.load(s)
    off = 111_111
    q = 0.9 * s[off - 1:off - 1 + m] + 3.0
    iv = datagen.chain_intervals(n, m, 20_000)
    got = gpu.verify_cnsm_dtw(q, eps, rho, 1.5, 5.0, iv)
    exp = oracle.verify_cnsm_dtw(s, q, eps, rho, 1.5, 5.0, iv)
    assert same(got, exp) and got.n_gate_pass == exp.n_gate_pass and off in got.offsets.tolist()
    assert got.n_lb_pass <= 1.5 * exp.n_dtw + 64, (got.n_lb_pass, exp.n_dtw)


@pytest.mark.parametrize("m,rho", [(4096, 204), (8192, 409), (3000, 300), (2500, 511), (1200, 160)])
def test_wide_bands(gpu, oracle, m, rho):
    """dtw_band_kernel<R> for R = 6 .. 16 (rho up to 511), raw and z-normalised, against the oracle."""
    n = 60_000
    s = datagen.generate(n, seed=35 + rho)
    gpu.load(s)
    off = 20_001
    rng = np.random.default_rng(rho)
    q = s[off - 1:off - 1 + m] + rng.normal(scale=0.02, size=m)
    iv = [(off - 40, off + 40), (40_000, 40_050)]
    got = gpu.verify_dtw(q, 5.0, rho, iv)
    exp = oracle.verify_dtw(s, q, 5.0, rho, iv)
    assert same(got, exp) and off in got.offsets.tolist()
    got = gpu.verify_cnsm_dtw(q, 2.0, rho, 1.5, 5.0, iv)
    exp = oracle.verify_cnsm_dtw(s, q, 2.0, rho, 1.5, 5.0, iv)
    assert same(got, exp) and off in got.offsets.tolist()


@pytest.mark.parametrize("m", [4096, 8192])
def test_cnsm_ed_long_queries(gpu, oracle, m):
    n = 500_000
    s = datagen.generate(n, seed=36)
    gpu.load(s)
    off = 300_003
    q = s[off - 1:off - 1 + m].copy()
    iv = datagen.chain_intervals(n, m, 30_000)
    got = gpu.verify_cnsm_ed(q, 8.0, 1.5, 5.0, iv)
    exp = oracle.verify_cnsm_ed(s, q, 8.0, 1.5, 5.0, iv)
    assert same(got, exp) and got.n_gate_pass == exp.n_gate_pass and off in got.offsets.tolist()
