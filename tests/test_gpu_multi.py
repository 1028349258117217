"""kvm_multi: several devices behind one C-ABI handle (one process, one host thread per device, host-side merge).
On a one-GPU box the same device is opened twice (two contexts, two shards, two threads); with two or more GPUs the
shards go to different devices."""
import numpy as np
import pytest

from kvmatch_b200 import _lib, datagen

pytestmark = pytest.mark.gpu


def n_gpus():
    import torch
    return torch.cuda.device_count()


def same(a, b):
    return (a.offsets.tolist() == b.offsets.tolist() and a.distances.tolist() == b.distances.tolist()
            and a.n_verified == b.n_verified and a.cnt_candidate == b.cnt_candidate)


@pytest.mark.parametrize("devices", [[0, 0], [0, 0, 0], "all"])
def test_multi_equals_single_device_and_oracle(oracle, devices):
    import kvmatch_b200
    if devices == "all":
        if n_gpus() < 2:
            pytest.skip("needs two GPUs")
        devices = list(range(min(n_gpus(), 8)))
    n, m = 500_000, 512
    s = datagen.generate(n, seed=61)
    chunk = 4096
    mg = kvmatch_b200.MultiGpuSeries(devices)
    mg.load(s, halo=m - 1, grid=chunk)
    single = kvmatch_b200.GpuSeries(0)
    single.load(s)
    off = 333_333
    q = s[off - 1:off - 1 + m].copy()
    iv = datagen.chain_intervals(n, m, chunk)
    a = mg.verify(_lib.KVM_ENGINE_CNSM_ED, q, 4.0, iv, alpha=1.5, beta=5.0)
    b = single.verify_cnsm_ed(q, 4.0, 1.5, 5.0, iv)
    e = oracle.verify_cnsm_ed(s, q, 4.0, 1.5, 5.0, iv)
    assert same(a, b) and same(a, e) and a.n_gate_pass == b.n_gate_pass == e.n_gate_pass and off in a.offsets.tolist()
    a = mg.verify(_lib.KVM_ENGINE_ED, q, 9.0, iv)
    assert same(a, single.verify_ed(q, 9.0, iv)) and same(a, oracle.verify_ed(s, q, 9.0, iv))
    a = mg.verify(_lib.KVM_ENGINE_DTW, q, 9.0, iv, rho=25)
    assert same(a, single.verify_dtw(q, 9.0, 25, iv)) and same(a, oracle.verify_dtw(s, q, 9.0, 25, iv))
    a = mg.verify(_lib.KVM_ENGINE_CNSM_DTW, q, 3.0, iv, rho=25, alpha=1.5, beta=5.0)
    assert same(a, single.verify_cnsm_dtw(q, 3.0, 25, 1.5, 5.0, iv))
    assert same(a, oracle.verify_cnsm_dtw(s, q, 3.0, 25, 1.5, 5.0, iv))
    # pruned intervals with a shift; an interval that would straddle a shard edge needs a larger halo
    rng = np.random.default_rng(4)
    lefts = np.sort(rng.choice(np.arange(100, n - m - 400), size=300, replace=False))
    piv = [(int(l), int(l) + int(rng.integers(0, 200))) for l in lefts[::2]]
    piv = [p for i, p in enumerate(piv) if i == 0 or p[0] > piv[i - 1][1] + 1]
    big = kvmatch_b200.MultiGpuSeries(devices)
    big.load(s, halo=m - 1 + 300, grid=1)
    a = big.verify(_lib.KVM_ENGINE_CNSM_ED, q, 4.0, piv, shift=25, alpha=2.0, beta=50.0)
    assert same(a, oracle.verify_cnsm_ed(s, q, 4.0, 2.0, 50.0, piv, 25))
    with pytest.raises(kvmatch_b200.KvmError) as err:
        mg.verify(_lib.KVM_ENGINE_ED, q, 9.0, [(1, n - m + 1)])  # one interval over all shards: beyond the halo
    assert err.value.code == _lib.KVM_E_RANGE
    for h in (mg, big, single):
        h.close()
