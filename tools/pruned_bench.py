"""Index-pruned shape of phase 2 (what the reference's engines normally see): K short merged intervals scattered over the
series instead of a full scan.  GPU through the ABI vs the oracle on one core, with parity.  usage: pruned_bench.py [n] [K]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200
from kvmatch_b200 import datagen
from oracle import kvm_oracle
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
K = int(float(sys.argv[2])) if len(sys.argv) > 2 else 100_000
m = 1024
s = datagen.generate(n); g = kvmatch_b200.GpuSeries(0); g.load(s)
rng = np.random.default_rng(11)
lefts = np.sort(rng.choice(np.arange(1, n - m - 200, 250), size=K, replace=False))
iv = np.stack([lefts, lefts + rng.integers(0, 100, K)], axis=1).astype(np.int32)
off = int(iv[K // 2, 0]); q = s[off - 1:off - 1 + m].copy()
for name, fn_g, fn_o in [
    ("cNSM-ED", lambda: g.verify_cnsm_ed(q, 5.0, 1.5, 5.0, iv), lambda: kvm_oracle.verify_cnsm_ed(s, q, 5.0, 1.5, 5.0, iv)),
    ("RSM-ED", lambda: g.verify_ed(q, 10.0, iv), lambda: kvm_oracle.verify_ed(s, q, 10.0, iv)),
    ("cNSM-DTW rho=51", lambda: g.verify_cnsm_dtw(q, 3.0, 51, 1.5, 5.0, iv), lambda: kvm_oracle.verify_cnsm_dtw(s, q, 3.0, 51, 1.5, 5.0, iv)),
]:
    fn_g(); t = time.perf_counter(); r = fn_g(); wall = (time.perf_counter() - t) * 1e3
    t = time.perf_counter(); e = fn_o(); cpu = (time.perf_counter() - t) * 1e3
    ok = r.offsets.tolist() == e.offsets.tolist() and r.distances.tolist() == e.distances.tolist()
    print(f"{name}: K={K} intervals, {r.n_verified} candidates, {r.s_total} samples; GPU kernel {r.kernel_ms:.3f} ms, wall {wall:.3f} ms; "
          f"oracle 1 core {cpu:.1f} ms; answers {r.count}; parity {ok}", flush=True)
