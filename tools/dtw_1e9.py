"""BASELINE config 4 on one GPU: cNSM-DTW, n = 1e9, m = 2048, rho = 102 (5 %), alpha 1.5, beta 5; epoch chains.
usage: dtw_1e9.py [n] [eps,eps,...]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200
from kvmatch_b200 import datagen
from oracle import kvm_oracle
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
epss = [float(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1.0, 5.0]
m, rho = 2048, 102
s = datagen.generate(n); g = kvmatch_b200.GpuSeries(0); t = time.perf_counter(); g.load(s); print(f"load {time.perf_counter()-t:.2f}s", flush=True)
iv = np.asarray(datagen.chain_intervals(n, m, 100_000 - m + 1), dtype=np.int64).reshape(-1, 2)
rng = np.random.default_rng(20260117)
offs = rng.integers(1, n - m, 3)
for eps in epss:
    for off in offs:
        q = s[off - 1:off - 1 + m].copy()
        t = time.perf_counter(); r = g.verify_cnsm_dtw(q, eps, rho, 1.5, 5.0, iv); wall = (time.perf_counter() - t) * 1e3
        # parity on the chains that hold answers + 2 random chains
        hit = np.unique(np.searchsorted(iv[:, 0], r.offsets, side="right") - 1)[:6]
        sub = iv[np.unique(np.concatenate([hit, rng.choice(len(iv), 2)]))]
        got = g.verify_cnsm_dtw(q, eps, rho, 1.5, 5.0, sub); exp = kvm_oracle.verify_cnsm_dtw(s, q, eps, rho, 1.5, 5.0, sub)
        ok = got.offsets.tolist() == exp.offsets.tolist() and got.distances.tolist() == exp.distances.tolist()
        print(f"cNSM-DTW n={n:.0e} eps={eps:g} off {off}: kernel {r.kernel_ms:.1f} ms wall {wall:.1f} ms  {r.n_verified/r.kernel_ms/1e6:.2f} Gsubseq/s  "
              f"gate {r.n_gate_pass} DTWs {r.n_lb_pass} answers {r.count} stages {r.stage_ms[0]:.2f}/{r.stage_ms[1]:.2f}/{r.stage_ms[2]:.2f}  parity on {len(sub)} chains {ok}", flush=True)
