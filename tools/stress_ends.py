"""Stress for the tile ring at chain ends: thousands of short, ragged chains (almost every turn is a masked, "non-steady"
one), single-query and query-set paths, every run compared bit for bit with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200
from kvmatch_b200 import datagen
from oracle import kvm_oracle
n = 2_000_000
s = datagen.generate(n, seed=99)
g = kvmatch_b200.GpuSeries(0); g.load(s)
rng = np.random.default_rng(1)
bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 12):
    m = int(rng.choice([25, 64, 127, 128, 200, 333]))
    K = int(rng.integers(500, 4000))
    lefts = np.sort(rng.choice(np.arange(1, n - m - 400, 401), size=K, replace=False))
    iv = np.stack([lefts, lefts + rng.integers(0, 300, K)], axis=1).astype(np.int32)
    offs = rng.integers(0, n - m, 4)
    qs = np.stack([s[o:o + m] for o in offs])
    exp = [kvm_oracle.verify_cnsm_ed(s, q, 6.0, 2.0, 50.0, iv) for q in qs]
    for rep in range(3):
        one = [g.verify_cnsm_ed(q, 6.0, 2.0, 50.0, iv) for q in qs]
        many = g.verify_cnsm_ed_batch(qs, 6.0, 2.0, 50.0, iv)
        for e, a, b in zip(exp, one, many):
            ok = (a.offsets.tolist() == e.offsets.tolist() and a.distances.tolist() == e.distances.tolist() and
                  b.offsets.tolist() == e.offsets.tolist() and b.distances.tolist() == e.distances.tolist() and
                  a.n_gate_pass == e.n_gate_pass == b.n_gate_pass)
            bad += 0 if ok else 1
    print(f"iter {it}: m={m} K={K} answers {[e.count for e in exp]} gate {[e.n_gate_pass for e in exp]} mismatches so far {bad}", flush=True)
print("STRESS", "FAILED" if bad else "ok")
