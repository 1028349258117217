"""RSM-ED latency floor on a small series (BASELINE configs[0] shape: n = 1e6, query cut at 123456): kernel time against query
length and epsilon, parity against the oracle.  usage: python tools/ed_small.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200, bench
from kvmatch_b200 import datagen
from oracle import kvm_oracle as o
# BASELINE configs[0] shape: n = 1e6, m = 8192, eps = 10, query at 123456
n = 1_000_000
s = datagen.generate(n)
g = kvmatch_b200.GpuSeries(0); g.load(s)
for m, eps in ((8192, 10.0), (1024, 10.0), (1024, 40.0), (50, 3.0), (139, 3.0), (140, 3.0), (141, 2.0), (165, 30.0), (300, 60.0)):
    q = s[123455:123455 + m].copy()
    iv = [(1, n - m + 1)]
    g.verify_ed(q, eps, iv)
    r = g.verify_ed(q, eps, iv)
    e = o.verify_ed(s, q, eps, iv)
    print(f"n=1e6 m={m} eps={eps}: kernel {r.kernel_ms:.4f} ms answers {r.count} parity", r.offsets.tolist() == e.offsets.tolist() and r.distances.tolist() == e.distances.tolist(), flush=True)
