"""Run a few cNSM-ED full-scan queries (for ncu captures).  usage: one_query.py n m chunk eps reps"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200
from kvmatch_b200 import datagen
n = int(float(sys.argv[1])); m = int(sys.argv[2]); chunk = int(sys.argv[3]); eps = float(sys.argv[4]); reps = int(sys.argv[5])
s = datagen.generate(n); g = kvmatch_b200.GpuSeries(0); g.load(s)
iv = datagen.chain_intervals(n, m, chunk); qo = int(sys.argv[6]) if len(sys.argv) > 6 else min(3941173, n // 3); q = s[qo:qo + m].copy()
for i in range(reps):
    r = g.verify_cnsm_ed(q, eps, 1.5, 5.0, iv)
print(r.kernel_ms, r.count, r.n_gate_pass)
