import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200
from kvmatch_b200 import datagen
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
s = datagen.generate(n); g = kvmatch_b200.GpuSeries(0); g.load(s)
m, rho = 2048, 102; q = s[70_000_000 % n:70_000_000 % n + m].copy(); iv = datagen.chain_intervals(n, m, 12288)
for eps in (1.0, 5.0):
    r = g.verify_cnsm_dtw(q, eps, rho, 1.5, 5.0, iv); r = g.verify_cnsm_dtw(q, eps, rho, 1.5, 5.0, iv)
    print(f"cNSM-DTW n={n:.0e} m=2048 eps={eps}: kernel {r.kernel_ms:.1f} ms stages {r.stage_ms[0]:.2f}/{r.stage_ms[1]:.2f}/{r.stage_ms[2]:.2f} dtws {r.n_lb_pass} answers {r.count}")
m, rho = 512, 25; q = s[60_000_000 % n:60_000_000 % n + m] + np.random.default_rng(7).normal(scale=0.05, size=m); iv = datagen.chain_intervals(n, m, 100000 - m + 1)
for eps in (75.0,):
    r = g.verify_dtw(q, eps, rho, iv); r = g.verify_dtw(q, eps, rho, iv)
    print(f"RSM-DTW n={n:.0e} m=512 eps={eps}: kernel {r.kernel_ms:.2f} ms stages {r.stage_ms[0]:.2f}/{r.stage_ms[2]:.2f} dtws {r.n_lb_pass} answers {r.count}")
