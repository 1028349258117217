import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, kvmatch_b200, bench
from kvmatch_b200 import datagen
n = 100_000_000
s = datagen.generate_range(n, 0, n, bench.SEED); g = kvmatch_b200.GpuSeries(0); g.load(s)
for m in (1024, 2048, 4096, 8192):
    for off in bench.query_offsets(n, m, 3):
        q = s[off - 1:off - 1 + m].copy()
        iv = datagen.chain_intervals(n, m, 2048)
        eps = 5.0 * (m / 1024) ** 0.5
        for _ in range(2): r = g.verify_cnsm_ed(q, eps, 1.5, 5.0, iv)
        print(f"m {m} off {off} eps {eps:.2f}: kernel {r.kernel_ms:.3f} stages {[round(x,3) for x in r.stage_ms]} gate {r.n_gate_pass} rewalked {r.n_rewalked}", flush=True)
