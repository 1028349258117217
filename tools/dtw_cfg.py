"""Developer timing of the DTW engines on BASELINE configs 3 / 4 shapes.  usage: dtw_cfg.py n"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200
from kvmatch_b200 import datagen
import bench
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 3
s = datagen.generate(n); g = kvmatch_b200.GpuSeries(0); g.load(s)
offs = bench.query_offsets(n, 2048, 10)[:nq]
iv = datagen.chain_intervals(n, 512, 50_000)
for eps in (50.0, 75.0, 100.0):
    for off in offs[:2]:
        q = s[off - 1:off - 1 + 512].copy()
        r = g.verify_dtw(q, eps, 25, iv); r = g.verify_dtw(q, eps, 25, iv)
        print(f"RSM-DTW m=512 rho=25 eps={eps} off {off}: kernel {r.kernel_ms:.2f} ms stages {[round(x,2) for x in r.stage_ms]} dtws {r.n_lb_pass} answers {r.count}", flush=True)
iv = datagen.chain_intervals(n, 2048, 2048)
for eps in (1.0, 5.0, 10.0):
    for off in offs:
        q = s[off - 1:off - 1 + 2048].copy()
        r = g.verify_cnsm_dtw(q, eps, 102, 1.5, 5.0, iv)
        print(f"cNSM-DTW m=2048 rho=102 eps={eps} off {off}: kernel {r.kernel_ms:.2f} ms stages {[round(x,2) for x in r.stage_ms]} gate {r.n_gate_pass} "
              f"rewalked {r.n_rewalked} dtws {r.n_lb_pass} answers {r.count}", flush=True)
