"""p50 / p95 query latency through the public API at a given n (default 1e9), cNSM-ED bench workload.
usage: latency_n.py [n] [chunk] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("KVM_PLAN_CACHE", "0")
import numpy as np, kvmatch_b200, bench
from kvmatch_b200 import datagen
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else bench.DEFAULT_CHUNK
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
m = bench.M
s = datagen.generate(n); g = kvmatch_b200.GpuSeries(0); g.load(s)
iv = datagen.chain_intervals(n, m, chunk)
offs = bench.query_offsets(n, m, bench.N_QUERIES)
qs = [s[o - 1:o - 1 + m].copy() for o in offs]
for q in qs[:2]: g.verify_cnsm_ed(q, bench.EPSILON, bench.ALPHA, bench.BETA, iv)
lat, ker = [], []
for _ in range(reps):
    for q in qs:
        t = time.perf_counter(); r = g.verify_cnsm_ed(q, bench.EPSILON, bench.ALPHA, bench.BETA, iv); lat.append(time.perf_counter() - t)
        ker.append(r.kernel_ms)
lat = np.array(lat) * 1e3
print(f"n={n:.0e} chunk={chunk} chains={len(iv)} queries={len(qs)}x{reps}: latency ms p50 {np.median(lat):.3f} p95 {np.percentile(lat,95):.3f} "
      f"min {lat.min():.3f} max {lat.max():.3f}; kernel ms p50 {np.median(ker):.3f}; verified/query {r.n_verified}; "
      f"throughput at p50 {r.n_verified/np.median(lat)*1e3:.3e} subseq/s")
