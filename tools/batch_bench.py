"""Query sets (kvm_verify_cnsm_ed_batch): the bench workload's 10 queries in one call vs one call per query."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("KVM_PLAN_CACHE", "0")
import numpy as np, kvmatch_b200, bench
from kvmatch_b200 import datagen
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else bench.N_CFG2
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else bench.DEFAULT_CHUNK
m = bench.M
s = datagen.generate(n); g = kvmatch_b200.GpuSeries(0); g.load(s)
iv = datagen.chain_intervals(n, m, chunk)
qs = np.stack([s[o - 1:o - 1 + m] for o in bench.query_offsets(n, m, bench.N_QUERIES)])
for Q in (1, 2, 5, 10):
    sub = qs[:Q]
    g.verify_cnsm_ed_batch(sub, bench.EPSILON, bench.ALPHA, bench.BETA, iv)
    t = time.perf_counter(); reps = 5
    for _ in range(reps): res = g.verify_cnsm_ed_batch(sub, bench.EPSILON, bench.ALPHA, bench.BETA, iv)
    wall = (time.perf_counter() - t) / reps * 1e3
    ker = sum(r.kernel_ms for r in res)
    t = time.perf_counter()
    for _ in range(reps):
        single = [g.verify_cnsm_ed(q, bench.EPSILON, bench.ALPHA, bench.BETA, iv) for q in sub]
    wall1 = (time.perf_counter() - t) / reps * 1e3
    ok = all(a.offsets.tolist() == b.offsets.tolist() and a.distances.tolist() == b.distances.tolist() for a, b in zip(res, single))
    v = res[0].n_verified * Q
    print(f"Q={Q:2d}: set call wall {wall:.3f} ms (kernels {ker:.3f}; statistics pass {res[0].stage_ms[0]*Q:.3f}) -> {v/wall*1e3:.3e} subseq/s;"
          f"  {Q} single calls wall {wall1:.3f} ms -> {v/wall1*1e3:.3e} subseq/s;  x{wall1/wall:.2f}  identical {ok}", flush=True)
