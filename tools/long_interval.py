"""The long-interval case: ONE merged interval (= one statistics chain, K/NormQueryEngine.java:487) of c candidates, as an
index-pruned phase 1 can hand it over unchunked, against the same windows cut into chains of 2048.  usage: long_interval.py [n]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200, bench
from kvmatch_b200 import datagen
from oracle import kvm_oracle
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
m = 1024
s = datagen.generate_range(n, 0, n, bench.SEED); g = kvmatch_b200.GpuSeries(0); g.load(s)
off = bench.query_offsets(n, m, 10)[0]
q = s[off - 1:off - 1 + m].copy()
for c in (100_000, 1_000_000, 10_000_000, n - m + 1):
    lo = max(1, min(off - c // 2, n - m + 1 - c + 1))
    one = [(lo, lo + c - 1)]
    g.verify_cnsm_ed(q, 5.0, 1.5, 5.0, one)
    t = time.perf_counter(); r = g.verify_cnsm_ed(q, 5.0, 1.5, 5.0, one); wall = time.perf_counter() - t
    grid = [(a, min(a + 2047, lo + c - 1)) for a in range(lo, lo + c, 2048)]
    g.verify_cnsm_ed(q, 5.0, 1.5, 5.0, grid)
    r2 = g.verify_cnsm_ed(q, 5.0, 1.5, 5.0, grid)
    line = f"c = {c:>10d}: one chain {r.kernel_ms:8.3f} ms (stream {r.stage_ms[0]:.3f}, re-walk {r.stage_ms[1]:.3f}, exact {r.stage_ms[2]:.3f}; {r.n_rewalked} windows re-walked), chains of 2048: {r2.kernel_ms:.3f} ms"
    if c <= 10_000_000:
        t = time.perf_counter(); e = kvm_oracle.verify_cnsm_ed(s, q, 5.0, 1.5, 5.0, one); cpu = time.perf_counter() - t
        ok = r.offsets.tolist() == e.offsets.tolist() and r.distances.tolist() == e.distances.tolist()
        line += f"; CPU oracle (1 core) {1e3 * cpu:.1f} ms, parity {ok}"
    print(line, flush=True)
