"""Per-query stage breakdown for the bench workload's 10 seeded queries."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200, bench
from kvmatch_b200 import datagen
n, m = bench.N_PER_GPU, bench.M
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else bench.DEFAULT_CHUNK
s = datagen.generate(n); g = kvmatch_b200.GpuSeries(0); g.load(s)
iv = datagen.chain_intervals(n, m, chunk)
for off in bench.query_offsets(n, m, bench.N_QUERIES):
    q = s[off - 1:off - 1 + m].copy()
    for _ in range(2): r = g.verify_cnsm_ed(q, bench.EPSILON, bench.ALPHA, bench.BETA, iv)
    print(f"off {off:9d} kernel {r.kernel_ms:7.3f} stages {r.stage_ms[0]:.3f}/{r.stage_ms[1]:.3f}/{r.stage_ms[2]:.3f} gate {r.n_gate_pass:9d} exact {r.n_exact} answers {r.count}", flush=True)
