"""Per-query stage breakdown of the bench workload's 10 seeded queries.  usage: per_query.py [n] [m] [chunk] [eps] [engine ed|dtw] [rho]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200, bench
from kvmatch_b200 import datagen
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else bench.M
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else bench.DEFAULT_CHUNK
eps = float(sys.argv[4]) if len(sys.argv) > 4 else bench.EPSILON
engine = sys.argv[5] if len(sys.argv) > 5 else "ed"
rho = int(sys.argv[6]) if len(sys.argv) > 6 else int(0.05 * m)
s = datagen.generate_range(n, 0, n, bench.SEED); g = kvmatch_b200.GpuSeries(0); g.load(s)
iv = datagen.chain_intervals(n, m, chunk)
for off in bench.query_offsets(n, m, bench.N_QUERIES):
    q = s[off - 1:off - 1 + m].copy()
    for _ in range(2):
        r = g.verify_cnsm_ed(q, eps, bench.ALPHA, bench.BETA, iv) if engine == "ed" else g.verify_cnsm_dtw(q, eps, rho, bench.ALPHA, bench.BETA, iv)
    st = "/".join(f"{x:.3f}" for x in r.stage_ms)
    print(f"off {off:9d} kernel {r.kernel_ms:9.3f} stages {st} gate {r.n_gate_pass:9d} rewalked {r.n_rewalked} in {r.n_chains_rewalked} chains "
          f"exact {r.n_exact} lb-pass {r.n_lb_pass} cells {r.n_dtw_cells} answers {r.count}", flush=True)
