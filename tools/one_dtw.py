"""One cNSM-DTW query (config 4 shape) for profiling.  usage: one_dtw.py [n] [eps] [query offset] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kvmatch_b200
from kvmatch_b200 import datagen
import bench
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
eps = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
s = datagen.generate_range(n, 0, n, bench.SEED); g = kvmatch_b200.GpuSeries(0); g.load(s)
off = int(sys.argv[3]) if len(sys.argv) > 3 else bench.query_offsets(n, 2048, 10)[0]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
q = s[off - 1:off - 1 + 2048].copy()
iv = datagen.chain_intervals(n, 2048, 2048)
for _ in range(reps):
    r = g.verify_cnsm_dtw(q, eps, 102, 1.5, 5.0, iv)
print(r.kernel_ms, r.stage_ms, r.n_lb_pass, r.count)
