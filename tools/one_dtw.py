"""One cNSM-DTW query (config 4 shape) for profiling the band DTW kernel.  usage: one_dtw.py [n] [eps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kvmatch_b200
from kvmatch_b200 import datagen
import bench
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
eps = float(sys.argv[2]) if len(sys.argv) > 2 else 5.0
s = datagen.generate(n); g = kvmatch_b200.GpuSeries(0); g.load(s)
off = bench.query_offsets(n, 2048, 10)[0]
q = s[off - 1:off - 1 + 2048].copy()
iv = datagen.chain_intervals(n, 2048, 2048)
for _ in range(2):
    r = g.verify_cnsm_dtw(q, eps, 102, 1.5, 5.0, iv)
print(r.kernel_ms, r.stage_ms, r.n_lb_pass, r.count)
