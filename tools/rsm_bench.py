"""RSM-ED / RSM-DTW full scans (BASELINE configs[0] shape on the GPU, configs[2]).  usage: rsm_bench.py [n] [m_ed] [m_dtw]"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200, bench
from kvmatch_b200 import datagen
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
m_ed = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
m_dtw = int(sys.argv[3]) if len(sys.argv) > 3 else 512
s = datagen.generate_range(n, 0, n, bench.SEED); g = kvmatch_b200.GpuSeries(0); g.load(s)
for off in bench.query_offsets(n, m_ed, 4):
    q = s[off - 1:off - 1 + m_ed].copy()
    for iv, name in ((datagen.chain_intervals(n, m_ed, 2048), "chunked 2048"), ([(1, n - m_ed + 1)], "one interval")):
        g.verify_ed(q, 10.0, iv)
        t = time.perf_counter(); r = g.verify_ed(q, 10.0, iv); wall = time.perf_counter() - t
        print(f"RSM-ED m={m_ed} eps=10 {name:13s} off {off}: kernel {r.kernel_ms:.3f} ms ({8 * n / (r.kernel_ms * 1e-3) / 6553e9:.3f} of HBM), wall {1e3 * wall:.3f} ms, answers {r.count}", flush=True)
rho = int(0.05 * m_dtw)
iv = datagen.chain_intervals(n, m_dtw, 100_000 - m_dtw + 1)
for eps in (50.0, 75.0, 100.0):
    for off in bench.query_offsets(n, m_dtw, 3):
        q = s[off - 1:off - 1 + m_dtw].copy()
        g.verify_dtw(q, eps, rho, iv)
        t = time.perf_counter(); r = g.verify_dtw(q, eps, rho, iv); wall = time.perf_counter() - t
        print(f"RSM-DTW m={m_dtw} rho={rho} eps={eps} off {off}: kernel {r.kernel_ms:.3f} ms stages {[round(x, 3) for x in r.stage_ms]}, wall {1e3 * wall:.3f} ms, "
              f"lb-pass {r.n_lb_pass} dtw {r.n_exact} cells {r.n_dtw_cells} answers {r.count}", flush=True)
