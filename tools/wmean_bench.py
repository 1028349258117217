"""Developer timing of the fused window-mean pass.  usage: wmean_bench.py n"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200
from kvmatch_b200 import datagen
for n in [int(float(a)) for a in sys.argv[1:]] or [1_000_000, 100_000_000]:
    s = datagen.generate(n); g = kvmatch_b200.GpuSeries(0); g.load(s)
    for rep in range(3):
        t = time.perf_counter(); r = g.window_mean_runs_all(); wall = time.perf_counter() - t
        print(f"n {n} all five widths: kernel {r.kernel_ms:.3f} ms wall {wall*1e3:.1f} ms runs {r.n_runs} rewalked epochs {r.n_chains_rewalked} "
              f"HBM frac {(8*n + 16*r.n_runs)/(r.kernel_ms*1e-3)/6553e9:.3f} (8n only: {8*n/(r.kernel_ms*1e-3)/6553e9:.3f})", flush=True)
    t = time.perf_counter(); k = g.window_mean_runs(50); print(f"  single width 50 (relay walker): kernel {k[3]:.3f} ms wall {(time.perf_counter()-t)*1e3:.1f} ms")
    g.close()
