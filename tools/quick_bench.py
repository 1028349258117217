"""Developer timing sweep (not the contract bench): cNSM-ED full scan kernel time vs chain chunk size."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kvmatch_b200
from kvmatch_b200 import datagen

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
chunks = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4096, 8192, 16384, 100000 - m + 1]
eps = float(sys.argv[4]) if len(sys.argv) > 4 else 5.0
t = time.time(); s = datagen.generate(n); print("gen %.1fs" % (time.time() - t), flush=True)
g = kvmatch_b200.GpuSeries(0)
t = time.time(); g.load(s); print("load %.2fs" % (time.time() - t), flush=True)
rng = np.random.default_rng(20260117)
offs = rng.integers(1, n - m, 4)
for chunk in chunks:
    iv = datagen.chain_intervals(n, m, chunk)
    for off in offs:
        q = s[off - 1:off - 1 + m].copy()
        best = 1e9
        for rep in range(3):
            t = time.perf_counter()
            r = g.verify_cnsm_ed(q, eps, 1.5, 5.0, iv)
            wall = (time.perf_counter() - t) * 1e3
            best = min(best, r.kernel_ms)
        print(f"chunk {chunk:6d} K {len(iv):6d} off {off:9d} kernel {best:8.3f} ms wall {wall:8.3f} ms  "
              f"{r.n_verified / best / 1e6:9.1f} Gsubseq/s  frac {8 * n / (best * 1e-3) / 6553e9:5.3f}  "
              f"gate {r.n_gate_pass} exact {r.n_exact} answers {r.count} stages {r.stage_ms[0]:.3f}/{r.stage_ms[1]:.3f}/{r.stage_ms[2]:.3f} "
              f"rewalked {r.n_rewalked} in {r.n_chains_rewalked} chains", flush=True)
iv = datagen.chain_intervals(n, m, 100000 - m + 1)
q = s[offs[0] - 1:offs[0] - 1 + m].copy()
r = g.verify_ed(q, 10.0, iv); r = g.verify_ed(q, 10.0, iv)
print(f"RSM-ED full scan kernel {r.kernel_ms:.3f} ms frac {8 * n / (r.kernel_ms * 1e-3) / 6553e9:.3f} answers {r.count}")
