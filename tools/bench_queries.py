"""Developer timing: the bench's 10 seeded cNSM-ED queries (n, m, chunk, eps from argv), per-stage kernel times."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kvmatch_b200
from kvmatch_b200 import datagen
import bench

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
chunks = [int(c) for c in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2048]
eps = float(sys.argv[4]) if len(sys.argv) > 4 else 5.0
s = datagen.generate(n)
g = kvmatch_b200.GpuSeries(0)
g.load(s)
offs = bench.query_offsets(n, m, 10)
for chunk in chunks:
    iv = datagen.chain_intervals(n, m, chunk)
    tot = 0.0
    st = np.zeros(4)
    for off in offs:
        q = s[off - 1:off - 1 + m].copy()
        best = None
        for rep in range(3):
            t = time.perf_counter()
            r = g.verify_cnsm_ed(q, eps, 1.5, 5.0, iv)
            wall = (time.perf_counter() - t) * 1e3
            if best is None or r.kernel_ms < best.kernel_ms:
                best = r
        tot += best.kernel_ms
        st += np.array(best.stage_ms)
        print(f"chunk {chunk:6d} off {off:9d} kernel {best.kernel_ms:7.3f} ms wall {wall:7.3f}  stages "
              f"{best.stage_ms[0]:.3f}/{best.stage_ms[1]:.3f}/{best.stage_ms[2]:.3f}  gate {best.n_gate_pass:9d} answers {best.count:4d} "
              f"rewalked {best.n_rewalked} in {best.n_chains_rewalked} chains", flush=True)
    print(f"chunk {chunk}: mean kernel {tot / 10:.3f} ms  stages {st[0] / 10:.3f}/{st[1] / 10:.3f}/{st[2] / 10:.3f}  "
          f"stream frac {8 * n / (st[0] / 10 * 1e-3) / 6553e9:.3f} step frac {8 * n / (tot / 10 * 1e-3) / 6553e9:.3f}", flush=True)
