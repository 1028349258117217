"""Stage breakdown of phase 2 over index-pruned interval lists (cNSM-ED, n = 1e8 by default).  usage: pruned_stages.py [n]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kvmatch_b200, bench
from kvmatch_b200 import datagen, phase1
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
m = bench.M
s = datagen.generate_range(n, 0, n, bench.SEED)
g = kvmatch_b200.GpuSeries(0); g.load(s)
images = kvmatch_b200.IndexBuilder(g).build_all()
indexes = [phase1.open_index(images[w]) for w in phase1.WU_LIST]
for off in bench.query_offsets(n, m, bench.N_QUERIES)[:4]:
    q = s[off - 1:off - 1 + m].copy()
    valid, last, _ = phase1.phase1_norm(q, bench.EPSILON, bench.ALPHA, bench.BETA, n, indexes)
    iv = np.asarray(valid, dtype=np.int32).reshape(-1, 2)
    for _ in range(3):
        t0 = time.perf_counter()
        r = g.verify_cnsm_ed(q, bench.EPSILON, bench.ALPHA, bench.BETA, iv, (last - 1) * 25)
        wall = 1e3 * (time.perf_counter() - t0)
    print(f"off {off}: intervals {len(iv)} candidates {int((iv[:,1]-iv[:,0]+1).sum())} kernel {r.kernel_ms:.3f} stages {[round(x,3) for x in r.stage_ms]} "
          f"launches {r.n_launches} wall {wall:.3f} gate {r.n_gate_pass} rewalked {r.n_rewalked} exact {r.n_exact} h2d {r.h2d_bytes}", flush=True)
