"""Where the host time of one sharded cNSM-ED step goes (rank 0 of `world` on one GPU, no collective).
usage: host_overhead.py [world=8] [n=1e9]"""
import os, sys, time
os.environ.setdefault("KVM_PLAN_CACHE", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200, bench
from kvmatch_b200 import datagen, sharding
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000_000
M, chunk = bench.M, bench.DEFAULT_CHUNK
sh = sharding.make_shard(n, M, 0, world, grid=chunk)
local = datagen.generate_range(n, sh.first - 1, sh.last, bench.SEED)
g = kvmatch_b200.GpuSeries(0); g.load(local, n=n, first=sh.first)
iv = sharding.assign_intervals(datagen.chain_intervals(n, M, chunk, lo=sh.start_lo, hi=min(sh.start_hi, n - M + 1)), 0, M, sh)
qs = [bench.query_of(n, o, M) for o in bench.query_offsets(n, M, 10)]
for i in range(5): g.verify_cnsm_ed(qs[i], 5.0, 1.5, 5.0, iv)
ts, ks = [], []
for i in range(40):
    t = time.perf_counter(); r = g.verify_cnsm_ed(qs[i % 10], 5.0, 1.5, 5.0, iv); ts.append(time.perf_counter() - t); ks.append(r.kernel_ms)
print(f"world {world}: K {len(iv)} intervals; wall/step {1e3*np.mean(ts):.3f} ms, device {np.mean(ks):.3f} ms, host overhead {1e3*np.mean(ts)-np.mean(ks):.3f} ms")
os.environ["KVM_TIMING"] = "1"
