"""BASELINE configs[0]/[1](ii): index-pruned RSM-ED queries on REAL phase-1 output.  Builds the five KV-indexes on the
GPU (one fused window-mean pass), runs phase 0/1 on the host (kvmatch_b200/phase1.py), verifies the candidate list on
the GPU and on the CPU oracle (1 core).  usage: pruned_real.py n [length eps]..."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, kvmatch_b200
from kvmatch_b200 import datagen, phase1
from oracle import kvm_oracle as o

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
cfgs = [(8192, 10.0), (1024, 3.0), (512, 1.0)]
s = datagen.generate(n)
g = kvmatch_b200.GpuSeries(0); g.load(s)
t = time.perf_counter(); images = kvmatch_b200.IndexBuilder(g).build_all(); t_build = time.perf_counter() - t
print(f"n={n}: index build (5 widths, GPU pass + host step 2 + images) {t_build:.2f} s, {sum(len(v) for v in images.values())} bytes", flush=True)
indexes = [phase1.IndexFile(images[w]) for w in phase1.WU_LIST]
rng = np.random.default_rng(7)
print("| length | eps | intervals | candidates | phase-1 ms (Python mirror) | GPU kernel ms | GPU wall ms | CPU oracle ms (1 core) | answers | parity |")
print("|---|---|---|---|---|---|---|---|---|---|")
for length, eps in cfgs:
    for off in rng.integers(1, n - length, 3):
        q = s[off - 1:off - 1 + length].copy()
        t = time.perf_counter(); valid, last_seg, plan = phase1.phase1(q, eps, n, indexes); t_p1 = (time.perf_counter() - t) * 1e3
        shift = (last_seg - 1) * 25
        g.verify_ed(q, eps, valid, shift)
        t = time.perf_counter(); r = g.verify_ed(q, eps, valid, shift); wall = (time.perf_counter() - t) * 1e3
        t = time.perf_counter(); e = o.verify_ed(s, q, eps, valid, shift); cpu = (time.perf_counter() - t) * 1e3
        ok = r.offsets.tolist() == e.offsets.tolist() and r.distances.tolist() == e.distances.tolist()
        print(f"| {length} | {eps} | {len(valid)} | {r.cnt_candidate} | {t_p1:.1f} | {r.kernel_ms:.3f} | {wall:.3f} | {cpu:.1f} | {r.count} | {ok} |", flush=True)
