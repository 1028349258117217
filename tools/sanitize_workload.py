import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import numpy as np, kvmatch_b200
from kvmatch_b200 import datagen
from oracle import kvm_oracle as o
n, m = 60_000, 128
s = datagen.generate(n, seed=3)
g = kvmatch_b200.GpuSeries(0); g.load(s)
q = s[20_000:20_000 + m].copy()
iv = datagen.chain_intervals(n, m, 3000)
r = g.verify_cnsm_ed(q, 4.0, 1.5, 5.0, iv); e = o.verify_cnsm_ed(s, q, 4.0, 1.5, 5.0, iv)
assert r.offsets.tolist() == e.offsets.tolist() and r.distances.tolist() == e.distances.tolist(), "cnsm-ed"
r = g.verify_cnsm_dtw(q, 3.0, 6, 1.5, 5.0, iv); e = o.verify_cnsm_dtw(s, q, 3.0, 6, 1.5, 5.0, iv)
assert r.offsets.tolist() == e.offsets.tolist(), "cnsm-dtw"
r = g.verify_ed(q, 30.0, [(1, n - m + 1)]); e = o.verify_ed(s, q, 30.0, [(1, n - m + 1)])
assert r.offsets.tolist() == e.offsets.tolist(), "ed"
k, f, l, ms, nl = g.window_mean_runs(50); ek, ef, el = o.window_mean_runs(s, 50)
assert f.tolist() == ef.tolist() and l.tolist() == el.tolist(), "runs"
qs = np.stack([s[o:o + m] for o in (100, 20_000, 40_000)])
rb = g.verify_cnsm_ed_batch(qs, 4.0, 1.5, 5.0, iv)
for qq, r1 in zip(qs, rb):
    one = g.verify_cnsm_ed(qq, 4.0, 1.5, 5.0, iv)
    assert r1.offsets.tolist() == one.offsets.tolist() and r1.distances.tolist() == one.distances.tolist(), "query set"
print("sanitizer workload ok")
