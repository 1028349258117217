"""Small all-entry-point workload for compute-sanitizer (memcheck / racecheck / synccheck): every round-2 kernel runs at
least once (stream + re-walk + screen + exact, fused LB in both paths, corner probe, both band DTW kernels, envelope,
RSM-ED, fused window means, UCR scans, query set), answers checked against the oracle in the same process."""
import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, kvmatch_b200
from kvmatch_b200 import datagen
from oracle import kvm_oracle as o

def same(a, b, what):
    assert a.offsets.tolist() == b.offsets.tolist() and a.distances.tolist() == b.distances.tolist(), what

n, m = 60_000, 128
s = datagen.generate(n, seed=3)
g = kvmatch_b200.GpuSeries(0); g.load(s)
q = s[20_000:20_000 + m].copy()
iv = datagen.chain_intervals(n, m, 3000)
same(g.verify_cnsm_ed(q, 4.0, 1.5, 5.0, iv), o.verify_cnsm_ed(s, q, 4.0, 1.5, 5.0, iv), "cnsm-ed")
same(g.verify_cnsm_dtw(q, 3.0, 6, 1.5, 5.0, iv), o.verify_cnsm_dtw(s, q, 3.0, 6, 1.5, 5.0, iv), "cnsm-dtw (warp DTW, short list)")
same(g.verify_cnsm_dtw(q, 9.0, 40, 3.0, 50.0, iv), o.verify_cnsm_dtw(s, q, 9.0, 40, 3.0, 50.0, iv), "cnsm-dtw (cohort LB path, probe)")
m2 = 512
q2 = s[30_000:30_000 + m2].copy(); iv2 = datagen.chain_intervals(n, m2, 3000)
same(g.verify_cnsm_dtw(q2, 6.0, 100, 2.0, 20.0, iv2), o.verify_cnsm_dtw(s, q2, 6.0, 100, 2.0, 20.0, iv2), "cnsm-dtw (cooperative DTW, probe K=101)")
same(g.verify_dtw(q, 12.0, 6, iv), o.verify_dtw(s, q, 12.0, 6, iv), "rsm-dtw")
same(g.verify_ed(q, 30.0, [(1, n - m + 1)]), o.verify_ed(s, q, 30.0, [(1, n - m + 1)]), "ed")
same(g.verify_ed(q, 30.0, iv), o.verify_ed(s, q, 30.0, iv), "ed (regular grid)")
k, f, l, ms, nl = g.window_mean_runs(50); ek, ef, el = o.window_mean_runs(s, 50)
assert f.tolist() == ef.tolist() and l.tolist() == el.tolist(), "runs"
res = g.window_mean_runs_all()
for w, (k5, f5, l5) in zip(kvmatch_b200.WU_LIST, res.runs):
    ek, ef, el = o.window_mean_runs(s, w)
    assert f5.tolist() == ef.tolist() and l5.tolist() == el.tolist() and k5.view(np.int64).tolist() == ek.view(np.int64).tolist(), w
same(g.scan_ucr_ed(q, 4.0, 1.5, 5.0), o.ucr_ed(s, q, 4.0, 1.5, 5.0), "ucr-ed")
same(g.scan_ucr_dtw(q, 3.0, 6, 1.5, 5.0), o.ucr_dtw(s, q, 3.0, 6, 1.5, 5.0), "ucr-dtw")
lo, up = g.envelope(6, 1000, 5000)
el_, eu_ = o.lower_upper_lemire(s[999:5999], 6)
assert lo.tolist() == el_.tolist() and up.tolist() == eu_.tolist(), "envelope"
qs = np.stack([s[x:x + m] for x in (100, 20_000, 40_000)])
for qq, r1 in zip(qs, g.verify_cnsm_ed_batch(qs, 4.0, 1.5, 5.0, iv)):
    same(r1, g.verify_cnsm_ed(qq, 4.0, 1.5, 5.0, iv), "query set")
print("sanitizer workload ok")
