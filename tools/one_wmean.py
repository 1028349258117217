"""ncu capture target: the fused window-mean pass.  usage: one_wmean.py n reps"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kvmatch_b200
from kvmatch_b200 import datagen
n = int(float(sys.argv[1])); reps = int(sys.argv[2])
s = datagen.generate(n); g = kvmatch_b200.GpuSeries(0); g.load(s)
for i in range(reps):
    r = g.window_mean_runs_all()
print(r.kernel_ms, r.n_runs, r.n_chains_rewalked)
