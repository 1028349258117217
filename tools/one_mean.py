import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kvmatch_b200
from kvmatch_b200 import datagen
n = int(float(sys.argv[1])); w = int(sys.argv[2])
s = datagen.generate(n); g = kvmatch_b200.GpuSeries(0); g.load(s)
for i in range(2): keys, first, last, ms, nl = g.window_mean_runs(w)
print(ms, len(keys))
