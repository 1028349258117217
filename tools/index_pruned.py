"""BASELINE configs[1] (ii): INDEX-PRUNED cNSM-ED queries — the reference's whole query() with phase 2 on the GPU.

Builds the five KV-indexes of one n-sample series on the GPU (one fused window-mean pass + host step 2 / file image),
then for each seeded query: phases 0 / 1 on the host (kvmatch_b200/phase1.phase1_norm over kvm_norm_intervals_* and the
library's row decoder), phase 2 on the GPU over the resulting interval list (kvm_verify_cnsm_ed with
shift = (lastSegment - 1) * 25), and the same query as an index-free full scan for comparison.  Checks: the index-pruned
answer offsets equal the full scan's (no false dismissals).  Writes a markdown table.

usage: python tools/index_pruned.py [n=1e8] [queries=4] [eps=5] [out=gpurun_out/index_pruned_r02.md]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kvmatch_b200, bench
from kvmatch_b200 import datagen, phase1

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
n_q = int(sys.argv[2]) if len(sys.argv) > 2 else 4
eps = float(sys.argv[3]) if len(sys.argv) > 3 else bench.EPSILON
out = sys.argv[4] if len(sys.argv) > 4 else "gpurun_out/index_pruned_r02.md"
m = bench.M
s = datagen.generate_range(n, 0, n, bench.SEED)
g = kvmatch_b200.GpuSeries(0)
g.load(s)
t0 = time.perf_counter()
images = kvmatch_b200.IndexBuilder(g).build_all()
build_s = time.perf_counter() - t0          # first call: includes allocating the pinned staging for the runs
t0 = time.perf_counter()
runs = g.window_mean_runs_all()
runs_s = time.perf_counter() - t0           # staging in place
kernel_ms = runs.kernel_ms
del runs
t0 = time.perf_counter()
kvmatch_b200.IndexBuilder(g).build_all()
build2_s = time.perf_counter() - t0
t0 = time.perf_counter()
indexes = [phase1.IndexFile(images[w]) for w in phase1.WU_LIST]
open_s = time.perf_counter() - t0
full_iv = datagen.chain_intervals(n, m, bench.DEFAULT_CHUNK)
rows = []
RSM_EPS = 10.0
DTW_NORM_EPS = 2.0
for off in bench.query_offsets(n, m, bench.N_QUERIES)[:n_q]:
    q = s[off - 1:off - 1 + m].copy()
    rho = int(0.05 * m)
    rsm_iv = datagen.chain_intervals(n, m, 100_000 - m + 1)
    engines = {
        "cNSM-ED": (lambda I, sh=0: g.verify_cnsm_ed(q, eps, bench.ALPHA, bench.BETA, I, sh),
                    lambda: phase1.phase1_norm(q, eps, bench.ALPHA, bench.BETA, n, indexes), full_iv),
        "RSM-ED": (lambda I, sh=0: g.verify_ed(q, RSM_EPS, I, sh), lambda: phase1.phase1(q, RSM_EPS, n, indexes),
                   np.array([[1, n - m + 1]], dtype=np.int32)),
        "cNSM-DTW": (lambda I, sh=0: g.verify_cnsm_dtw(q, DTW_NORM_EPS, rho, bench.ALPHA, bench.BETA, I, sh),
                     lambda: phase1.phase1_norm_dtw(q, DTW_NORM_EPS, rho, bench.ALPHA, bench.BETA, n, indexes), full_iv),
        "RSM-DTW": (lambda I, sh=0: g.verify_dtw(q, RSM_EPS, rho, I, sh), lambda: phase1.phase1_dtw(q, RSM_EPS, rho, n, indexes), rsm_iv),
    }
    for engine, (verify, run_phase1, scan_iv) in engines.items():
        norm = engine.startswith("cNSM")
        t0 = time.perf_counter()
        valid, last_segment, plan = run_phase1()
        t1_ms = 1e3 * (time.perf_counter() - t0)
        iv = np.asarray(valid, dtype=np.int32).reshape(-1, 2)
        shift = (last_segment - 1) * 25
        verify(iv, shift)
        t0 = time.perf_counter()
        r = verify(iv, shift)
        t2_ms = 1e3 * (time.perf_counter() - t0)
        verify(scan_iv)
        t0 = time.perf_counter()
        f = verify(scan_iv)
        tf_ms = 1e3 * (time.perf_counter() - t0)
        assert r.offsets.tolist() == f.offsets.tolist(), (engine, off)
        assert off in r.offsets.tolist(), (engine, off)
        if not norm:
            assert r.distances.tolist() == f.distances.tolist(), off   # no running statistics: bit-identical distances
        lens = iv[:, 1] - iv[:, 0] + 1
        row = (off, len(plan), last_segment, t1_ms, len(iv), int(lens.sum()), int(lens.max()), r.kernel_ms, t2_ms, f.kernel_ms, tf_ms, r.count, engine)
        rows.append(row)
        print(row, flush=True)
with open(out, "w") as fh:
    fh.write(f"# Index-pruned queries (BASELINE configs[1] (ii)), n = {n}, m = {m}; cNSM-ED: eps = {eps}, alpha = {bench.ALPHA}, beta = {bench.BETA}; cNSM-DTW: eps = {DTW_NORM_EPS}, rho = {int(0.05 * m)}; RSM-ED / RSM-DTW: eps = {RSM_EPS} (round 2)\n\n"
             f"`python tools/index_pruned.py {n} {n_q} {eps}` on one B200.  Index build (five widths: one fused window-mean pass on the GPU, "
             f"runs to the host, step 2 + file images on the host, the five widths on five host threads): {build_s:.2f} s for the first build "
             f"(it allocates the pinned staging of the runs), {build2_s:.2f} s for a second one, of which the GPU pass is {kernel_ms:.2f} ms and the pass "
             f"with its runs copied out to numpy arrays {runs_s:.2f} s; {sum(len(b) for b in images.values()) / 1e6:.0f} MB of index files; "
             f"opening them (offset + statistic tables): {open_s:.2f} s.  T_1 = phases 0 / 1 on the host (plan DP on one core, a cNSM query's segments probed on up to four threads, "
             f"`kvm_norm_intervals_*` / `kvm_intervals_*`); T_2 = `kvm_verify_*` over the phase-1 interval list with host buffers (wall) and its CUDA-event "
             f"kernel time; full scan = the same query over every window start (cNSM: chains of {bench.DEFAULT_CHUNK}; RSM-ED: one interval; RSM-DTW: chains of 100000 - m + 1).  Every row: index-pruned answer "
             f"offsets == full-scan answer offsets.\n\n"
             "| engine | query offset | segments | lastSegment | T_1 host ms | intervals | candidates | longest interval | T_2 kernel ms | T_2 wall ms | full-scan kernel ms | full-scan wall ms | answers |\n"
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        fh.write(f"| {r[12]} | {r[0]} | {r[1]} | {r[2]} | {r[3]:.0f} | {r[4]} | {r[5]} ({100.0 * r[5] / n:.2g} % of n) | {r[6]} | {r[7]:.3f} | {r[8]:.3f} | {r[9]:.3f} | {r[10]:.3f} | {r[11]} |\n")
print(open(out).read())
