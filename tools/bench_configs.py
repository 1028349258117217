"""Measures the other BASELINE.json configs (kernel time from the library's CUDA events, CPU oracle beside it on a
bounded sample) and prints a markdown table for profiles/.  usage: bench_configs.py [n_dtw=1e8] [big_n=0]"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kvmatch_b200
from kvmatch_b200 import datagen
from oracle import kvm_oracle

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
big_n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 0
FP64_DADD = 18.36e12  # measured, tools/fp64_peak.cu
rows = []

def timed(fn, reps=3):
    best = None
    for _ in range(reps):
        r = fn()
        if best is None or r.kernel_ms < best.kernel_ms:
            best = r
    t = time.perf_counter(); fn(); wall = (time.perf_counter() - t) * 1e3
    return best, wall

def cpu_sample(fn_cpu, sample_n):
    t = time.perf_counter(); r = fn_cpu(); dt = time.perf_counter() - t
    return r, dt

s = datagen.generate(n)
g = kvmatch_b200.GpuSeries(0); g.load(s)
rng = np.random.default_rng(7)
sample_n = 5_000_000

# config 1 (README demo): n=1e6 prefix, offset 123456, m=8192, eps 10, full scan
m = 8192; s1 = s[:1_000_000]; g1 = kvmatch_b200.GpuSeries(0); g1.load(s1)
q = s1[123455:123455 + m].copy(); iv = [(1, len(s1) - m + 1)]
r, wall = timed(lambda: g1.verify_ed(q, 10.0, iv))
ro, dt = cpu_sample(lambda: kvm_oracle.verify_ed(s1, q, 10.0, iv), 0)
ok = r.offsets.tolist() == ro.offsets.tolist() and r.distances.tolist() == ro.distances.tolist()
rows.append(("1 RSM-ED n=1e6 m=8192 eps=10 (README demo, full scan)", r.n_verified, r.kernel_ms, wall, r.count, ro.n_verified / dt, f"best {r.offsets[np.argmin(r.distances)]} dist {r.distances.min()}", ok))
g1.close()

# RSM-ED full scan at n (config 5 flavour)
m = 1024; q = s[40_000_000:40_000_000 + m].copy(); iv = datagen.chain_intervals(n, m, 100000 - m + 1)
r, wall = timed(lambda: g.verify_ed(q, 10.0, iv))
civ = datagen.chain_intervals(sample_n, m, 100000 - m + 1)
ro, dt = cpu_sample(lambda: kvm_oracle.verify_ed(s[:sample_n], q, 10.0, civ), sample_n)
rows.append((f"5 RSM-ED n={n:.0e} m=1024 eps=10 full scan", r.n_verified, r.kernel_ms, wall, r.count, ro.n_verified / dt, f"HBM frac {8*n/(r.kernel_ms*1e-3)/6553e9:.3f}", None))

# config 3: RSM-DTW m=512 rho=25
m, rho = 512, 25
q = s[60_000_000:60_000_000 + m] + rng.normal(scale=0.05, size=m); iv = datagen.chain_intervals(n, m, 100000 - m + 1)
for eps in (50.0, 75.0, 100.0):
    r, wall = timed(lambda: g.verify_dtw(q, eps, rho, iv), reps=2)
    civ = datagen.chain_intervals(sample_n, m, 100000 - m + 1)
    ro, dt = cpu_sample(lambda: kvm_oracle.verify_dtw(s[:sample_n], q, eps, rho, civ), sample_n)
    keep = r.offsets <= sample_n - m + 1
    ok = r.offsets[keep].tolist() == ro.offsets.tolist() and r.distances[keep].tolist() == ro.distances.tolist()
    cells = r.n_lb_pass * (m * (2 * rho + 1) - rho * (rho + 1))
    rows.append((f"3 RSM-DTW n={n:.0e} m=512 rho=25 eps={eps:g}", r.n_verified, r.kernel_ms, wall, r.count, ro.n_verified / dt,
                 f"DTWs {r.n_lb_pass} (ref cascade on sample: {ro.n_dtw}); stage ms {r.stage_ms[0]:.2f}/{r.stage_ms[2]:.2f}; FP64 frac {5*cells/(max(r.stage_ms[2],1e-6)*1e-3)/FP64_DADD:.3f}", ok))

# config 4 flavour on one GPU: cNSM-DTW m=2048 rho=102
m, rho = 2048, 102
q = s[70_000_000:70_000_000 + m].copy(); iv = datagen.chain_intervals(n, m, 12288)
for eps in (1.0, 5.0, 10.0):
    r, wall = timed(lambda: g.verify_cnsm_dtw(q, eps, rho, 1.5, 5.0, iv), reps=2)
    civ = datagen.chain_intervals(sample_n, m, 12288)
    ro, dt = cpu_sample(lambda: kvm_oracle.verify_cnsm_dtw(s[:sample_n], q, eps, rho, 1.5, 5.0, civ), sample_n)
    keep = r.offsets <= sample_n - m + 1
    ok = r.offsets[keep].tolist() == ro.offsets.tolist() and r.distances[keep].tolist() == ro.distances.tolist()
    cells = r.n_lb_pass * (m * (2 * rho + 1) - rho * (rho + 1))
    rows.append((f"4 cNSM-DTW n={n:.0e} m=2048 rho=102 eps={eps:g}", r.n_verified, r.kernel_ms, wall, r.count, ro.n_verified / dt,
                 f"gate {r.n_gate_pass} DTWs {r.n_lb_pass} (ref on sample: {ro.n_dtw}); stage ms {r.stage_ms[0]:.2f}/{r.stage_ms[1]:.2f}/{r.stage_ms[2]:.2f}; FP64 frac {5*cells/(max(r.stage_ms[2],1e-6)*1e-3)/FP64_DADD:.3f}", ok))

# index build, n=1e6 prefix (config 1's IndexBuilder)
g2 = kvmatch_b200.GpuSeries(0); g2.load(s[:1_000_000])
tot_ms = 0; okr = True; t_cpu = 0
for w in (25, 50, 100, 200, 400):
    keys, first, last, ms, _ = g2.window_mean_runs(w); keys, first, last, ms, _ = g2.window_mean_runs(w)
    tot_ms += ms
    t = time.perf_counter(); ek, ef, el = kvm_oracle.window_mean_runs(s[:1_000_000], w); t_cpu += time.perf_counter() - t
    okr &= first.tolist() == ef.tolist() and last.tolist() == el.tolist() and keys.view(np.int64).tolist() == ek.view(np.int64).tolist()
rows.append(("1 IndexBuilder step 1, n=1e6, 5 windows", 5_000_000, tot_ms, tot_ms, 0, 5e6 / t_cpu, "windows/s; unit = window means", okr))
# whole single-width build (steps 1+2 + file image), 5 widths
import tempfile
tot_k = tot_wall = t_cpu = 0.0; okf = True; nbytes = 0
with tempfile.TemporaryDirectory() as td:
    for w in (25, 50, 100, 200, 400):
        path = os.path.join(td, f"index-1000000-{w}")
        g2.build_index_file(w, path)
        t = time.perf_counter(); info = g2.build_index_file(w, path); tot_wall += (time.perf_counter() - t) * 1e3
        tot_k += info.kernel_ms
        t = time.perf_counter(); exp, _, _ = kvm_oracle.index_file_image(s[:1_000_000], w); t_cpu += time.perf_counter() - t
        okf &= open(path, "rb").read() == exp
        nbytes += len(exp)
rows.append(("1 IndexBuilder steps 1+2 + index files, n=1e6, 5 windows", 5_000_000, tot_k, tot_wall, 0, 5e6 / t_cpu,
             f"windows/s; kernel = window-mean pass, wall adds host step 2 + codec + write ({nbytes} file bytes, byte-identical)", okf))
g2.close()

print("| config | verified | kernel ms | wall ms | subseq/s (GPU kernel) | #answers | CPU oracle 1 core subseq/s | note | parity on CPU sample |")
print("|---|---|---|---|---|---|---|---|---|")
for name, v, kms, wall, cnt, cpu, note, ok in rows:
    print(f"| {name} | {v} | {kms:.3f} | {wall:.3f} | {v/(kms*1e-3):.3e} | {cnt} | {cpu:.3e} | {note} | {ok} |")

if big_n:
    del s; g.close()
    m = 2048
    t = time.perf_counter(); sb = datagen.generate(big_n); print(f"\ngen n={big_n:.0e}: {time.perf_counter()-t:.1f}s")
    gb = kvmatch_b200.GpuSeries(0); t = time.perf_counter(); gb.load(sb); print(f"load {time.perf_counter()-t:.2f}s")
    iv = datagen.chain_intervals(big_n, m, 100000 - m + 1)
    lat = []
    for off in rng.integers(1, big_n - m, 6):
        q = sb[off - 1:off - 1 + m].copy()
        t = time.perf_counter(); r = gb.verify_cnsm_dtw(q, 5.0, 102, 1.5, 5.0, iv); lat.append((time.perf_counter() - t) * 1e3)
        print(f"cNSM-DTW n={big_n:.0e} m=2048 rho=102 eps=5 off {off}: kernel {r.kernel_ms:.2f} ms wall {lat[-1]:.2f} ms answers {r.count} gate {r.n_gate_pass} dtws {r.n_lb_pass} stages {r.stage_ms[0]:.2f}/{r.stage_ms[1]:.2f}/{r.stage_ms[2]:.2f}", flush=True)
    print(f"p50 query latency n={big_n:.0e}: {np.median(lat[1:]):.2f} ms")
    m = 1024; iv = datagen.chain_intervals(big_n, m, 100000 - m + 1); lat = []
    for off in rng.integers(1, big_n - m, 6):
        q = sb[off - 1:off - 1 + m].copy()
        t = time.perf_counter(); r = gb.verify_cnsm_ed(q, 5.0, 1.5, 5.0, iv); lat.append((time.perf_counter() - t) * 1e3)
        print(f"cNSM-ED n={big_n:.0e} m=1024 eps=5 off {off}: kernel {r.kernel_ms:.2f} ms wall {lat[-1]:.2f} ms answers {r.count} gate {r.n_gate_pass} HBM frac {8*big_n/(r.kernel_ms*1e-3)/6553e9:.3f}", flush=True)
    print(f"p50 query latency n={big_n:.0e} cNSM-ED: {np.median(lat[1:]):.2f} ms")
