"""BASELINE.json configs[4]: index-free full-scan verification sweep, query length 128-8192, ED and DTW, one GPU.

For every m in {128, 256, ..., 8192}: RSM-ED, cNSM-ED, RSM-DTW and cNSM-DTW (rho = floor(0.05 m)) over EVERY window start
of one n-sample series (chain-chunked interval list, the same list on the CPU), one seeded query per length cut from
the series itself, epsilon scaled with sqrt(m) so that selectivity stays comparable across lengths.  Each engine is also
run on the first `cpu_n` samples and compared bit for bit with the CPU oracle there (the oracle's time on that prefix,
one core, gives the CPU column).  Writes a markdown table.

usage: python tools/sweep_cfg5.py [n=1e9] [cpu_n=2e6] [out=gpurun_out/cfg5_r02.md]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import kvmatch_b200, bench
from kvmatch_b200 import datagen
from oracle import kvm_oracle

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
cpu_n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 2_000_000
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/cfg5_r02.md"
ALPHA, BETA, CHUNK = 1.5, 5.0, 2048
s = datagen.generate_range(n, 0, n, bench.SEED)
g = kvmatch_b200.GpuSeries(0)
g.load(s)
gp = kvmatch_b200.GpuSeries(0)       # the CPU-sized prefix, for the parity column
gp.load(s[:cpu_n])
rng = np.random.default_rng(55)
rows = []
for m in (128, 256, 512, 1024, 2048, 4096, 8192):
    off = int(rng.integers(cpu_n // 4, cpu_n // 2))      # inside the prefix: the planted match is checked on the CPU too
    q = s[off - 1:off - 1 + m].copy()
    rho = int(0.05 * m)
    sc = float(np.sqrt(m / 1024.0))
    iv = datagen.chain_intervals(n, m, CHUNK)
    ivp = datagen.chain_intervals(cpu_n, m, CHUNK)
    engines = [
        ("RSM-ED", 10.0 * sc, lambda G, e, I: G.verify_ed(q, e, I), lambda e, I: kvm_oracle.verify_ed(s[:cpu_n], q, e, I)),
        ("cNSM-ED", 5.0 * sc, lambda G, e, I: G.verify_cnsm_ed(q, e, ALPHA, BETA, I),
         lambda e, I: kvm_oracle.verify_cnsm_ed(s[:cpu_n], q, e, ALPHA, BETA, I)),
        ("RSM-DTW", 10.0 * sc, lambda G, e, I: G.verify_dtw(q, e, rho, I), lambda e, I: kvm_oracle.verify_dtw(s[:cpu_n], q, e, rho, I)),
        ("cNSM-DTW", 2.0 * sc, lambda G, e, I: G.verify_cnsm_dtw(q, e, rho, ALPHA, BETA, I),
         lambda e, I: kvm_oracle.verify_cnsm_dtw(s[:cpu_n], q, e, rho, ALPHA, BETA, I)),
    ]
    for name, eps, run_gpu, run_cpu in engines:
        run_gpu(g, eps, iv)
        t0 = time.perf_counter()
        r = run_gpu(g, eps, iv)
        wall = time.perf_counter() - t0
        rp = run_gpu(gp, eps, ivp)
        t0 = time.perf_counter()
        e = run_cpu(eps, ivp)
        cpu_s = time.perf_counter() - t0
        same = rp.offsets.tolist() == e.offsets.tolist() and rp.distances.tolist() == e.distances.tolist()
        assert off in r.offsets.tolist(), (name, m)
        ed = "ED" in name
        row = (m, rho if not ed else "-", name, eps, r.kernel_ms, 1e3 * wall, r.n_verified / (r.kernel_ms * 1e-3),
               8.0 * n / (r.kernel_ms * 1e-3) / 1e9 / 6553.0 if ed else float("nan"), r.count, r.n_lb_pass if not ed else r.n_exact,
               e.n_verified / cpu_s, (r.n_verified / (r.kernel_ms * 1e-3)) / (e.n_verified / cpu_s), same)
        rows.append(row)
        print(row, flush=True)
with open(out, "w") as f:
    f.write(f"# BASELINE configs[4]: index-free full-scan sweep, n = {n}, one B200 (round 2)\n\n"
            f"`python tools/sweep_cfg5.py {n} {cpu_n}`: every window start of one series, chains of {CHUNK} candidates, alpha = {ALPHA}, beta = {BETA}, "
            f"rho = floor(0.05 m), one query per length cut from the series.  `kernel ms` = CUDA events over all stages of the call, "
            f"`wall ms` = through the C ABI with host buffers.  HBM fraction = 8 n bytes / kernel time / 6553 GB/s (ED engines).  "
            f"CPU = the oracle (C++ restatement of the reference's loops) on ONE core over the first {cpu_n} samples, same chains; "
            f"`parity` = GPU and oracle agree bit for bit (offsets and distances) on that prefix.\n\n"
            "| m | rho | engine | eps | kernel ms | wall ms | subsequences/s | HBM frac | answers | exact / DTW candidates | CPU subseq/s (1 core) | GPU/CPU | parity |\n"
            "|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for r in rows:
        f.write(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]:.2f} | {r[4]:.3f} | {r[5]:.3f} | {r[6]:.3e} | {r[7]:.3f} | {r[8]} | {r[9]} | {r[10]:.3e} | {r[11]:.0f} | {r[12]} |\n")
print(open(out).read())
