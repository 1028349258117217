#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native KV-match phase-2 path.

A "step" is one cNSM-ED verification pass (BASELINE.json configs[1]: NormQueryEngine, query length 1024,
alpha=1.5, beta=5) of one query over every window start of a DataGenerator-style synthetic series, index-free:
the merged-interval list handed to the C ABI is [1, n-m+1] cut into statistic chains of --chunk candidates
(<= 100000-m+1, the reference's own epoch chunking, K/experiments/ucr/UcrDtwQueryExecutor.java:97).

  python bench.py --gpus N --steps K --warmup W              # ours (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU loops (oracle port), rank 0 only

N GPUs: weak scaling — every rank holds its own n_per_gpu samples (+ m-1 halo) of one global series of length
N*n_per_gpu and verifies its own window starts; the only exchange is the tail (counts, sparse answers, best
match) over torch.distributed/NCCL.  `value` = window starts verified by all ranks / max-over-ranks device time
of the K timed steps (CUDA events on the library's stream, series resident in HBM).  `e2e` = the same through
the C ABI call with host buffers (query + interval list copied in, answers copied out, every step).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# Every timed step plans and uploads its own candidate-interval list, as a query of the index-based engines would:
# the library's reuse of an identical previous plan (a convenience for fixed-grid scans) is switched off here.
os.environ.setdefault("KVM_PLAN_CACHE", "0")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from kvmatch_b200 import datagen, sharding  # noqa: E402

M = 1024
ALPHA, BETA = 1.5, 5.0
EPSILON = 5.0            # middle of the reference's cNSM grid {1,5,10} (NormQueryDtwSelectivityGenerate.java:72-86)
N_PER_GPU = 100_000_000
N_QUERIES = 10           # seeded query offsets, cycled over the steps
SEED = datagen.DEFAULT_SEED
DEFAULT_CHUNK = 1024     # candidates per statistic chain (see DESIGN.md "chain chunking")


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def query_offsets(n_total, m, k):
    rng = np.random.default_rng(SEED)
    return [int(x) for x in rng.integers(1, n_total - m, k)]


def query_of(n_total, off, m):
    return datagen.generate_range(n_total, off - 1, off - 1 + m, SEED)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, device_index):
        self.idx = device_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_sample(series, q, intervals, n_threads):
    """Time the CPU oracle (the restated reference loops) on `intervals`; n_threads > 1 shards the interval list
    over Python threads (ctypes releases the GIL) — every chain is still walked by exactly one thread."""
    from oracle import kvm_oracle
    kvm_oracle.lib()
    parts = [p for p in np.array_split(np.asarray(intervals), n_threads) if len(p)]
    out = [None] * len(parts)

    def run(i):
        out[i] = kvm_oracle.verify_cnsm_ed(series, q, EPSILON, ALPHA, BETA, parts[i])

    t0 = time.perf_counter()
    threads = [threading.Thread(target=run, args=(i,)) for i in range(len(parts))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    dt = time.perf_counter() - t0
    return sum(o.n_verified for o in out), dt, out


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  It is Java (no JVM in this image), so
    this times the C++ restatement in oracle/ on the host cores, rank 0 only; other ranks exit without work."""
    if rank != 0:
        return
    n = N_PER_GPU * world
    cores = os.cpu_count() or 1
    chunk = min(args.chunk, 100000 - M + 1)
    # bounded sample: the first `sample_n` samples of the same series, same chains, same queries
    sample_n = int(args.ref_sample)
    series = datagen.generate_range(n, 0, sample_n, SEED)
    iv = datagen.chain_intervals(sample_n, M, chunk)
    offs = query_offsets(n, M, N_QUERIES)
    times, verified = [], 0
    for step in range(args.warmup + args.steps):
        q = query_of(n, offs[step % N_QUERIES], M)
        v, dt, _ = oracle_sample(series, q, iv, cores)
        if step >= args.warmup:
            times.append(dt)
            verified += v
    total = sum(times)
    value = verified / total
    line = {
        "impl": "reference", "metric": "verified subsequences/sec (cNSM-ED)", "value": value, "unit": "subsequences/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(n, chunk, world),
        "cpu_baseline": {"value": value, "unit": "subsequences/s", "cores": cores, "kind": "port",
                         "sample": f"first {sample_n} samples ({len(iv)} chains of <= {chunk} candidates) of the same "
                                   f"series per step, interval list sharded over {cores} threads"},
        "e2e": {"value": value, "unit": "subsequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference is Java 8 with no JVM in this image: timed the line-by-line C++ restatement (oracle/), "
                "series resident in RAM (kinder than the reference, which re-reads its file per interval)",
    }
    print(json.dumps(line), flush=True)


def workload_config(n_total, chunk, world):
    return {"workload": "cNSM-ED NormQueryEngine phase-2 verification, index-free scan of every window start "
                        "(BASELINE.json configs[1])",
            "n_per_gpu": N_PER_GPU, "n_total": n_total, "query_length": M, "alpha": ALPHA, "beta": BETA,
            "epsilon": EPSILON, "chain_chunk": chunk, "queries": N_QUERIES, "parallelism": f"offset-shard x{world}",
            "l2": "inputs (800 MB per GPU) exceed the 126 MB L2; no explicit flush between steps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-query-set", action="store_true", help="skip the query-set side measurement")
    ap.add_argument("--no-dtw", action="store_true", help="skip the cNSM-DTW side measurement")
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("KVM_BENCH_CHUNK", DEFAULT_CHUNK)))
    ap.add_argument("--ref-sample", type=float, default=N_PER_GPU)   # samples scanned per reference step
    ap.add_argument("--cpu-queries", type=int, default=10)          # queries the 1-core CPU baseline times
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import kvmatch_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: kvmatch_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n_total = N_PER_GPU * world
    chunk = min(args.chunk, 100000 - M + 1)
    shard = sharding.make_shard(n_total, M, rank, world, grid=chunk)
    t0 = time.perf_counter()
    local = datagen.generate_range(n_total, shard.first - 1, shard.last, SEED)
    t_gen = time.perf_counter() - t0
    g = kvmatch_b200.GpuSeries(local_rank)
    t0 = time.perf_counter()
    g.load(local, n=n_total, first=shard.first)
    t_load = time.perf_counter() - t0
    all_iv = datagen.chain_intervals(n_total, M, chunk, lo=shard.start_lo, hi=min(shard.start_hi, n_total - M + 1))
    iv = sharding.assign_intervals(all_iv, 0, M, shard)
    offs = query_offsets(n_total, M, N_QUERIES)
    queries = [query_of(n_total, o, M) for o in offs]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(i):
        t = time.perf_counter()
        r = g.verify_cnsm_ed(queries[i % N_QUERIES], EPSILON, ALPHA, BETA, iv)  # host buffers in, host answers out
        return r, time.perf_counter() - t

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        one_step(i)
    sampler.lines.clear()  # keep only samples taken from here on (the timed region)
    barrier()
    t_wall0 = time.perf_counter()
    dev_ms = walker_ms = wall_s = 0.0
    verified = launches = answers = s_total = gate = h2d = 0
    lat = []
    for i in range(args.steps):
        r, dt = one_step(args.warmup + i)
        dev_ms += r.kernel_ms
        walker_ms += r.stage_ms[0]
        wall_s += dt
        lat.append(dt)
        verified += r.n_verified
        launches += r.n_launches
        answers += r.count
        s_total += r.s_total
        gate += r.n_gate_pass
        h2d += r.h2d_bytes
        last = r
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()

    # the multi-GPU tail of one query (not inside the timed kernels): counts, answers, best match
    offs_m, dists_m, totals, best = sharding.merge_answers(last.offsets, last.distances,
                                                           {"n_verified": last.n_verified}, device=dev)

    stats = torch.tensor([dev_ms, t_wall, wall_s, walker_ms], dtype=torch.float64, device=dev)
    sums = torch.tensor([verified, launches, answers, s_total, gate], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_ms_max, t_wall_max, wall_s_max, walker_ms_max = [float(x) for x in stats.tolist()]
    verified_all, launches_all, answers_all, s_total_all, gate_all = [float(x) for x in sums.tolist()]

    if rank == 0:
        peak, peak_src = load_peaks()
        k = args.steps
        value = verified_all / (dev_ms_max * 1e-3)
        e2e_value = verified_all / wall_s_max
        # roofline of the dominant kernel (the statistics walker): algorithmic bytes = 8 B per touched sample
        # (bytes_alg = 8*S_total + 12*#answers), per launch, over its own CUDA-event duration
        # S_total counts every touched sample once: the m-1 halo re-read between adjacent chains is NOT algorithmic
        ivs = np.asarray(iv, dtype=np.int64).reshape(-1, 2)
        lo_s, hi_s = ivs[:, 0], np.minimum(ivs[:, 1] + M - 1, n_total)
        order = np.argsort(lo_s, kind="stable")
        lo_s, hi_s = lo_s[order], hi_s[order]
        reach = np.maximum.accumulate(hi_s)
        prev_reach = np.concatenate(([lo_s[0] - 1], reach[:-1]))
        unique_samples = int(np.sum(np.maximum(0, hi_s - np.maximum(lo_s - 1, prev_reach))))
        bytes_per_launch = 8.0 * unique_samples + 12.0 * answers / k     # rank 0's shard
        walker_s = (walker_ms / k) * 1e-3
        achieved = bytes_per_launch / walker_s / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "roofline_r01.json")
        if os.path.exists(prof):
            pj = json.load(open(prof))
            if pj.get("chain_chunk") == chunk and pj.get("n_per_gpu") == N_PER_GPU:
                traffic = pj.get("walker_dram_bytes_per_launch")
        line = {
            "metric": "verified subsequences/sec (cNSM-ED)", "value": value, "unit": "subsequences/s",
            "n_gpus": world, "steps": k, "warmup": args.warmup, "ms_per_step": dev_ms_max / k,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n_total, chunk, world),
            "e2e": {"value": e2e_value, "unit": "subsequences/s",
                    "h2d_bytes_per_step": int(h2d / k),
                    "d2h_bytes_per_step": int(12 * answers / k + 64),
                    "ms_per_step": 1e3 * wall_s_max / k,
                    "latency_ms_p50": 1e3 * float(np.median(lat)), "latency_ms_p95": 1e3 * float(np.percentile(lat, 95)),
                    "series_upload_s_once": t_load},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "cnsm_relay_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch, "kernel_ms_per_launch": walker_ms / k,
                         "samples_read_incl_chain_halos_per_launch": s_total / k,
                         "whole_step_frac": (bytes_per_launch / ((dev_ms / k) * 1e-3) / 1e9) / peak},
            "answers_per_step": answers_all / k, "gate_pass_per_step": gate_all / k,
            "wall_s_timed_region": t_wall_max, "datagen_s": t_gen,
            "best_of_last_query": best,
        }
        # Beside the headline (not part of it): the same 10 queries as ONE query-set call (kvm_verify_cnsm_ed_batch, an
        # addition to the reference's per-query engines: one statistics pass serves the set), host buffers in, answers out
        if world == 1 and not args.no_query_set:
            qs = np.stack(queries[:N_QUERIES])
            res = g.verify_cnsm_ed_batch(qs, EPSILON, ALPHA, BETA, iv)
            reps = 5
            t0 = time.perf_counter()
            for _ in range(reps):
                res = g.verify_cnsm_ed_batch(qs, EPSILON, ALPHA, BETA, iv)
            per_call = (time.perf_counter() - t0) / reps
            single = [g.verify_cnsm_ed(q, EPSILON, ALPHA, BETA, iv) for q in qs]
            line["query_set"] = {
                "queries_per_call": int(len(qs)), "ms_per_call": 1e3 * per_call,
                "value": float(sum(r.n_verified for r in res)) / per_call, "unit": "subsequences/s",
                "kernel_ms_per_call": float(sum(r.kernel_ms for r in res)),
                "statistics_pass_ms": float(res[0].stage_ms[0] * len(qs)),
                "identical_to_single_calls": bool(all(a.offsets.tolist() == b.offsets.tolist() and
                                                      a.distances.tolist() == b.distances.tolist()
                                                      for a, b in zip(res, single))),
                "note": "through the ABI with host buffers (compare with e2e, not with value)"}
        # Also beside the headline: the metric's other half, cNSM-DTW (BASELINE configs[3] shape on this GPU's shard:
        # m = 2048, rho = 5 % = 102, the headline's chain grid), two of the seeded offsets, kernel time from CUDA events
        if world == 1 and not args.no_dtw:
            m2, rho2, eps2 = 2048, 102, 1.0
            iv2 = datagen.chain_intervals(n_total, m2, chunk)
            rows = []
            for o2 in offs[:2]:
                q2 = query_of(n_total, o2, m2)
                g.verify_cnsm_dtw(q2, eps2, rho2, ALPHA, BETA, iv2)
                t0 = time.perf_counter()
                r2 = g.verify_cnsm_dtw(q2, eps2, rho2, ALPHA, BETA, iv2)
                rows.append({"offset": int(o2), "kernel_ms": r2.kernel_ms, "wall_ms": 1e3 * (time.perf_counter() - t0),
                             "verified": int(r2.n_verified), "gate_pass": int(r2.n_gate_pass), "dtws": int(r2.n_lb_pass),
                             "answers": int(r2.count), "stage_ms": [float(x) for x in r2.stage_ms[:3]]})
            line["cnsm_dtw"] = {"config": f"m={m2} rho={rho2} eps={eps2} alpha={ALPHA} beta={BETA}, n={n_total}, chain_chunk={chunk}",
                                "value": float(sum(r["verified"] for r in rows) / sum(r["kernel_ms"] * 1e-3 for r in rows)),
                                "unit": "subsequences/s", "queries": rows}
        # CPU baseline beside it (N=1 only): the oracle port on 1 core (the reference is single-threaded), the same
        # series, chains and queries; bounded to --cpu-queries whole-series queries (~0.7 s each)
        if world == 1:
            nq = max(1, min(args.cpu_queries, k))
            cpu_v = cpu_t = 0.0
            ok = True
            for j in range(nq):
                qi = (args.warmup + k - 1 - j) % N_QUERIES
                v, dt, outs = oracle_sample(local, queries[qi], iv, 1)
                cpu_v += v
                cpu_t += dt
                if j == 0:  # parity spot check on the last timed query (the oracle as checker, never as the thing measured)
                    ok = bool(last.offsets.tolist() == outs[0].offsets.tolist() and
                              last.distances.tolist() == outs[0].distances.tolist() and
                              last.n_gate_pass == outs[0].n_gate_pass)
            line["cpu_baseline"] = {"value": cpu_v / cpu_t, "unit": "subsequences/s", "cores": 1, "kind": "port",
                                    "sample": f"{nq} of the timed queries over the whole series ({len(iv)} chains), "
                                              f"{cpu_t:.1f} s on 1 core ({os.cpu_count()} cores on the box); series "
                                              f"resident in RAM (kinder than the reference, which re-reads its file)"}
            line["parity_vs_oracle_last_query"] = ok
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    g.close()


if __name__ == "__main__":
    main()
