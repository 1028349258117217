#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native KV-match phase-2 path.

A "step" is one cNSM-ED verification query (BASELINE.json: NormQueryEngine, query length 1024, alpha=1.5, beta=5,
epsilon=5) over EVERY window start of ONE DataGenerator-style synthetic series of n = 1e9 samples (the size
BASELINE.json's metric quotes its latency on; 8 GB, fits one B200), index-free: the merged-interval list handed to
the C ABI is [1, n-m+1] cut into statistic chains of --chunk candidates (<= 100000-m+1, the reference's own epoch
chunking, K/experiments/ucr/UcrDtwQueryExecutor.java:97).

  python bench.py --gpus N --steps K --warmup W              # ours (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU loops (oracle port), rank 0 only

N GPUs: STRONG scaling — the same series sharded by offset range (rank r holds its 1/N of the window starts plus an
m-1 halo; chains are never split); every step ends with the multi-GPU tail inside the timed region: ONE packed
all_gather (count, counters, best match, the first 256 answers of every rank; kvm_gather_result in the library).
`value` = window starts verified by all ranks / max-over-ranks device time of the K timed steps (the library's CUDA
events per kernel stage + CUDA events around the exchange; series resident in HBM).  The exchange is the library's own
(kvm_comm_init / kvm_gather_result: a peer-memory exchange kernel on the ctx's stream, NCCL for the overflow round);
torch.distributed carries the NCCL id and the IPC handles, the
barriers and the final statistics.  `e2e` = the same through the C
ABI with host buffers (query + interval list in, answers out, exchange included), from the barrier-bracketed wall
clock of the K steps.  At N=1 the line also carries BASELINE.json configs[1] itself (n = 1e8) as `cfg2_n1e8`.
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
N_TOTAL = 1_000_000_000  # one series for every N (strong scaling); KVM_BENCH_N overrides (developer runs)
N_CFG2 = 100_000_000     # BASELINE.json configs[1]
N_QUERIES = 10           # seeded query offsets, cycled over the steps
SEED = datagen.DEFAULT_SEED
DEFAULT_CHUNK = 2048     # candidates per statistic chain (see DESIGN.md "chain chunking")


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
    n = args.n
    cores = os.cpu_count() or 1
    chunk = min(args.chunk, 100000 - M + 1)
    # bounded sample: the first `sample_n` samples of the same series, same chains, same queries
    sample_n = int(min(args.ref_sample, n))
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
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
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
    return {"workload": "cNSM-ED NormQueryEngine phase-2 verification, index-free scan of every window start of one "
                        "series of n samples (BASELINE.json metric: n = 1e9; configs[1] = the same at n = 1e8, reported "
                        "as cfg2_n1e8 at N=1)",
            "n_total": n_total, "query_length": M, "alpha": ALPHA, "beta": BETA, "epsilon": EPSILON,
            "chain_chunk": chunk, "queries": N_QUERIES, "parallelism": f"offset-shard x{world} (strong scaling)",
            "tail": "one packed all_gather per query, inside the timed step",
            "l2": "inputs (8 GB / N per GPU) exceed the 126 MB L2; no explicit flush between steps"}


def unique_samples_of(iv, n_total):
    """every touched series sample counted once (the m-1 halo between adjacent chains is not
    algorithmic)."""
    ivs = np.asarray(iv, dtype=np.int64).reshape(-1, 2)
    lo_s, hi_s = ivs[:, 0], np.minimum(ivs[:, 1] + M - 1, n_total)
    order = np.argsort(lo_s, kind="stable")
    lo_s, hi_s = lo_s[order], hi_s[order]
    reach = np.maximum.accumulate(hi_s)
    prev_reach = np.concatenate(([lo_s[0] - 1], reach[:-1]))
    return int(np.sum(np.maximum(0, hi_s - np.maximum(lo_s - 1, prev_reach))))


def offline_traffic(n, chunk):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture of this configuration (a constant
    measured offline, labelled as such), or None."""
    prof = os.path.join(ROOT, "profiles", "roofline_r02.json")
    if os.path.exists(prof):
        pj = json.load(open(prof))
        for row in pj.get("captures", []):
            if row.get("n") == n and row.get("chain_chunk") == chunk:
                return row.get("dram_bytes_per_launch"), row.get("source")
    return None, None


def timed_queries(g, queries, iv, steps, warmup, merger=None, barrier=None):
    """warmup untimed + `steps` timed cNSM-ED queries through the C ABI (host buffers in, answers out), each followed
    by the multi-GPU exchange when a merger is given.  Returns per-step lists and the barrier-bracketed wall time."""
    def one(i):
        t = time.perf_counter()
        r = g.verify_cnsm_ed(queries[i % len(queries)], EPSILON, ALPHA, BETA, iv)
        merged = None
        tail_ms = 0.0
        if merger == "library":   # the library's own NCCL tail: one packed ncclAllGather on its stream (kvm_gather_result)
            mr, best = g.gather(r)
            merged = (mr.offsets, mr.distances, {"n_verified": mr.n_verified, "gate": mr.n_gate_pass}, best)
            tail_ms = mr.stage_ms[0]
        elif merger is not None:
            merged = merger.merge(r.offsets, r.distances, {"n_verified": r.n_verified, "gate": r.n_gate_pass})
            tail_ms = merger.last_device_ms
        return r, merged, tail_ms, time.perf_counter() - t
    for i in range(warmup):
        one(i)
    if barrier:
        barrier()
    t0 = time.perf_counter()
    rows = [one(warmup + i) for i in range(steps)]
    if barrier:
        barrier()
    return rows, time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=float, default=float(os.environ.get("KVM_BENCH_N", N_TOTAL)))
    ap.add_argument("--no-side", action="store_true", help="skip the side measurements (cfg2_n1e8, cNSM-DTW, window means)")
    ap.add_argument("--chunk", type=int, default=int(os.environ.get("KVM_BENCH_CHUNK", DEFAULT_CHUNK)))
    ap.add_argument("--ref-sample", type=float, default=1e8)   # samples scanned per reference step
    ap.add_argument("--cpu-queries", type=int, default=10)     # queries the 1-core CPU baseline times
    args = ap.parse_args()
    args.n = int(args.n)
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

    n_total = args.n
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
    # multi-GPU tail: the library's communicator (KVM_BENCH_TAIL=torch selects the torch.distributed PackedMerger instead)
    if world > 1 and os.environ.get("KVM_BENCH_TAIL", "library") == "library":
        g.comm_init(rank, world)
        merger = "library"
    else:
        merger = sharding.PackedMerger(["n_verified", "gate"], device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    # the clock samples of the warm-up are dropped at the first barrier: only the timed region counts
    first_barrier = [True]

    def barrier_hook():
        barrier()
        if first_barrier[0]:
            sampler.lines.clear()
            first_barrier[0] = False
    rows, t_wall = timed_queries(g, queries, iv, args.steps, args.warmup, merger, barrier=barrier_hook)
    clocks = sampler.stop()

    k = args.steps
    dev_ms = sum(r.kernel_ms + tail for r, _, tail, _ in rows)
    kern_ms = sum(r.kernel_ms for r, _, _, _ in rows)
    stream_ms = sum(r.stage_ms[0] for r, _, _, _ in rows)
    tail_ms = sum(tail for _, _, tail, _ in rows)
    lat = [dt for _, _, _, dt in rows]
    verified = sum(r.n_verified for r, _, _, _ in rows)
    launches = sum(r.n_launches for r, _, _, _ in rows)
    answers = sum(r.count for r, _, _, _ in rows)
    gate = sum(r.n_gate_pass for r, _, _, _ in rows)
    h2d = sum(r.h2d_bytes for r, _, _, _ in rows)
    rewalked = sum(r.n_rewalked for r, _, _, _ in rows)
    last, last_merged = rows[-1][0], rows[-1][1]

    tail_min = torch.tensor([tail_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tail_min, op=dist.ReduceOp.MIN)
    stats = torch.tensor([dev_ms, t_wall, kern_ms, stream_ms, tail_ms], dtype=torch.float64, device=dev)
    sums = torch.tensor([verified, launches, answers, gate, rewalked], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    dev_ms_max, t_wall_max, kern_ms_max, stream_ms_max, tail_ms_max = [float(x) for x in stats.tolist()]
    verified_all, launches_all, answers_all, gate_all, rewalked_all = [float(x) for x in sums.tolist()]

    if rank == 0:
        peak, peak_src = load_peaks()
        value = verified_all / (dev_ms_max * 1e-3)
        e2e_value = verified_all / t_wall_max
        # roofline of the dominant kernel (the streaming statistics pass): algorithmic bytes = 8 B per touched sample
        # (bytes_alg = 8*S_total + 12*#answers), rank 0's shard, per launch, over its own CUDA-event time
        unique_samples = unique_samples_of(iv, n_total)
        bytes_per_launch = 8.0 * unique_samples + 12.0 * answers / k
        stream_s = (stream_ms / k) * 1e-3
        achieved = bytes_per_launch / stream_s / 1e9
        traffic, traffic_src = offline_traffic(int(n_total // world), chunk)
        line = {
            "metric": "verified subsequences/sec (cNSM-ED)", "value": value, "unit": "subsequences/s",
            "n_gpus": world, "steps": k, "warmup": args.warmup, "ms_per_step": dev_ms_max / k,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n_total, chunk, world),
            "e2e": {"value": e2e_value, "unit": "subsequences/s",
                    "h2d_bytes_per_step": int(h2d / k),
                    "d2h_bytes_per_step": int(12 * answers / k + 64),
                    "ms_per_step": 1e3 * t_wall_max / k,
                    "latency_ms_p50": 1e3 * float(np.median(lat)), "latency_ms_p95": 1e3 * float(np.percentile(lat, 95)),
                    "series_upload_s_once": t_load},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "cnsm_stream_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "traffic_offline_ncu": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_per_launch, "kernel_ms_per_launch": stream_ms / k,
                         "whole_step_frac": (bytes_per_launch / ((kern_ms / k) * 1e-3) / 1e9) / peak,
                         "stage_ms_per_step": {"stream": stream_ms / k,
                                               "rewalk": sum(r.stage_ms[1] for r, _, _, _ in rows) / k,
                                               "exact": sum(r.stage_ms[2] for r, _, _, _ in rows) / k}},
            "tail": {"ms_per_step_device_max_over_ranks": tail_ms_max / k,
                     # the rank that arrives last at the collective waits for nobody: its time is the exchange itself, the
                     # difference to the maximum is load imbalance between the shards of one query
                     "ms_per_step_device_min_over_ranks": float(tail_min.item()) / k,
                     "collectives_per_step": 1 if world > 1 else 0,
                     "implementation": ("kvm_gather_result: one peer-memory exchange kernel on the library's stream (CUDA IPC over NVLink / NVSwitch; "
                                        "ncclAllGather only for a rank with more than 256 answers), timed by CUDA events"
                                        if os.environ.get("KVM_GATHER_P2P", "1") != "0" and world <= 8 else
                                        "kvm_gather_result: one packed ncclAllGather on the library's stream, timed by CUDA events")
                     if merger == "library" else "torch.distributed all_gather_into_tensor (sharding.PackedMerger)",
                     "packed_bytes_per_rank": 8 * (16 + 2 * 256) if merger == "library" else 8 * merger.len},
            "parity": "oracle-only (reference unpinned: Java 8, no JVM in this image)",
            "answers_per_step": answers_all / k, "gate_pass_per_step": gate_all / k, "rewalked_windows_per_step": rewalked_all / k,
            "wall_s_timed_region": t_wall_max, "datagen_s": t_gen,
            "best_of_last_query": last_merged[3] if last_merged else None,
        }
        if world == 1 and not args.no_side:
            side_measurements(line, g, kvmatch_b200, local, n_total, chunk, queries, offs, iv, last, args, peak)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    g.close()


def side_measurements(line, g, kvmatch_b200, local, n_total, chunk, queries, offs, iv, last, args, peak):
    """N=1 only, beside the headline: BASELINE configs[1] at n = 1e8, the CPU baseline with a parity spot check, the
    metric's other half (cNSM-DTW), the query set and the fused window-mean pass."""
    k = args.steps
    # ---- CPU baseline: the oracle port on 1 core (the reference is single-threaded), first 1e8 samples of the same
    # series and chains; the oracle also checks the last timed query on that prefix
    sample_n = int(min(1e8, n_total))
    iv_s = datagen.chain_intervals(sample_n, M, chunk)
    prefix = local[:sample_n]
    nq = max(1, min(args.cpu_queries, N_QUERIES))
    cpu_v = cpu_t = 0.0
    ok = True
    for j in range(nq):
        qi = (args.warmup + k - 1 - j) % N_QUERIES
        v, dt, outs = oracle_sample(prefix, queries[qi], iv_s, 1)
        cpu_v += v
        cpu_t += dt
        if j == 0:  # parity spot check (the oracle as checker, never as the thing measured)
            hi = sample_n - M + 1
            keep = last.offsets <= hi
            ok = bool(last.offsets[keep].tolist() == outs[0].offsets.tolist() and
                      last.distances[keep].tolist() == outs[0].distances.tolist())
    line["cpu_baseline"] = {"value": cpu_v / cpu_t, "unit": "subsequences/s", "cores": 1, "kind": "port",
                            "sample": f"{nq} of the timed queries over the first {sample_n} samples ({len(iv_s)} chains), "
                                      f"{cpu_t:.1f} s on 1 core ({os.cpu_count()} cores on the box); series resident in "
                                      f"RAM (kinder than the reference, which re-reads its file)"}
    line["parity_vs_oracle_last_query"] = ok
    # ---- BASELINE configs[1]: the same engine on an n = 1e8 series (the series BENCH_r01 used)
    s8 = datagen.generate(N_CFG2, SEED)
    g8 = kvmatch_b200.GpuSeries(0)
    g8.load(s8)
    iv8 = datagen.chain_intervals(N_CFG2, M, chunk)
    q8 = [query_of(N_CFG2, o, M) for o in query_offsets(N_CFG2, M, N_QUERIES)]
    rows8, wall8 = timed_queries(g8, q8, iv8, max(20, k), 3)
    k8 = len(rows8)
    kern8 = sum(r.kernel_ms for r, _, _, _ in rows8)
    stream8 = sum(r.stage_ms[0] for r, _, _, _ in rows8)
    b8 = 8.0 * unique_samples_of(iv8, N_CFG2) + 12.0 * sum(r.count for r, _, _, _ in rows8) / k8
    tr8, tr8_src = offline_traffic(N_CFG2, chunk)
    exp8 = oracle_sample(s8, q8[(3 + k8 - 1) % N_QUERIES], iv8, os.cpu_count() or 1)[2]
    got8 = rows8[-1][0]
    exp_off = np.concatenate([o.offsets for o in exp8])
    exp_dist = np.concatenate([o.distances for o in exp8])
    line["cfg2_n1e8"] = {
        "value": sum(r.n_verified for r, _, _, _ in rows8) / (kern8 * 1e-3), "unit": "subsequences/s",
        "ms_per_step": kern8 / k8, "e2e_value": sum(r.n_verified for r, _, _, _ in rows8) / wall8,
        "e2e_ms_per_step": 1e3 * wall8 / k8, "steps": k8,
        "roofline": {"bound": "hbm", "kernel": "cnsm_stream_kernel", "achieved": b8 / (stream8 / k8 * 1e-3) / 1e9,
                     "peak": peak, "unit": "GB/s", "frac": b8 / (stream8 / k8 * 1e-3) / 1e9 / peak,
                     "whole_step_frac": b8 / (kern8 / k8 * 1e-3) / 1e9 / peak, "traffic_offline_ncu": tr8,
                     "traffic_source": tr8_src, "kernel_ms_per_launch": stream8 / k8},
        "parity_vs_oracle_whole_series": bool(got8.offsets.tolist() == exp_off.tolist() and
                                              got8.distances.tolist() == exp_dist.tolist() and
                                              got8.n_gate_pass == sum(o.n_gate_pass for o in exp8))}
    # ---- the same 10 queries as ONE query-set call (an addition to the reference's per-query engines; n = 1e8)
    try:
        qs = np.stack(q8[:N_QUERIES])
        res = g8.verify_cnsm_ed_batch(qs, EPSILON, ALPHA, BETA, iv8)
        t0 = time.perf_counter()
        res = g8.verify_cnsm_ed_batch(qs, EPSILON, ALPHA, BETA, iv8)
        per_call = time.perf_counter() - t0
        line["query_set"] = {"n": N_CFG2, "queries_per_call": int(len(qs)), "ms_per_call": 1e3 * per_call,
                             "value": float(sum(r.n_verified for r in res)) / per_call, "unit": "subsequences/s",
                             "note": "relay walker, through the ABI with host buffers (compare with cfg2_n1e8.e2e_value)"}
    except Exception as e:
        line["query_set"] = {"error": repr(e)}
    g8.close()
    del s8
    # ---- the metric's other half: cNSM-DTW, BASELINE configs[3] (n = 1e9, m = 2048, rho = 5 % = 102), the seeded
    # queries at eps in {1, 5, 10} within a time budget; DTW roofline = 5 flops per executed band cell / FP64 peak
    try:
        smp = ClockSampler(0)   # the heavy queries run seconds of FP64 / integer work: the clocks they ran at belong beside them
        smp.start()
        line["cnsm_dtw"] = dtw_block(g, n_total, chunk, offs)
        line["cnsm_dtw"]["clocks"] = smp.stop()
    except Exception as e:
        line["cnsm_dtw"] = {"error": repr(e)}
    # ---- IndexBuilder's window-mean pass, all five windows of Sigma
    try:
        line["window_mean"] = window_mean_block(kvmatch_b200, peak)
    except Exception as e:  # keep the headline even if a side block fails
        line["window_mean"] = {"error": repr(e)}


FP64_PEAK_FLOPS = 36.5e12  # measured on this pool's B200 with tools/fp64_peak.cu (18.27e12 DFMA/s), DESIGN.md


def dtw_block(g, n_total, chunk, offs, budget_s=25.0):
    m2, rho2 = 2048, 102
    iv2 = datagen.chain_intervals(n_total, m2, chunk)
    cells_full = m2 * (2 * rho2 + 1) - rho2 * (rho2 + 1)
    out = {"config": f"m={m2} rho={rho2} alpha={ALPHA} beta={BETA}, n={n_total}, chain_chunk={chunk}", "eps": {}}
    t_start = time.perf_counter()
    for eps in (1.0, 5.0, 10.0):
        rows = []
        for o2 in offs:
            if time.perf_counter() - t_start > budget_s and rows:
                break
            q2 = query_of(n_total, o2, m2)
            t0 = time.perf_counter()
            r2 = g.verify_cnsm_dtw(q2, eps, rho2, ALPHA, BETA, iv2)
            rows.append({"offset": int(o2), "kernel_ms": r2.kernel_ms, "wall_ms": 1e3 * (time.perf_counter() - t0),
                         "verified": int(r2.n_verified), "gate_pass": int(r2.n_gate_pass), "dtws": int(r2.n_lb_pass),
                         "cells": int(r2.n_dtw_cells), "answers": int(r2.count), "band_dtws": int(r2.n_exact),
                         "stage_ms": [float(x) for x in r2.stage_ms]})
        kt = sum(r["kernel_ms"] for r in rows) * 1e-3
        dtw_t = sum(r["stage_ms"][3] for r in rows) * 1e-3
        cells = sum(r["cells"] for r in rows)
        out["eps"][str(eps)] = {
            "queries": len(rows), "value": sum(r["verified"] for r in rows) / kt if kt else None, "unit": "subsequences/s",
            "kernel_ms_p50": float(np.median([r["kernel_ms"] for r in rows])),
            "dtws": sum(r["dtws"] for r in rows), "cells_executed": cells,
            "roofline": {"bound": "fp64", "kernel": "dtw_band_kernel", "unit": "TFLOP/s", "peak": FP64_PEAK_FLOPS / 1e12,
                         "achieved": 5.0 * cells / dtw_t / 1e12 if dtw_t and cells else None,
                         "frac": 5.0 * cells / dtw_t / FP64_PEAK_FLOPS if dtw_t and cells else None,
                         "cells_if_no_abandon": sum(r["dtws"] for r in rows) * cells_full},
            "rows": rows}
        if time.perf_counter() - t_start > budget_s:
            break
    return out


def window_mean_block(kvmatch_b200, peak):
    out = {}
    for n in (1_000_000, N_CFG2):
        s = datagen.generate(n, SEED)
        g = kvmatch_b200.GpuSeries(0)
        g.load(s)
        if not hasattr(g, "window_mean_runs_all"):
            g.close()
            return {"error": "fused pass not built"}
        g.window_mean_runs_all(kvmatch_b200.WU_LIST)
        res = g.window_mean_runs_all(kvmatch_b200.WU_LIST)
        ms = res.kernel_ms
        out[f"n={n}"] = {"windows": list(kvmatch_b200.WU_LIST), "kernel_ms_all_windows": ms, "runs": int(res.n_runs),
                        "rewalked_chains": int(res.n_chains_rewalked),
                        "roofline": {"bound": "hbm", "kernel": "window_mean_stream_kernel", "unit": "GB/s", "peak": peak,
                                     "achieved": (8.0 * n + 16.0 * res.n_runs) / (ms * 1e-3) / 1e9,
                                     "frac": (8.0 * n + 16.0 * res.n_runs) / (ms * 1e-3) / 1e9 / peak}}
        if n == N_CFG2:  # BASELINE configs[1] (ii): the whole index-pruned query on this series
            try:
                out["index_pruned_n1e8"] = index_pruned_block(kvmatch_b200, g, s, n)
            except Exception as e:
                out["index_pruned_n1e8"] = {"error": repr(e)}
        g.close()
    return out


def index_pruned_block(kvmatch_b200, g, s, n, n_queries=3):
    """The reference's whole query() for cNSM-ED: the five indexes built from the fused window-mean pass (+ host step 2 and
    file images), phases 0 / 1 on the host (kvmatch_b200/phase1.py over kvm_norm_intervals_*), phase 2 on the GPU over
    the phase-1 interval list; beside it the same query as a full scan, whose answer offsets it must reproduce."""
    from kvmatch_b200 import phase1
    t0 = time.perf_counter()
    images = kvmatch_b200.IndexBuilder(g).build_all()
    build_s = time.perf_counter() - t0
    indexes = [phase1.IndexFile(images[w]) for w in phase1.WU_LIST]
    full_iv = datagen.chain_intervals(n, M, DEFAULT_CHUNK)
    rows = []
    for off in query_offsets(n, M, N_QUERIES)[:n_queries]:
        q = query_of(n, off, M)
        t0 = time.perf_counter()
        valid, last_segment, plan = phase1.phase1_norm(q, EPSILON, ALPHA, BETA, n, indexes)
        t1_ms = 1e3 * (time.perf_counter() - t0)
        iv = np.asarray(valid, dtype=np.int32).reshape(-1, 2)
        shift = (last_segment - 1) * 25
        g.verify_cnsm_ed(q, EPSILON, ALPHA, BETA, iv, shift)
        t0 = time.perf_counter()
        r = g.verify_cnsm_ed(q, EPSILON, ALPHA, BETA, iv, shift)
        t2_ms = 1e3 * (time.perf_counter() - t0)
        g.verify_cnsm_ed(q, EPSILON, ALPHA, BETA, full_iv)
        f = g.verify_cnsm_ed(q, EPSILON, ALPHA, BETA, full_iv)
        lens = iv[:, 1] - iv[:, 0] + 1
        rows.append({"query_offset": int(off), "segments": len(plan), "last_segment": int(last_segment), "phase1_host_ms": t1_ms,
                     "intervals": int(len(iv)), "candidates": int(lens.sum()), "longest_interval": int(lens.max()),
                     "phase2_kernel_ms": r.kernel_ms, "phase2_wall_ms": t2_ms, "full_scan_kernel_ms": f.kernel_ms,
                     "answers": int(r.count), "same_answer_offsets_as_full_scan": bool(r.offsets.tolist() == f.offsets.tolist())})
    return {"index_build_s": build_s, "index_bytes": int(sum(len(b) for b in images.values())), "queries": rows,
            "note": "phase 1 on the host (plan on one core, segments probed on up to four threads); incremental index visiting and wall-clock early termination off (DESIGN 1, row f1)"}


if __name__ == "__main__":
    main()
