"""Generates tests/golden/*.json: answers of the CPU oracle (oracle/kvm_oracle.cpp) on seeded inputs.

The reference ships no golden vectors and cannot run here (Java, no JVM), so these fixtures pin the ORACLE
(regression) and give the GPU tests vectors that need neither the oracle nor /root/reference at run time.
Distances are stored as IEEE-754 hex strings (bit-exact).  Run from the repo root: python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from kvmatch_b200 import datagen  # noqa: E402
from oracle import kvm_oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = [
    # name, kind, n, seed, query offset, m, params
    ("ed_n200k_m256", "ed", 200_000, 101, 150_001, 256, dict(eps=12.0, shift=25, chunk=None)),
    ("cnsm_ed_n300k_m512", "cnsm_ed", 300_000, 102, 77_777, 512, dict(eps=14.0, alpha=1.5, beta=5.0, chunk=20_000)),
    ("cnsm_ed_n300k_m1023", "cnsm_ed", 300_000, 103, 200_002, 1023, dict(eps=22.0, alpha=2.0, beta=20.0, chunk=8_191)),
    ("dtw_n100k_m128", "dtw", 100_000, 104, 50_000, 128, dict(eps=10.0, rho=6, chunk=None)),
    ("cnsm_dtw_n150k_m512", "cnsm_dtw", 150_000, 105, 120_000, 512, dict(eps=5.0, rho=25, alpha=1.5, beta=5.0, chunk=30_000)),
    ("runs_n250k", "runs", 250_000, 106, 0, 0, dict(ws=[25, 50, 100, 200, 400])),
]


def intervals_for(n, m, chunk, shift=0):
    if chunk is None:  # a phase-1-like pruned list: a few ragged intervals
        rng = np.random.default_rng(n + m)
        lefts = np.sort(rng.choice(np.arange(1 + shift, n - m - 400), size=40, replace=False))
        out, end = [], 0
        for l in lefts:
            l = max(int(l), end + 2)
            r = l + int(rng.integers(0, 300))
            out.append((l, r))
            end = r
        return out
    return datagen.chain_intervals(n, m, chunk).tolist()


def run_case(name, kind, n, seed, off, m, p):
    s = datagen.generate(n, seed)
    doc = {"name": name, "kind": kind, "n": n, "seed": seed, "query_offset": off, "m": m, "params": p}
    if kind == "runs":
        doc["runs"] = {}
        for w in p["ws"]:
            keys, first, last = kvm_oracle.window_mean_runs(s, w)
            doc["runs"][str(w)] = {"count": int(len(keys)), "first_head": first[:50].tolist(), "last_head": last[:50].tolist(),
                                   "keys_head": [float(k).hex() for k in keys[:50]],
                                   "checksum_first": int(first.astype(np.int64).sum()), "checksum_last": int(last.astype(np.int64).sum()),
                                   "checksum_keys": int(np.bitwise_xor.reduce(keys.view(np.int64)))}
        return doc
    q = s[off - 1:off - 1 + m].copy()
    if kind in ("ed", "dtw"):
        q = q + np.random.default_rng(seed).normal(scale=0.02, size=m)  # not an exact copy
    iv = intervals_for(n, m, p.get("chunk"), p.get("shift", 0))
    if kind == "ed":
        iv = sorted(iv + [(off + p["shift"] - 4, off + p["shift"] + 4)])
        r = kvm_oracle.verify_ed(s, q, p["eps"], iv, p["shift"])
    elif kind == "cnsm_ed":
        r = kvm_oracle.verify_cnsm_ed(s, q, p["eps"], p["alpha"], p["beta"], iv)
    elif kind == "dtw":
        iv = sorted(iv + [(off - 50, off + 50)])
        r = kvm_oracle.verify_dtw(s, q, p["eps"], p["rho"], iv)
    else:
        r = kvm_oracle.verify_cnsm_dtw(s, q, p["eps"], p["rho"], p["alpha"], p["beta"], iv)
    doc["query_noise_seed"] = seed if kind in ("ed", "dtw") else None
    doc["intervals"] = [list(map(int, x)) for x in iv]
    doc["offsets"] = r.offsets.tolist()
    doc["distances_hex"] = [float(d).hex() for d in r.distances]
    doc["cnt_candidate"], doc["n_verified"], doc["s_total"], doc["n_gate_pass"] = r.cnt_candidate, r.n_verified, r.s_total, r.n_gate_pass
    return doc


if __name__ == "__main__":
    for case in CASES:
        doc = run_case(*case)
        with open(os.path.join(HERE, case[0] + ".json"), "w") as f:
            json.dump(doc, f)
        print(case[0], "answers:", len(doc.get("offsets", [])), "intervals:", len(doc.get("intervals", [])))
