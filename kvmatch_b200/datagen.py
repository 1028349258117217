"""Seeded DataGenerator-style synthetic series (measurement / test inputs).

Restates the *distribution* of the reference's generator, which is unseeded
(K/DataGenerator.java:80-119, K/data/RandomWalkGenerator.java:40-50,
K/data/GaussianGenerator.java:43-84, K/data/SineGenerator.java:45-56,
K/utils/RandomUtils.java:34-47: random(min,max) is U[min, max+1e-5)):

  repeat until n points: pick one of 3 generators uniformly; segment length
  l ~ U{min(1000, L) .. L} with L = min(left, n/100);
    random walk : start U[-5,5], step U[0,1] with a random sign, running sum
    gaussian    : mean U[-5,5], std U[0,2], i.i.d. normal
    noisy sine  : freq U[2,10], amp U[2,10], mean U[-5,5], phase U[0,2pi],
                  mean + amp*sin(2*i*(pi/l)*freq + phase) + U[-0.05 amp, 0.05 amp]

PRNG: numpy PCG64 seeded with SeedSequence([seed, k]); k = 0 plans the segments, k = 1 + s
generates segment s — so any offset range can be generated independently (multi-GPU shards).
Series are FP64, 0-based numpy arrays; sample k (1-based reference offset) is series[k-1].
"""
from __future__ import annotations

import numpy as np

DEFAULT_SEED = 20260117
_EPS = 0.00001


def _uniform(rng, lo, hi, size=None):
    return rng.uniform(lo, hi + _EPS, size)


def plan_segments(n: int, seed: int = DEFAULT_SEED):
    """[(start, length, generator_id)] covering [0, n)."""
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence([seed, 0])))
    segs = []
    start = 0
    while start < n:
        left = n - start
        L = max(1, min(left, n // 100))
        t = int(rng.integers(0, 3))
        l = int(rng.integers(min(1000, L), L + 1))
        segs.append((start, l, t))
        start += l
    return segs


def _segment(seed: int, index: int, length: int, kind: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(np.random.SeedSequence([seed, 1 + index])))
    if kind == 0:  # random walk
        out = np.empty(length)
        out[0] = _uniform(rng, -5, 5)
        if length > 1:
            sign = np.where(rng.random(length - 1) < 0.5, -1.0, 1.0)
            out[1:] = sign * _uniform(rng, 0, 1, length - 1)
        return np.cumsum(out)  # sequential running sum, as ts[i] = ts[i-1] + sign*step
    if kind == 1:  # gaussian
        mean = _uniform(rng, -5, 5)
        std = _uniform(rng, 0, 2)
        return mean + std * rng.standard_normal(length)
    freq = _uniform(rng, 2, 10)
    amp = _uniform(rng, 2, 10)
    mean = _uniform(rng, -5, 5)
    phase = _uniform(rng, 0, 2 * np.pi)
    i = np.arange(length, dtype=np.float64)
    noise = rng.uniform(amp * 0.05 * -1, amp * 0.05 + _EPS, length)
    return mean + amp * np.sin(2 * i * (np.pi / length) * freq + phase) + noise


def generate_range(n: int, lo: int, hi: int, seed: int = DEFAULT_SEED) -> np.ndarray:
    """Samples [lo, hi) (0-based) of the length-n series with this seed."""
    lo = max(0, lo)
    hi = min(n, hi)
    out = np.empty(max(0, hi - lo))
    for index, (start, length, kind) in enumerate(plan_segments(n, seed)):
        if start + length <= lo or start >= hi:
            continue
        seg = _segment(seed, index, length, kind)
        a = max(lo, start)
        b = min(hi, start + length)
        out[a - lo:b - lo] = seg[a - start:b - start]
    return out


def generate(n: int, seed: int = DEFAULT_SEED) -> np.ndarray:
    return generate_range(n, 0, n, seed)


def chain_intervals(n: int, m: int, chunk: int, lo: int = 1, hi: int | None = None) -> np.ndarray:
    """Index-free candidate list: window starts [lo, hi] (1-based, default the whole series) cut into
    chains of at most `chunk` candidates.  chunk = 100000 - m + 1 is the epoch chunking of
    K/experiments/ucr/UcrDtwQueryExecutor.java:97,169-214.  Returns int32 [K, 2] (left, right)."""
    hi = n - m + 1 if hi is None else hi
    if hi < lo:
        return np.zeros((0, 2), dtype=np.int32)
    left = np.arange(lo, hi + 1, chunk, dtype=np.int64)
    right = np.minimum(left + chunk - 1, hi)
    return np.stack([left, right], axis=1).astype(np.int32)
