"""Multi-GPU layout of the phase-2 path: one process per GPU, the series sharded by offset range.

Every candidate window start is verified independently, so the start offsets [1, n-m+1] are cut into
`world` contiguous ranges; rank r keeps the samples its starts need, i.e. its range plus a halo of
m-1 samples (precedent in the reference: the (w-1)-point overlap of the MapReduce index build,
K/mapreduce/BuildIndexMapReduce.java:216-221).  No series data crosses GPUs after load.  An interval
(= one running-statistics chain) is never split: it belongs to the rank whose range contains its first
scanned sample, and that rank's halo must reach its last one — otherwise the chain's rounding history,
and with it bit-exactness, would change.

The only exchange is the tail: per-rank counts, the sparse answers and the best match
(torch.distributed: NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    n: int            # global series length
    start_lo: int     # first window start (1-based) this rank owns
    start_hi: int     # last window start this rank owns
    first: int        # first sample held (1-based)
    count: int        # samples held, halo included

    @property
    def last(self) -> int:
        return self.first + self.count - 1


def make_shard(n: int, m_max: int, rank: int, world: int, halo: int | None = None, grid: int = 1) -> Shard:
    """Rank `rank` of `world`: owned starts are a 1/world slice of [1, n] rounded to multiples of `grid`
    (pass the chain chunk as `grid` so index-free chains never straddle shards); the samples held extend
    `halo` (default m_max-1) past the last owned start."""
    halo = m_max - 1 if halo is None else halo
    per = -(-n // world)
    per = -(-per // grid) * grid
    lo = rank * per + 1
    hi = min(n, (rank + 1) * per)
    if lo > n:
        return Shard(rank, world, n, lo, lo - 1, n, 1)
    last = min(n, hi + halo)
    return Shard(rank, world, n, lo, hi, lo, last - lo + 1)


def assign_intervals(intervals, shift: int, m: int, shard: Shard) -> np.ndarray:
    """The intervals this rank verifies: those whose first scanned sample max(left-shift, 1) lies in its
    owned range.  Raises if one of them needs samples beyond the rank's halo."""
    lr = np.asarray(intervals, dtype=np.int64).reshape(-1, 2)
    begin = np.maximum(lr[:, 0] - shift, 1)
    end = np.minimum(lr[:, 1] - shift + m - 1, shard.n)
    mine = (begin >= shard.start_lo) & (begin <= shard.start_hi)
    if np.any(end[mine] > shard.last):
        worst = int(end[mine].max())
        raise ValueError(f"rank {shard.rank}: an interval reaches sample {worst} but the shard ends at {shard.last}; "
                         f"load the shard with a larger halo (chains are never split across GPUs)")
    return lr[mine].astype(np.int32)


def merge_answers(offsets: np.ndarray, distances: np.ndarray, counters: dict, device=None):
    """All-gather the per-rank answers and counters.  Returns (offsets, distances, totals, best) on every
    rank: answers in ascending offset order (ranks own increasing ranges), totals = summed counters,
    best = (distance, offset) of the reference's `Best:` line (stable sort by distance: lowest offset wins
    ties, K/QueryEngine.java:373-376)."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        totals = dict(counters)
        best = None
        if len(offsets):
            i = int(np.lexsort((offsets, distances))[0])
            best = (float(distances[i]), int(offsets[i]))
        return offsets, distances, totals, best
    world = dist.get_world_size()
    dev = device if device is not None else torch.device("cpu")
    keys = sorted(counters)
    head = torch.tensor([len(offsets)] + [int(counters[k]) for k in keys], dtype=torch.int64, device=dev)
    heads = [torch.zeros_like(head) for _ in range(world)]
    dist.all_gather(heads, head)                      # (1) counts
    counts = [int(h[0]) for h in heads]
    totals = {k: int(sum(int(h[1 + i]) for h in heads)) for i, k in enumerate(keys)}
    cap = max(max(counts), 1)
    pad_o = torch.zeros(cap, dtype=torch.int32, device=dev)
    pad_d = torch.full((cap,), float("inf"), dtype=torch.float64, device=dev)
    if len(offsets):
        pad_o[:len(offsets)] = torch.from_numpy(np.ascontiguousarray(offsets)).to(dev)
        pad_d[:len(offsets)] = torch.from_numpy(np.ascontiguousarray(distances)).to(dev)
    all_o = [torch.zeros_like(pad_o) for _ in range(world)]
    all_d = [torch.zeros_like(pad_d) for _ in range(world)]
    dist.all_gather(all_o, pad_o)                     # (2) sparse answers, padded to the largest count
    dist.all_gather(all_d, pad_d)
    offs = np.concatenate([all_o[r][:counts[r]].cpu().numpy() for r in range(world)])
    dists = np.concatenate([all_d[r][:counts[r]].cpu().numpy() for r in range(world)])
    # (3) best match: min over distance, then the lowest offset among the minimisers (NCCL has no arg-min)
    local_min = torch.tensor([float(distances.min()) if len(distances) else float("inf")], dtype=torch.float64,
                             device=dev)
    dist.all_reduce(local_min, op=dist.ReduceOp.MIN)
    best = None
    if np.isfinite(local_min.item()):
        mine = offsets[distances == local_min.item()]
        cand = torch.tensor([int(mine.min()) if len(mine) else np.iinfo(np.int64).max], dtype=torch.int64, device=dev)
        dist.all_reduce(cand, op=dist.ReduceOp.MIN)
        best = (float(local_min.item()), int(cand.item()))
    return offs, dists, totals, best
