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


class PackedMerger:
    """The multi-GPU tail as ONE fixed-size collective per query: every rank packs
    [#answers, best distance, best offset, counters..., first `cap` offsets, first `cap` distances] into a float64
    buffer (offsets are int32: exact in float64), one all_gather moves it, every rank unpacks.  Only a query with more
    than `cap` answers on some rank pays a second, padded all_gather for the overflow.  Buffers are allocated once.

    merge() returns (offsets, distances, totals, best) like merge_answers(); `last_device_ms` is the device time of
    the exchange (CUDA events on torch's current stream; 0 on CPU / single process)."""

    def __init__(self, counter_keys, device=None, cap: int = 512):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.keys = list(counter_keys)
        self.cap = int(cap)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.dev = device if device is not None else torch.device("cpu")
        self.cuda = self.dev.type == "cuda"
        self.head = 3 + len(self.keys)
        self.len = self.head + 2 * self.cap
        self.host = torch.zeros(self.len, dtype=torch.float64, pin_memory=self.cuda)
        self.host_np = self.host.numpy()
        self.all_host = torch.zeros(self.world * self.len, dtype=torch.float64, pin_memory=self.cuda)
        if self.world > 1:
            self.send = torch.zeros(self.len, dtype=torch.float64, device=self.dev)
            self.recv = torch.zeros(self.world * self.len, dtype=torch.float64, device=self.dev)
            if self.cuda:
                self.ev0, self.ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.last_device_ms = 0.0
        self.overflows = 0

    def merge(self, offsets: np.ndarray, distances: np.ndarray, counters: dict):
        n = len(offsets)
        best_d, best_o = np.inf, float(np.iinfo(np.int64).max)
        if n:
            i = int(np.lexsort((offsets, distances))[0])
            best_d, best_o = float(distances[i]), float(offsets[i])
        if self.world == 1:
            self.last_device_ms = 0.0
            best = (best_d, int(best_o)) if n else None
            return offsets, distances, dict(counters), best
        torch, dist = self.torch, self.dist
        h = self.host_np
        h[0], h[1], h[2] = n, best_d, best_o
        for j, k in enumerate(self.keys):
            h[3 + j] = counters[k]
        c = min(n, self.cap)
        h[self.head:self.head + c] = offsets[:c]
        h[self.head + self.cap:self.head + self.cap + c] = distances[:c]
        if self.cuda:
            self.ev0.record()
        self.send.copy_(self.host, non_blocking=True)
        dist.all_gather_into_tensor(self.recv, self.send)
        self.all_host.copy_(self.recv, non_blocking=True)
        if self.cuda:
            self.ev1.record()
            self.ev1.synchronize()
            self.last_device_ms = float(self.ev0.elapsed_time(self.ev1))
        a = self.all_host.numpy().reshape(self.world, self.len)
        counts = a[:, 0].astype(np.int64)
        totals = {k: int(a[:, 3 + j].sum()) for j, k in enumerate(self.keys)}
        offs = [a[r, self.head:self.head + min(counts[r], self.cap)].astype(np.int32) for r in range(self.world)]
        dists = [a[r, self.head + self.cap:self.head + self.cap + min(counts[r], self.cap)].copy() for r in range(self.world)]
        over = int(counts.max()) - self.cap
        if over > 0:  # rare: some rank holds more answers than the packed buffer carries
            self.overflows += 1
            pad = torch.zeros(2 * over, dtype=torch.float64)
            k = max(0, n - self.cap)
            pad[:k] = torch.from_numpy(np.ascontiguousarray(offsets[self.cap:], dtype=np.float64))
            pad[over:over + k] = torch.from_numpy(np.ascontiguousarray(distances[self.cap:]))
            pad = pad.to(self.dev)
            got = torch.zeros(self.world * 2 * over, dtype=torch.float64, device=self.dev)
            dist.all_gather_into_tensor(got, pad)
            g = got.cpu().numpy().reshape(self.world, 2 * over)
            for r in range(self.world):
                k = max(0, int(counts[r]) - self.cap)
                offs[r] = np.concatenate([offs[r], g[r, :k].astype(np.int32)])
                dists[r] = np.concatenate([dists[r], g[r, over:over + k]])
        best = None
        if counts.sum() > 0:
            r = int(np.lexsort((a[:, 2], a[:, 1]))[0])  # min distance, then the lowest offset (stable sort of the reference)
            best = (float(a[r, 1]), int(a[r, 2]))
        return np.concatenate(offs), np.concatenate(dists), totals, best
