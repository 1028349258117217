"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes, each verifying its shard with the oracle
standing in for the device, then merging with the same collectives the GPU path uses."""
import os
import socket

import numpy as np
import pytest

from kvmatch_b200 import datagen, sharding


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_make_shard_partitions_all_starts():
    n, m = 1_000_003, 1024
    for world in (1, 2, 4, 8):
        shards = [sharding.make_shard(n, m, r, world, grid=4096) for r in range(world)]
        assert shards[0].start_lo == 1 and shards[-1].start_hi == n
        for a, b in zip(shards, shards[1:]):
            assert b.start_lo == a.start_hi + 1
            assert a.last == min(n, a.start_hi + m - 1)  # halo of m-1 samples
            assert (a.start_hi) % 4096 == 0              # shard edges sit on the chain grid


def test_assign_intervals_never_splits_a_chain():
    n, m = 100_000, 128
    iv = datagen.chain_intervals(n, m, 4096)
    shards = [sharding.make_shard(n, m, r, 2, grid=4096) for r in range(2)]
    parts = [sharding.assign_intervals(iv, 0, m, s) for s in shards]
    assert np.concatenate(parts).tolist() == iv.tolist()
    with pytest.raises(ValueError):  # a chain longer than the halo cannot be verified bit-exactly on one shard
        sharding.assign_intervals([(shards[0].start_hi - 10, shards[0].start_hi + 5000)], 0, m, shards[0])


def _worker(rank, world, port, n, m, out_dir):
    import torch.distributed as dist

    from oracle import kvm_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chunk = 4096
    shard = sharding.make_shard(n, m, rank, world, grid=chunk)
    local = datagen.generate_range(n, shard.first - 1, shard.last)      # this rank's samples only
    full_q_src = datagen.generate_range(n, 30_000 - 1, 30_000 - 1 + m)  # the query, same on every rank
    iv = sharding.assign_intervals(datagen.chain_intervals(n, m, chunk), 0, m, shard)
    # the oracle works on a whole series: embed the shard at its global position
    series = np.zeros(shard.last)
    series[shard.first - 1:] = local
    res = kvm_oracle.verify_cnsm_ed(series, full_q_src, 6.0, 1.5, 5.0, iv)
    offs, dists, totals, best = sharding.merge_answers(res.offsets, res.distances,
                                                       {"n_verified": res.n_verified, "gate": res.n_gate_pass})
    if rank == 0:
        np.savez(os.path.join(out_dir, "merged.npz"), offs=offs, dists=dists, n_verified=totals["n_verified"],
                 gate=totals["gate"], best=np.array(best))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_merge_equals_single_process(tmp_path, oracle):
    import torch.multiprocessing as mp
    n, m = 120_000, 256
    mp.spawn(_worker, args=(2, free_port(), n, m, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "merged.npz")
    s = datagen.generate(n)
    q = s[30_000 - 1:30_000 - 1 + m].copy()
    exp = oracle.verify_cnsm_ed(s, q, 6.0, 1.5, 5.0, datagen.chain_intervals(n, m, 4096))
    assert got["offs"].tolist() == exp.offsets.tolist()
    assert got["dists"].tolist() == exp.distances.tolist()
    assert int(got["n_verified"]) == exp.n_verified and int(got["gate"]) == exp.n_gate_pass
    i = int(np.lexsort((exp.offsets, exp.distances))[0])
    assert got["best"].tolist() == [exp.distances[i], exp.offsets[i]] and exp.offsets[i] == 30_000


def _packed_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    merger = sharding.PackedMerger(["n_verified", "gate"], cap=8)
    rng = np.random.default_rng(100 + rank)
    results = []
    for step, n_ans in enumerate([0, 3, 8, 9, 40 if rank == 1 else 2, 0 if rank == 0 else 5]):
        offs = np.sort(rng.choice(np.arange(1, 10_000), size=n_ans, replace=False)).astype(np.int32) + 10_000 * rank
        dists = rng.random(n_ans)
        if step == 2 and n_ans:
            dists[:] = 0.125  # ties across ranks: the lowest offset must win
        o, d, totals, best = merger.merge(offs, dists, {"n_verified": 1000 + rank, "gate": step})
        results.append((o, d, totals, best, offs, dists))
    np.save(os.path.join(out_dir, f"rank{rank}.npy"), np.array(results, dtype=object), allow_pickle=True)
    dist.barrier()
    dist.destroy_process_group()


def test_packed_merge_two_ranks_incl_overflow(tmp_path):
    """PackedMerger: one fixed-size all_gather per query, a second one only when a rank holds more than `cap`
    answers; both ranks must see the concatenation in rank order, the summed counters and the reference's best."""
    import torch.multiprocessing as mp
    mp.spawn(_packed_worker, args=(2, free_port(), str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "rank0.npy", allow_pickle=True)
    r1 = np.load(tmp_path / "rank1.npy", allow_pickle=True)
    for step in range(len(r0)):
        o0, d0, t0, b0, lo0, ld0 = r0[step]
        o1, d1, t1, b1, lo1, ld1 = r1[step]
        exp_o, exp_d = np.concatenate([lo0, lo1]), np.concatenate([ld0, ld1])
        assert o0.tolist() == exp_o.tolist() == o1.tolist()
        assert d0.tolist() == exp_d.tolist() == d1.tolist()
        assert t0 == t1 == {"n_verified": 2001, "gate": 2 * step}
        if len(exp_o):
            i = int(np.lexsort((exp_o, exp_d))[0])
            assert b0 == b1 == (float(exp_d[i]), int(exp_o[i]))
        else:
            assert b0 is None and b1 is None
