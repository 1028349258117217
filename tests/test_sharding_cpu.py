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
