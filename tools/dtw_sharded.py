"""BASELINE config 4, sharded: cNSM-DTW over ONE series of n samples (default 1e9) split by offset range across the
ranks of a torchrun job (halo m-1, chains never split), answers merged over NCCL.  Strong scaling: total work fixed.
usage: torchrun --nproc-per-node N tools/dtw_sharded.py [n] [eps]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import kvmatch_b200, bench
from kvmatch_b200 import datagen, sharding

n_total = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
eps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
m, rho, alpha, beta = 2048, 102, 1.5, 5.0
rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
chunk = 100_000 - m + 1                                      # the reference's epoch grid
shard = sharding.make_shard(n_total, m, rank, world, grid=chunk)
local = datagen.generate_range(n_total, shard.first - 1, shard.last, bench.SEED)
g = kvmatch_b200.GpuSeries(local_rank)
g.load(local, n=n_total, first=shard.first)
all_iv = datagen.chain_intervals(n_total, m, chunk, lo=shard.start_lo, hi=min(shard.start_hi, n_total - m + 1))
iv = sharding.assign_intervals(all_iv, 0, m, shard)
offs = [int(x) for x in np.random.default_rng(bench.SEED).integers(1, n_total - m, 3)]
rows = []
for off in offs:
    q = datagen.generate_range(n_total, off - 1, off - 1 + m, bench.SEED)
    g.verify_cnsm_dtw(q, eps, rho, alpha, beta, iv)          # warm-up (buffer growth)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    r = g.verify_cnsm_dtw(q, eps, rho, alpha, beta, iv)
    o_m, d_m, totals, best = sharding.merge_answers(r.offsets, r.distances, {"n_verified": r.n_verified, "dtws": r.n_lb_pass}, device=dev)
    wall = time.perf_counter() - t
    stats = torch.tensor([r.kernel_ms, wall * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    if rank == 0:
        k_ms, w_ms = [float(x) for x in stats.tolist()]
        rows.append({"offset": off, "kernel_ms_max_over_ranks": k_ms, "wall_ms_incl_merge": w_ms, "answers": int(len(o_m)),
                     "verified": totals["n_verified"], "dtws": totals["dtws"], "best": best,
                     "subseq_per_s": totals["n_verified"] / (k_ms * 1e-3)})
if rank == 0:
    print(json.dumps({"config": f"cNSM-DTW n={n_total:.0e} m={m} rho={rho} eps={eps} alpha={alpha} beta={beta}", "n_gpus": world, "queries": rows}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
