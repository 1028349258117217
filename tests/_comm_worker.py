"""Worker of tests/test_gpu_comm.py: one rank of a torchrun job (one GPU per rank).  Every rank verifies its offset shard
through the C ABI, the library merges over NCCL (kvm_gather_result), rank 0 compares with the oracle on the whole series."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch.distributed as dist
import kvmatch_b200
from kvmatch_b200 import datagen, sharding

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
n, m, chunk = 600_000, 512, 4096
s = datagen.generate(n, seed=77)
sh = sharding.make_shard(n, m, rank, world, grid=chunk)
g = kvmatch_b200.GpuSeries(int(os.environ.get("LOCAL_RANK", rank)))
g.load(s[sh.first - 1:sh.last], n=n, first=sh.first)
g.comm_init()
iv_all = datagen.chain_intervals(n, m, chunk)
iv = sharding.assign_intervals(iv_all, 0, m, sh)
ok = True
for off, eps in ((100_000, 4.0), (450_000, 4.0), (450_000, 60.0)):   # the last one: > 256 answers on a rank (second round)
    q = s[off - 1:off - 1 + m].copy()
    local = g.verify_cnsm_ed(q, eps, 1.5, 5.0, iv)
    merged, best = g.gather(local)
    if rank == 0:
        from oracle import kvm_oracle
        e = kvm_oracle.verify_cnsm_ed(s, q, eps, 1.5, 5.0, iv_all)
        i = int(np.lexsort((e.offsets, e.distances))[0])
        good = (merged.offsets.tolist() == e.offsets.tolist() and merged.distances.tolist() == e.distances.tolist()
                and merged.n_verified == e.n_verified and merged.n_gate_pass == e.n_gate_pass
                and best == (float(e.distances[i]), int(e.offsets[i])))
        print(f"query {off} eps {eps}: {merged.count} answers, best {best}, exchange {merged.stage_ms[0]:.3f} ms: {'OK' if good else 'MISMATCH'}", flush=True)
        ok = ok and good
    # RSM-DTW through the same tail
loc = g.verify_dtw(s[99_999:99_999 + m].copy(), 9.0, 25, iv)
mer, best = g.gather(loc)
if rank == 0:
    from oracle import kvm_oracle
    e = kvm_oracle.verify_dtw(s, s[99_999:99_999 + m].copy(), 9.0, 25, iv_all)
    good = mer.offsets.tolist() == e.offsets.tolist() and mer.distances.tolist() == e.distances.tolist()
    print(f"RSM-DTW: {mer.count} answers: {'OK' if good else 'MISMATCH'}", flush=True)
    ok = ok and good
g.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
