"""Multi-GPU check (torchrun, one rank per GPU): the row-sharded search -- fused peer-memory exchange
and the NCCL variant -- must equal the unsharded index bit for bit on every rank."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import persian_rag_system_b200 as P
from persian_rag_system_b200.sharded import ShardedFlatIndex, shard_bounds

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
bad = 0
cases = [(20000, 128, 33, 10, "fp16", P.METRIC_IP), (20000, 128, 33, 10, "fp16", P.METRIC_L2), (5000, 64, 7, 100, "fp16", P.METRIC_L2),
         (3000, 96, 5, 10, "fp32", P.METRIC_L2), (50000, 768, 128, 10, "bf16", P.METRIC_IP), (1000, 64, 300, 16, "fp16", P.METRIC_IP)]
for exchange in ("p2p", "nccl"):
    for (n, d, nq, k, storage, metric) in cases:
        rng = np.random.default_rng(n + d)                 # same data on every rank
        base = rng.standard_normal((n, d)).astype(np.float32)
        base[n // 2: n // 2 + 50] = base[:50]              # duplicates across shard boundaries -> ties on global id
        q = rng.standard_normal((nq, d)).astype(np.float32); q[:5] = base[:5]
        whole = P.FlatIndex(d, metric, storage, device=local); whole.add(base)
        qd = torch.from_numpy(q).to(dev)
        Dw, Iw = whole.search(qd, k)
        sh = ShardedFlatIndex(d, metric, storage, device=local, exchange=exchange, nq_cap=512, k_cap=128)
        lo, hi = shard_bounds(n, world, rank)
        sh.add_local(base[lo:hi], lo, n)
        for rep in range(3):                               # slot alternation
            D, I = sh.search(qd, k)
        sh.check_exchange()
        ok = bool(torch.equal(I, Iw)) and bool(torch.equal(D, Dw))
        flag = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if rank == 0:
            print(f"{exchange} n={n} d={d} nq={nq} k={k} {storage} metric={metric}: {'ok' if flag.item() else 'MISMATCH'}", flush=True)
        bad += 0 if flag.item() else 1
        del sh, whole
if rank == 0:
    print("TOTAL BAD", bad)
dist.destroy_process_group()
sys.exit(1 if bad else 0)
