"""CPU, world_size 2, gloo: the host-side logic of the N>1 path (shard bounds, global-id offsets,
all-gather layout, merge rule).  The per-shard search is the oracle here because there is no GPU;
on the GPU box tests/test_retrieval_gpu.py::test_sharded_equals_unsharded runs the same shards
through the CUDA scan and the CUDA merge kernel."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from persian_rag_system_b200.sharded import merge_topk_host_lists, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)                         # same data on every rank
    base = rng.standard_normal((301, 32)).astype(np.float32)
    x = np.concatenate([base, base[:50]])                    # duplicates straddle the shard boundary
    q = np.concatenate([base[:6], rng.standard_normal((5, 32)).astype(np.float32)])
    k = 9
    lo, hi = shard_bounds(x.shape[0], world, rank)
    D, I = O.flat_search_c(x[lo:hi], q, k, O.METRIC_L2, form=1)
    I = np.where(I >= 0, I + lo, -1)
    Dt, It = torch.from_numpy(D), torch.from_numpy(I)
    Dg = torch.empty((world * Dt.shape[0], Dt.shape[1]), dtype=Dt.dtype)      # rank-major concatenation
    Ig = torch.empty((world * It.shape[0], It.shape[1]), dtype=It.dtype)
    dist.all_gather_into_tensor(Dg, Dt)
    dist.all_gather_into_tensor(Ig, It)
    Dm, Im = merge_topk_host_lists(Dg.view(world, *Dt.shape).numpy(), Ig.view(world, *It.shape).numpy(), largest=False)
    Dw, Iw = O.flat_search_c(x, q, k, O.METRIC_L2, form=1)
    ok = bool(np.array_equal(Im, Iw) and np.array_equal(Dm, Dw))
    # every rank must hold the same merged answer
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([int(flag.item()), lo, hi]))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_merge_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(tmp_path / "ok_0.npy")
    r1 = np.load(tmp_path / "ok_1.npy")
    assert r0[0] == 1 and r1[0] == 1
    assert (r0[1], r0[2], r1[1], r1[2]) == (0, 176, 176, 351)
