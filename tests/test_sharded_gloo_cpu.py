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


def _sparse_worker(rank, world, port, out_dir):
    """Doc-range sharding of the sparse path on gloo: global BM25 weights sliced by rows, local
    top-k with global ids (the per-shard scoring is the oracle here: no GPU), all-gather, merge with ties on
    global ids DESCENDING == the unsharded canonical order."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from persian_rag_system_b200.sharded import merge_topk_host_lists, shard_bounds
    from persian_rag_system_b200.sparse import build_bm25_csr
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(9)
    vocab = [f"w{i}" for i in range(60)]
    base = [[vocab[j] for j in rng.integers(0, 60, size=int(rng.integers(1, 15)))] for _ in range(150)]
    docs = base + base[:40]                                   # duplicates straddle the shard boundary
    qs = [base[3][:3], base[100][:5], ["zzz"]]
    b = build_bm25_csr(docs)
    n, k = len(docs), 7
    lo, hi = shard_bounds(n, world, rank)
    import scipy.sparse as sp
    M = sp.csr_matrix((b["weights"], b["indices"], b["indptr"]), shape=(n, len(b["vocab"])))[lo:hi].tocsc()
    S = np.zeros((len(qs), k))
    I = np.full((len(qs), k), -1, np.int64)
    full = O.BM25OkapiOracle(docs)
    for r, q in enumerate(qs):
        sc = np.zeros(hi - lo)
        for tok in q:
            t = b["vocab"].get(tok, -1)
            if t >= 0:
                sc += M[:, t].toarray().ravel()
        top = O.argsort_topk_canonical(sc, k)
        S[r, :len(top)], I[r, :len(top)] = sc[top], top + lo
    St, It = torch.from_numpy(S), torch.from_numpy(I)
    Sg = torch.empty((world * St.shape[0], k), dtype=St.dtype)
    Ig = torch.empty((world * It.shape[0], k), dtype=It.dtype)
    dist.all_gather_into_tensor(Sg, St)
    dist.all_gather_into_tensor(Ig, It)
    Sm, Im = merge_topk_host_lists(Sg.view(world, *St.shape).numpy(), Ig.view(world, *It.shape).numpy(), largest=True, tie_high_id=True)
    ok = True
    for r, q in enumerate(qs):
        sc = full.get_scores(q)
        want = O.argsort_topk_canonical(sc, k)
        ok = ok and Im[r].tolist() == want.tolist() and np.array_equal(Sm[r], sc[want])
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    np.save(os.path.join(out_dir, f"sp_{rank}.npy"), np.array([int(flag.item())]))
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_sparse_merge_world2_gloo(tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_sparse_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert np.load(tmp_path / "sp_0.npy")[0] == 1 and np.load(tmp_path / "sp_1.npy")[0] == 1
