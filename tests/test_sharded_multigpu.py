"""GPU, >= 2 devices: the row-sharded search on REAL peer GPUs (one process per GPU, spawned here) --
`merge_xchg_kernel` (fused local merge + NVLink peer-memory exchange + global merge) and the NCCL
all-gather variant -- must equal the unsharded index bit for bit on every rank, including ties on
global ids across shard boundaries, 16-bit L2 (direct-form re-rank), wide k and slot alternation.
Skipped on a single-GPU box (there `test_retrieval_gpu.py::test_sharded_equals_unsharded` runs emulated
shards through the non-fused merge kernel)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


CASES = [  # n, d, nq, k, storage, metric (0 IP, 1 L2)
    (20000, 128, 33, 10, "fp16", 0), (20000, 128, 33, 10, "fp16", 1), (5000, 64, 7, 100, "fp16", 1),
    (3000, 96, 5, 10, "fp32", 1), (50000, 768, 128, 10, "bf16", 0), (1000, 64, 300, 16, "fp16", 0),
    (40000, 384, 64, 5, "bf16", 1),
]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import persian_rag_system_b200 as P
    from persian_rag_system_b200.sharded import ShardedFlatIndex, shard_bounds
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    bad = []
    for exchange in ("p2p", "nccl"):
        for (n, d, nq, k, storage, metric) in CASES:
            rng = np.random.default_rng(n + d)                 # same data on every rank
            base = rng.standard_normal((n, d)).astype(np.float32)
            base[n // 2: n // 2 + 50] = base[:50]              # duplicates across shard boundaries -> ties on global id
            q = rng.standard_normal((nq, d)).astype(np.float32)
            q[:5] = base[:5]
            whole = P.FlatIndex(d, metric, storage, device=rank)
            whole.add(base)
            qd = torch.from_numpy(q).to(dev)
            Dw, Iw = whole.search(qd, k)
            sh = ShardedFlatIndex(d, metric, storage, device=rank, exchange=exchange, nq_cap=512, k_cap=128)
            lo, hi = shard_bounds(n, world, rank)
            sh.add_local(base[lo:hi], lo, n)
            ok = True
            for rep in range(4):                               # slot alternation / generation counter
                if rep == 2:
                    sh.local.set_fused(2)                      # push in the scan kernel's tail + pull kernel (nq <= 128, k <= 16)
                D, I = sh.search(qd, k)
                ok = ok and bool(torch.equal(I, Iw)) and bool(torch.equal(D, Dw))
            sh.check_exchange()
            flag = torch.tensor([1 if ok else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if not flag.item():
                bad.append((exchange, n, d, nq, k, storage, metric))
            del sh, whole
    # container: every rank writes its block in parallel, a fresh sharded index loads it back (f-2)
    n, d = 20000, 128
    base = np.random.default_rng(n + d).standard_normal((n, d)).astype(np.float32)
    sh = ShardedFlatIndex(d, 1, "fp16", device=rank, exchange="p2p", nq_cap=64, k_cap=16)
    lo, hi = shard_bounds(n, world, rank)
    sh.add_local(base[lo:hi], lo, n)
    qd = torch.from_numpy(base[:33]).to(dev)
    D0, I0 = sh.search(qd, 10)
    cdir = os.path.join(out_dir, "container")
    sh.write(cdir)
    sh2 = ShardedFlatIndex(d, 1, "fp16", device=rank, exchange="p2p", nq_cap=64, k_cap=16)
    sh2.load(cdir)
    D1, I1 = sh2.search(qd, 10)
    ok = sh2.ntotal == n and bool(torch.equal(I1, I0)) and bool(torch.equal(D1, D0))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if not flag.item():
        bad.append(("container",))
    del sh, sh2
    # sparse path, doc-range shards on real GPUs: NCCL all-gather of the [nq, k] float64 lists + device merge
    from persian_rag_system_b200.sharded import ShardedSparseIndex
    from persian_rag_system_b200.sparse import build_bm25_csr
    rng = np.random.default_rng(5)
    vocab = [f"w{i}" for i in range(300)]
    docs = [[vocab[j] for j in rng.integers(0, 300, size=int(rng.integers(1, 30)))] for _ in range(4000)]
    docs = docs + docs[:500]
    b = build_bm25_csr(docs)
    qs = [docs[i][:5] for i in range(20)] + [["nope"]]
    helper = P.BM25Index(docs[:8], device=rank)
    helper.vocab = b["vocab"]
    ip, qt, qw = helper.encode_queries(qs)
    dip, dqt, dqw = (torch.from_numpy(v).to(dev) for v in (ip, qt, qw))
    sparse_bad = 0
    for mode in ("exact", "throughput"):
        whole = P.SparseIndex(b["indptr"], b["indices"], b["weights"], len(b["vocab"]), device=rank, mode=mode)
        Sw, Iw = whole.search_device(dip, dqt, dqw, 10)
        shs = ShardedSparseIndex.from_global_csr(b["indptr"], b["indices"], b["weights"], len(b["vocab"]), device=rank, mode=mode)
        S, I = shs.search_device(dip, dqt, dqw, 10)
        ok = bool(torch.equal(S, Sw)) and bool(torch.equal(I, Iw))
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        sparse_bad += 0 if flag.item() else 1
        del shs, whole
    if sparse_bad:
        bad.append(("sparse", sparse_bad))
    # a peer that never searches: the fused exchange gives up after the timeout, answers -1 and reports
    sh = ShardedFlatIndex(32, 1, "fp16", device=rank, exchange="p2p", nq_cap=8, k_cap=16, lanes=1)
    x = np.random.default_rng(1).standard_normal((256, 32)).astype(np.float32)
    lo, hi = shard_bounds(256, world, rank)
    sh.add_local(x[lo:hi], lo, 256)
    sh.set_exchange_timeout(0.2)
    timeout_ok = True
    if rank == 0:
        D, I = sh.search(torch.from_numpy(x[:4]).to(dev), 5)
        torch.cuda.synchronize()
        timeout_ok = bool((I == -1).all().item())
        try:
            sh.check_exchange()
            timeout_ok = False                                  # must raise
        except P.PrsError:
            pass
    dist.barrier()
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([len(bad), int(timeout_ok)]))
    if bad:
        print(f"rank {rank}: MISMATCH {bad}", flush=True)
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs on one box (gpurun --gpus 2)")
@pytest.mark.timeout(900)
def test_sharded_search_on_real_gpus_equals_unsharded(tmp_path):
    import torch.multiprocessing as mp
    world = min(_ngpu(), 8)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        res = np.load(tmp_path / f"r{r}.npy")
        assert res[0] == 0, f"rank {r}: {res[0]} sharded cases differ from the unsharded index"
        assert res[1] == 1, f"rank {r}: timeout handling"


@pytest.mark.skipif(_ngpu() < 2, reason="needs >= 2 GPUs on one box (gpurun --gpus 2)")
@pytest.mark.timeout(600)
def test_single_process_multi_device_index_equals_single_device(tmp_path):
    """SURVEY 8b `devices=[...]`: ONE process (like the reference's RetrievalSystem / Gradio app) splits the rows
    over the GPUs of the box -- worker thread per device, peer access, the fused merge + exchange kernel -- and
    returns exactly what the single-device index returns."""
    import torch
    import persian_rag_system_b200 as P
    from oracle import oracle as O
    devs = list(range(min(_ngpu(), 8)))
    for (n, d, nq, k, storage, metric) in CASES + [(125, 384, 9, 5, "fp32", 1), (125, 384, 9, 5, "fp16", 1), (3, 32, 2, 5, "fp32", 0)]:
        rng = np.random.default_rng(n + d)
        base = rng.standard_normal((n, d)).astype(np.float32)
        if n > 200:
            base[n // 2: n // 2 + 50] = base[:50]              # ties on global ids across the device blocks
        q = rng.standard_normal((nq, d)).astype(np.float32)
        q[:2] = base[:2]
        whole = P.FlatIndex(d, metric, storage, device=0)
        whole.add(base)
        grp = P.FlatIndex(d, metric, storage, devices=devs, nq_cap=128, k_cap=128)
        grp.reserve(n)
        a, b = n // 3, 2 * n // 3
        grp.add(base[:a])                                       # host rows
        grp.add(torch.from_numpy(base[a:b]).to(f"cuda:{devs[-1]}"))   # rows produced on another device of the group
        grp.add(base[b:])
        assert grp.ntotal == n and sum(grp.shard_rows) == n and grp.d == d and grp.storage == storage
        Dw, Iw = whole.search(q, k)
        D, I = grp.search(q, k)                                 # nq > nq_cap is chunked
        assert np.array_equal(I, Iw) and np.array_equal(D, Dw), (n, d, nq, k, storage, metric)
        Dt, It = grp.search(torch.from_numpy(q).to(f"cuda:{devs[0]}"), k)
        assert np.array_equal(It.cpu().numpy(), Iw) and np.array_equal(Dt.cpu().numpy(), Dw)
        assert np.array_equal(grp.reconstruct_n(0, n), whole.reconstruct_n(0, n))
    # the drop-in surface: read_index / RetrievalSystem over several devices, byte-exact write-back
    gold = os.path.join(ROOT, "tests", "golden", "indices", "drugs_sentence_chunks.index")
    idx = P.read_index(gold, devices=devs)
    x, _ = O.read_faiss_flat(gold)
    D, I = idx.search(x[:7], 1)
    assert I[:, 0].tolist() == list(range(7)) and (D[:, 0] == 0).all()
    P.write_index(idx, str(tmp_path / "back.index"))
    assert open(tmp_path / "back.index", "rb").read() == open(gold, "rb").read()
