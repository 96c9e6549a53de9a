"""GPU: the drop-in RetrievalSystem end to end (same calls a user of src/retrieval.py makes),
Hit@K / MRR equality with the oracle pipeline, hybrid fusion, and the sharded merge."""
import os

import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def P():
    import persian_rag_system_b200 as P
    assert P.lib().prs_device_arch(0) == 100, P._lib.last_error()
    return P


class FakeEncoder:
    """Stands in for SentenceTransformer (absent here): deterministic text -> vector."""

    def __init__(self, table, d):
        self.table, self.d = table, d

    def encode(self, sentences, device=None, **_):
        out = np.zeros((len(sentences), self.d), np.float32)
        for i, s in enumerate(sentences):
            out[i] = self.table[s]
        return out


@pytest.fixture(scope="module")
def world(gold_dir, golden_indices, golden_texts, tmp_path_factory):
    """A 125-chunk corpus in the reference's CSV schema (src/chunking.py:45-53) whose row i is
    `word_chunk_i`, the reference's shipped MiniLM-ft index as its dense index, and queries that are
    seeded perturbations of corpus rows (relevance label = the perturbed row)."""
    import pandas as pd
    chunks_known, _ = golden_texts
    f = "paraphrase-multilingual-MiniLM-L12-v2_finetuned_drugs_word_chunks.index"
    x, _ = golden_indices[f]
    n = x.shape[0]
    rng = np.random.default_rng(77)
    vocab = sorted({w for c in chunks_known for w in c["text"].split()})
    texts = [" ".join(rng.choice(vocab, size=int(rng.integers(40, 150)))) for _ in range(n)]
    rows = [{"id": f"word_chunk_{i}", "text": texts[i], "start_word": i * 125, "end_word": i * 125 + 150,
             "num_words": len(texts[i].split()), "chunk_type": "word_based", "overlap_words": 25} for i in range(n)]
    tmp = tmp_path_factory.mktemp("world")
    csv = str(tmp / "drugs_word_chunks.csv")
    pd.DataFrame(rows).to_csv(csv, index=False, encoding="utf-8")
    g = np.load(os.path.join(gold_dir, "flat_golden.npz"))
    q = g["q_0"]
    queries, table, relevant = [], {}, {}
    for i in range(q.shape[0]):
        words = texts[i % n].split()
        text = " ".join(words[3:9])
        text = f"{text} #{i}"
        queries.append({"id": f"q{i}", "question": text})
        table[text] = q[i]
        relevant[f"q{i}"] = [f"word_chunk_{i % n}"] if i % 3 else []          # some queries have no labels
    return dict(csv=csv, index=os.path.join(gold_dir, "indices", f), x=x, q=q, texts=texts, queries=queries,
                table=table, relevant=relevant, rows=rows)


def test_dense_retriever_matches_reference_semantics(P, world):
    r = P.RetrievalSystem(method="dense", encoder=FakeEncoder(world["table"], 384))
    assert r.retrieve("x") == []                                     # not ready yet (src/retrieval.py:224-226)
    assert r.load_chunks_and_index(world["csv"], world["index"]) is True
    assert r.faiss_index.ntotal == 125 and r.is_ready
    Dr, Ir = O.flat_search_c(world["x"], world["q"], 5, O.METRIC_L2, form=1)
    for i, qd in enumerate(world["queries"][:20]):
        res = r.retrieve(qd["question"], top_k=5)
        assert [c["id"] for c, _ in res] == [f"word_chunk_{j}" for j in Ir[i]]
        for (c, s), dist in zip(res, Dr[i]):
            assert s == pytest.approx(1 / (1 + dist), rel=1e-5)      # src/retrieval.py:108
        ctx, meta = r.get_contexts_for_rag(qd["question"], top_k=5, max_context_length=2000)
        want_ctx, want_meta = O.pack_contexts(res, 2000)
        assert ctx == want_ctx and [m["chunk_id"] for m in meta] == [m["chunk_id"] for m in want_meta]
        assert sum(len(c) for c in ctx) <= 2000 + 3


def test_missing_index_is_skipped_like_the_reference(P, world):
    r = P.RetrievalSystem(method="dense", encoder=FakeEncoder(world["table"], 384))
    assert r.load_chunks_and_index(world["csv"], "/no/such.index") is True      # src/retrieval.py:52 silently skips
    assert r.retrieve(world["queries"][0]["question"]) == []                      # :94-95
    assert P.RetrievalSystem(method="bm25").load_chunks_and_index("/no/such.csv") is False


def test_dimension_mismatch_returns_empty_list(P, world):
    """results/phase4_rag_evaluation_results.json:999-1060: wrong-d queries -> exception -> []."""
    r = P.RetrievalSystem(method="dense", encoder=FakeEncoder({k: np.zeros(512, np.float32) for k in world["table"]}, 512))
    assert r.load_chunks_and_index(world["csv"], world["index"])
    assert r.retrieve(world["queries"][0]["question"], 5) == []


def test_hit_at_k_and_mrr_identical_to_oracle_pipeline(P, world):
    r = P.RetrievalSystem(method="dense", encoder=FakeEncoder(world["table"], 384))
    assert r.load_chunks_and_index(world["csv"], world["index"])
    got = r.evaluate_retrieval_quality(world["queries"], world["relevant"])
    _, Ir = O.flat_search_c(world["x"], world["q"], 10, O.METRIC_L2, form=1)
    ids = {qd["id"]: [f"word_chunk_{j}" for j in Ir[i]] for i, qd in enumerate(world["queries"])}
    want = O.retrieval_quality(ids, world["queries"], world["relevant"])
    assert got == want and 0 < got["hit_at_1"] <= got["hit_at_5"] <= 1 and got["mrr"] > 0


@pytest.mark.parametrize("method", ["bm25", "tfidf"])
def test_sparse_retrievers_match_oracle(P, world, method):
    r = P.RetrievalSystem(method=method)
    assert r.load_chunks_and_index(world["csv"])
    texts = world["texts"]
    if method == "bm25":
        bm = O.BM25OkapiOracle([t.split() for t in texts])
        score_fn = lambda q: bm.get_scores(q.split())
    else:
        vec, mat = O.tfidf_fit(texts)
        score_fn = lambda q: O.tfidf_scores(vec, mat, q)
    for qd in world["queries"][:16]:
        sc = score_fn(qd["question"])
        res = r.retrieve(qd["question"], top_k=10)
        want = O.argsort_topk_canonical(sc, 10)
        assert [c["id"] for c, _ in res] == [f"word_chunk_{j}" for j in want]
        assert [float(s) for _, s in res] == [float(v) for v in sc[want]]
    ids = {qd["id"]: [c["id"] for c, _ in r.retrieve(qd["question"], 10)] for qd in world["queries"]}
    assert r.evaluate_retrieval_quality(world["queries"], world["relevant"]) == O.retrieval_quality(ids, world["queries"], world["relevant"])


def test_hybrid_matches_reference_fusion(P, world):
    r = P.RetrievalSystem(method="hybrid", encoder=FakeEncoder(world["table"], 384))
    assert r.load_chunks_and_index(world["csv"], world["index"])
    bm = O.BM25OkapiOracle([t.split() for t in world["texts"]])
    chunks = r.chunks
    for i, qd in enumerate(world["queries"][:12]):
        Dr, Ir = O.flat_search_c(world["x"], world["q"][i:i + 1], 10, O.METRIC_L2, form=1)
        dense = [(chunks[j], 1 / (1 + d)) for d, j in zip(Dr[0], Ir[0])]
        sc = bm.get_scores(qd["question"].split())
        sparse = [(chunks[j], sc[j]) for j in O.argsort_topk_canonical(sc, 10)]
        want = O.hybrid_fuse(dense, sparse, 5)
        got = r.retrieve(qd["question"], top_k=5)
        assert [c["id"] for c, _ in got] == [c["id"] for c, _ in want]
        np.testing.assert_allclose([s for _, s in got], [s for _, s in want], rtol=1e-6)


def test_hybrid_batch_on_device_equals_per_query_and_oracle(P, world):
    """f-1: top-2k dense and top-2k BM25 lists stay in HBM, one kernel fuses the whole batch."""
    import torch
    r = P.RetrievalSystem(method="hybrid", encoder=FakeEncoder(world["table"], 384))
    assert r.load_chunks_and_index(world["csv"], world["index"])
    qs = [qd["question"] for qd in world["queries"]]
    batch = r.retrieve_batch(qs, top_k=5)
    assert len(batch) == len(qs)
    bm = O.BM25OkapiOracle([t.split() for t in world["texts"]])
    chunks = r.chunks
    for i in range(0, len(qs), 5):
        Dr, Ir = O.flat_search_c(world["x"], world["q"][i:i + 1], 10, O.METRIC_L2, form=1)
        dense = [(chunks[j], 1 / (1 + d)) for d, j in zip(Dr[0], Ir[0])]
        sc = bm.get_scores(qs[i].split())
        sparse = [(chunks[j], sc[j]) for j in O.argsort_topk_canonical(sc, 10)]
        want = O.hybrid_fuse(dense, sparse, 5)
        assert [c["id"] for c, _ in batch[i]] == [c["id"] for c, _ in want]
        np.testing.assert_allclose([s for _, s in batch[i]], [s for _, s in want], rtol=1e-6)
        assert [c["id"] for c, _ in r.retrieve(qs[i], top_k=5)] == [c["id"] for c, _ in batch[i]]
    # raw kernel: rows outside [0, n_chunks) are dropped before the maxima, stable order on ties, -1 padding
    D = torch.tensor([[0.0, 1.0, 3.0, 7.0]], dtype=torch.float32, device="cuda")
    Id = torch.tensor([[3, 99, 1, -1]], dtype=torch.int64, device="cuda")
    S = torch.tensor([[4.0, 2.0, 0.0, 0.0]], dtype=torch.float64, device="cuda")
    Is = torch.tensor([[1, 2, 0, 77]], dtype=torch.int64, device="cuda")
    F, I = P.hybrid_fuse(D, Id, S, Is, n_chunks=10, top_k=6)
    ch = [{"id": f"c{i}"} for i in range(10)]
    want = O.hybrid_fuse([(ch[3], 1.0), (ch[1], 0.25)], [(ch[1], 4.0), (ch[2], 2.0), (ch[0], 0.0)], 6)
    assert I[0].tolist() == [int(c["id"][1:]) for c, _ in want] + [-1, -1]
    np.testing.assert_allclose(F[0, :4].cpu().numpy(), [s for _, s in want], rtol=1e-12)
    # one retriever missing: the other list alone (reference: guarded sub-call returns [])
    F, I = P.hybrid_fuse(D[:, :0], Id[:, :0], S, Is, n_chunks=10, top_k=3)
    assert I[0].tolist() == [1, 2, 0] and F[0].tolist() == [0.4, 0.2, 0.0]
    nodense = P.RetrievalSystem(method="hybrid", encoder=FakeEncoder(world["table"], 384))
    assert nodense.load_chunks_and_index(world["csv"], "/no/such.index")
    got = nodense.retrieve(qs[0], top_k=5)
    sc = bm.get_scores(qs[0].split())
    want = O.hybrid_fuse([], [(chunks[j], sc[j]) for j in O.argsort_topk_canonical(sc, 10)], 5)
    assert [c["id"] for c, _ in got] == [c["id"] for c, _ in want]


class DeviceEncoder(FakeEncoder):
    """An encoder whose output is already on the GPU (what pooling.FusedPoolingEncoder.encode_device or
    SentenceTransformer.encode(convert_to_tensor=True) return): no host copy of the embeddings may happen."""

    def encode(self, sentences, device=None, convert_to_tensor=False, **_):
        import torch
        out = super().encode(sentences)
        if not convert_to_tensor:
            raise AssertionError("the retriever must ask for a tensor (convert_to_tensor=True)")
        return torch.from_numpy(out).cuda()


def test_device_resident_embeddings_go_straight_into_the_scan(P, world):
    """f-3: src/retrieval.py:98-102 does device -> host -> device; here CUDA embeddings feed the scan kernel."""
    host = P.RetrievalSystem(method="dense", encoder=FakeEncoder(world["table"], 384))
    devr = P.RetrievalSystem(method="dense", encoder=DeviceEncoder(world["table"], 384), storage="fp32")
    assert host.load_chunks_and_index(world["csv"], world["index"]) and devr.load_chunks_and_index(world["csv"], world["index"])
    qs = [qd["question"] for qd in world["queries"][:24]]
    a = host.retrieve_batch(qs, 7)
    b = devr.retrieve_batch(qs, 7)
    assert [[c["id"] for c, _ in h] for h in a] == [[c["id"] for c, _ in h] for h in b]
    np.testing.assert_allclose([[s for _, s in h] for h in a], [[s for _, s in h] for h in b], rtol=1e-6)
    assert [c["id"] for c, _ in devr.retrieve(qs[3], 7)] == [c["id"] for c, _ in b[3]]
    # the evaluator's loop as one batch (src/evaluation.py:273-299): same contexts as the per-query calls
    ctx = devr.get_contexts_for_rag_batch(qs[:6], top_k=5, max_context_length=2000)
    assert ctx == [devr.get_contexts_for_rag(q, top_k=5, max_context_length=2000) for q in qs[:6]]
    assert devr.evaluate_retrieval_quality(world["queries"], world["relevant"]) == host.evaluate_retrieval_quality(world["queries"], world["relevant"])


def test_multi_model_retrieval(P, world):
    m = P.MultiModelRetrieval(["models/a-model"], encoders={"a-model": FakeEncoder(world["table"], 384)})
    m.setup_retrievers(world["csv"], {"a-model": world["index"]})
    out = m.compare_retrieval_performance(world["queries"], world["relevant"])
    assert set(out) == {"a-model"} and out["a-model"]["total_queries"] == len(world["queries"])
    m.cleanup_all()
    assert m.retrievers == {}


# ------------------------------------------------------------------ row sharding (emulated on one device)
@pytest.mark.parametrize("G", [2, 3, 8])
@pytest.mark.parametrize("metric", [O.METRIC_L2, O.METRIC_IP])
def test_sharded_equals_unsharded(P, G, metric):
    """G shards as G sub-indices on one device + the same merge kernel the
    multi-GPU path runs after ncclAllGather.  Sharded == unsharded bit for bit, ties included."""
    import torch
    from persian_rag_system_b200.sharded import merge_topk, shard_bounds
    rng = np.random.default_rng(G)
    base = rng.standard_normal((1500, 64)).astype(np.float32)
    x = np.concatenate([base, base[:300], base[:100]])          # duplicates across shards -> ties on global id
    n = x.shape[0]
    q = np.concatenate([base[:40] + 0.01 * rng.standard_normal((40, 64)).astype(np.float32), base[:8]])
    k = 10
    whole = P.FlatIndex(64, metric)
    whole.add(x)
    Dw, Iw = whole.search(q, k)
    Dp, Ip = [], []
    for g in range(G):
        lo, hi = shard_bounds(n, G, g)
        sh = P.FlatIndex(64, metric)
        sh.add(x[lo:hi])
        sh.set_id_offset(lo)
        D, I = sh.search(torch.from_numpy(q).cuda(), k)
        Dp.append(D)
        Ip.append(I)
    D, I = merge_topk(torch.stack(Dp), torch.stack(Ip), largest=metric == O.METRIC_IP)
    assert np.array_equal(I.cpu().numpy(), Iw) and np.array_equal(D.cpu().numpy(), Dw)
    Dr, Ir = O.flat_search_c(x, q, k, metric, form=1)
    assert np.array_equal(Iw, Ir)


# ------------------------------------------------------------------ index build mirror (§8 a-2)
def test_create_model_embeddings_writes_the_faiss_file_the_reference_would(P, world, tmp_path, gold_dir, golden_indices):
    """encode in batches of 32 -> IndexFlatL2.add -> write_index: the file must equal, byte for byte,
    what faiss writes for the same float32 rows (oracle writer pinned on the reference's shipped files)."""
    import pandas as pd
    name = "drugs_sentence_chunks.index"
    x, _ = golden_indices[name]
    texts = [f"chunk {i}" for i in range(x.shape[0])]
    csv = tmp_path / "chunks.csv"
    pd.DataFrame({"id": [f"sentence_chunk_{i}" for i in range(len(texts))], "text": texts}).to_csv(csv, index=False)
    enc = FakeEncoder({t: x[i] for i, t in enumerate(texts)}, x.shape[1])
    out_dir = str(tmp_path / "faiss")
    assert P.create_model_embeddings("models/distiluse-ft", str(csv), "sentence", encoder=enc, faiss_dir=out_dir) is True
    path = P.index_path_for("models/distiluse-ft", "sentence", out_dir)
    want = tmp_path / "want.index"
    O.write_faiss_flat(str(want), x, O.METRIC_L2)
    assert open(path, "rb").read() == open(want, "rb").read() == open(os.path.join(gold_dir, "indices", name), "rb").read()
    # second call: skip-if-exists
    assert P.create_model_embeddings("models/distiluse-ft", str(csv), "sentence", encoder=None, faiss_dir=out_dir) is True
    # setup_faiss_index: same rows, exact L2, searchable at once; the IVF branch is replaced by the exact scan
    idx = P.setup_faiss_index(x, index_type="ivf")
    D, I = idx.search(x[:5], 1)
    assert I[:, 0].tolist() == [0, 1, 2, 3, 4] and np.allclose(D[:, 0], 0.0, atol=1e-6)


# ------------------------------------------------------------------ IVF-Flat branch of the reference's builder (§8 f-4)
def test_ivf_flat_branch_matches_the_restated_faiss_semantics(P):
    """scripts/phase3_pdf_chunking.py:45-57: >= 1000 embeddings -> IndexIVFFlat(IndexFlatL2(d), d, min(100, max(10, n//20))),
    trained on the first 10 000 rows, searched with nprobe = 1.  k-means (faiss's permutation-based initialisation,
    10 Lloyd iterations, float32 row-order centroid sums), list assignment and the probed-list scan are compared
    with the oracle restatement; every distance computation runs on the flat-scan kernels."""
    rng = np.random.default_rng(4)
    d, n = 32, 2400
    centers = rng.standard_normal((40, d)).astype(np.float32) * 8
    x = (centers[rng.integers(0, 40, n)] + 0.4 * rng.standard_normal((n, d))).astype(np.float32)
    q = (x[rng.integers(0, n, 50)] + 0.05 * rng.standard_normal((50, d))).astype(np.float32)
    idx = P.setup_faiss_index(x, index_type="ivf")
    assert isinstance(idx, P.IndexIVFFlat) and idx.nlist == min(100, max(10, n // 20)) == 100 and idx.nprobe == 1
    assert idx.ntotal == n and idx.is_trained and int(idx.list_sizes().sum()) == n
    ora = O.IVFFlatOracle(d, 100)
    ora.train(x[:10000])
    # same permutation, same row-order float32 sums: the centroids agree to rounding wherever no training row sits on a
    # cluster boundary within fp32 rounding (assignment near-ties may fall on either side in any implementation)
    same = np.isclose(idx.centroids, ora.centroids, rtol=1e-5, atol=1e-6).all(axis=1)
    assert same.mean() > 0.9, same.mean()
    # from identical centroids on: identical lists and identical search results
    both = P.IndexIVFFlat(d, 100)
    both.set_centroids(ora.centroids)
    for s in range(0, n, 1000):
        both.add(x[s:s + 1000])
    ora.add(x)
    assert [sorted(l) for l in ora.lists] == [a.tolist() for a in both._ids]
    for k in (1, 5, 64):
        D, I = both.search(q, k)
        Do, Io = ora.search(q, k)
        O.check_topk_lists(I, D, Io, Do, rtol=1e-5, atol=1e-6, what=f"ivf k{k}")
        assert (I == Io).mean() > 0.99
    # a list with fewer than k rows pads with (FLT_MAX, -1), like faiss's heap
    small = int(np.argmin(np.where(both.list_sizes() > 0, both.list_sizes(), 10**9)))
    qs = ora.centroids[small:small + 1]
    D, I = both.search(qs, 64)
    nrows = int(both.list_sizes()[small])
    assert (I[0, :nrows] >= 0).all() and (I[0, nrows:] == -1).all() and (D[0, nrows:] == np.finfo(np.float32).max).all()
    # the retriever works on top of it unchanged (src/retrieval.py:102 only needs .search / .ntotal)
    r = P.RetrievalSystem(method="dense", encoder=FakeEncoder({f"q{i}": q[i] for i in range(50)}, d))
    r.chunks = [{"id": f"c{i}", "text": "t"} for i in range(n)]
    r.faiss_index, r.is_ready = both, True
    hits = r.retrieve("q3", top_k=5)
    assert [c["id"] for c, _ in hits] == [f"c{j}" for j in Io[3, :5] if j >= 0][:len(hits)] or len(hits) == 5
    # below 1000 rows, or index_type "flat", the builder stays flat; exact=True opts out of the approximate branch
    assert isinstance(P.setup_faiss_index(x[:999], index_type="ivf"), P.FlatIndex)
    assert isinstance(P.setup_faiss_index(x, index_type="flat"), P.FlatIndex)
    assert isinstance(P.setup_faiss_index(x, index_type="ivf", exact=True), P.FlatIndex)
