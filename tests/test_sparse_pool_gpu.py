"""GPU parity: BM25 / TF-IDF scoring + top-k and the pooling epilogue vs the CPU oracle."""
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


# ------------------------------------------------------------------ BM25
def test_bm25_on_reference_chunks_bit_exact(P, gold_dir, golden_texts):
    chunks, queries = golden_texts
    texts = [c["text"] for c in chunks]
    gold = np.load(os.path.join(gold_dir, "bm25_golden.npz"))["scores"]       # oracle scores [nq, n]
    bm = P.BM25Index([t.split() for t in texts])
    n = len(texts)
    for k in (1, 5, 10, n, n + 5):
        S, I = bm.search([q.split() for q in queries], k)
        for r in range(len(queries)):
            want = O.argsort_topk_canonical(gold[r], k)
            kk = min(k, n)
            assert I[r, :kk].tolist() == want.tolist(), f"q{r} k{k}"
            assert np.array_equal(S[r, :kk], gold[r][want])                    # float64, bit for bit
            assert (I[r, kk:] == -1).all()
            # and it is a valid answer to what the reference computes (unstable argsort, finding 6)
            ref = O.argsort_topk_reference(gold[r], k)
            assert np.array_equal(np.sort(gold[r][ref])[::-1], S[r, :kk])


def test_bm25_fp32_weights_within_tolerance(P, gold_dir, golden_texts):
    chunks, queries = golden_texts
    texts = [c["text"] for c in chunks]
    gold = np.load(os.path.join(gold_dir, "bm25_golden.npz"))["scores"]
    bm = P.BM25Index([t.split() for t in texts], dtype="float32")
    S, I = bm.search([q.split() for q in queries], 10)
    for r in range(len(queries)):
        O.check_topk_against_scores(I[r], S[r], gold[r], 10, True, rtol=1e-5, atol=1e-9, what=f"q{r}")


def test_bm25_random_corpus_vs_oracle(P):
    rng = np.random.default_rng(11)
    vocab = [f"w{i}" for i in range(3000)]
    p = 1.0 / np.arange(1, 3001) ** 1.07
    p /= p.sum()
    docs = [[vocab[j] for j in rng.choice(3000, size=int(rng.integers(1, 200)), p=p)] for _ in range(20011)]
    qs = [[vocab[j] for j in rng.choice(3000, size=int(rng.integers(1, 9)), p=p)] for _ in range(24)] + [["zzz"], []]
    bm = P.BM25Index(docs)
    ob = O.BM25OkapiOracle(docs)
    S, I = bm.search(qs, 10)
    for r, q in enumerate(qs):
        sc = ob.get_scores(q)
        O.check_topk_against_scores(I[r], S[r], sc, 10, True, rtol=1e-5, atol=1e-12, what=f"q{r}")
        assert np.array_equal(S[r], sc[I[r]])                                   # exact float64 scores
    # no query token matches: every doc scores 0, order is id descending (stable argsort reversed)
    assert I[-1].tolist() == list(range(20010, 20000, -1)) and I[-2].tolist() == I[-1].tolist()
    assert bm.index.last_postings > 0


# ------------------------------------------------------------------ TF-IDF
def test_tfidf_on_reference_chunks_bit_exact_vs_sklearn(P, gold_dir, golden_texts):
    chunks, queries = golden_texts
    texts = [c["text"] for c in chunks]
    gold = np.load(os.path.join(gold_dir, "tfidf_golden.npz"))["scores"]      # sklearn cosine_similarity
    tf = P.TfidfIndex(texts, max_features=10000, ngram_range=(1, 2))
    for k in (1, 5, 10):
        S, I = tf.search(queries, k)
        for r in range(len(queries)):
            want = O.argsort_topk_canonical(gold[r], k)
            assert I[r].tolist() == want.tolist(), f"q{r} k{k}"
            assert np.array_equal(S[r], gold[r][want])


def test_tfidf_random_text_vs_sklearn(P):
    rng = np.random.default_rng(12)
    words = ["".join(chr(0x0627 + int(c)) for c in rng.integers(0, 30, size=int(rng.integers(2, 7)))) for _ in range(800)]
    texts = [" ".join(rng.choice(words, size=int(rng.integers(5, 120)))) for _ in range(700)]
    queries = [" ".join(rng.choice(words, size=int(rng.integers(1, 8)))) for _ in range(16)]
    vec, mat = O.tfidf_fit(texts)             # > 10000 (1,2)-gram features: max_features pruning is exercised
    tf = P.TfidfIndex(texts)
    assert tf.vocabulary_ == {k: int(v) for k, v in vec.vocabulary_.items()}
    S, I = tf.search(queries, 10)
    for r, q in enumerate(queries):
        sc = O.tfidf_scores(vec, mat, q)
        O.check_topk_against_scores(I[r], S[r], sc, 10, True, rtol=1e-5, atol=1e-12, what=f"q{r}")


def test_sparse_raw_csr_api_large_docs_multi_tile(P):
    """> 8192 docs (several accumulator tiles, several CTAs per query) and a stop-word-like term."""
    rng = np.random.default_rng(13)
    n_docs, n_terms = 50_000, 500
    rows = []
    for dct in range(n_docs):
        t = np.unique(np.concatenate([[0], rng.integers(1, n_terms, size=int(rng.integers(1, 12)))]))
        rows.append(t)
    indptr = np.zeros(n_docs + 1, np.int64)
    indptr[1:] = np.cumsum([len(r) for r in rows])
    indices = np.concatenate(rows).astype(np.int32)
    vals = rng.random(indices.shape[0]).astype(np.float32) + 0.1
    sp = P.SparseIndex(indptr, indices, vals, n_terms)
    assert sp.ndocs == n_docs and sp.nnz == indices.shape[0]
    q_terms = np.array([0, 7, 7, 499, 1000, 3, 0], np.int32)          # 1000 is out of vocabulary
    q_indptr = np.array([0, 5, 7], np.int64)
    q_w = np.array([1.0, 0.5, 0.5, 2.0, 9.0, 1.0, 0.25])
    S, I = sp.search(q_indptr, q_terms, q_w, 20)
    import scipy.sparse as ssp
    M = ssp.csr_matrix((vals.astype(np.float64), indices, indptr), shape=(n_docs, n_terms)).tocsc()
    for r in range(2):
        sc = np.zeros(n_docs)
        for e in range(q_indptr[r], q_indptr[r + 1]):
            if q_terms[e] < n_terms:
                sc += q_w[e] * M[:, q_terms[e]].toarray().ravel()
        O.check_topk_against_scores(I[r], S[r], sc, 20, True, rtol=1e-6, atol=1e-12, what=f"q{r}")
        assert np.array_equal(S[r], sc[I[r]])


# ------------------------------------------------------------------ throughput mode (batched queries, fixed-point selection)
def test_throughput_mode_on_reference_chunks_equals_exact_mode(P, gold_dir, golden_texts):
    chunks, queries = golden_texts
    texts = [c["text"] for c in chunks]
    gold = np.load(os.path.join(gold_dir, "bm25_golden.npz"))["scores"]
    bm = P.BM25Index([t.split() for t in texts], mode="throughput")
    assert bm.index.mode == "throughput"
    n = len(texts)
    for k in (1, 5, 10, n, n + 5):
        S, I = bm.search([q.split() for q in queries], k)
        for r in range(len(queries)):
            want = O.argsort_topk_canonical(gold[r], k)
            kk = min(k, n)
            assert I[r, :kk].tolist() == want.tolist(), f"q{r} k{k}"
            assert np.array_equal(S[r, :kk], gold[r][want])                    # exact float64 re-score
            assert (I[r, kk:] == -1).all()
    tgold = np.load(os.path.join(gold_dir, "tfidf_golden.npz"))["scores"]
    tf = P.TfidfIndex(texts, mode="throughput")
    S, I = tf.search(queries, 10)
    for r in range(len(queries)):
        O.check_topk_against_scores(I[r], S[r], tgold[r], 10, True, rtol=1e-5, atol=1e-12, what=f"tfidf q{r}")
        assert np.array_equal(S[r], tgold[r][I[r]])


def test_throughput_mode_random_corpus_vs_oracle(P):
    rng = np.random.default_rng(11)
    vocab = [f"w{i}" for i in range(3000)]
    p = 1.0 / np.arange(1, 3001) ** 1.07
    p /= p.sum()
    docs = [[vocab[j] for j in rng.choice(3000, size=int(rng.integers(1, 200)), p=p)] for _ in range(20011)]
    qs = [[vocab[j] for j in rng.choice(3000, size=int(rng.integers(1, 9)), p=p)] for _ in range(43)] + [["zzz"], []]
    qs.append([vocab[j] for j in rng.choice(3000, size=150, p=p)])            # > 64 entries: the multi-round path of its group
    ob = O.BM25OkapiOracle(docs)
    exact = P.BM25Index(docs)
    fast = P.BM25Index(docs, mode="throughput")
    Se, Ie = exact.search(qs, 10)
    S, I = fast.search(qs, 10)
    for r, q in enumerate(qs):
        sc = ob.get_scores(q)
        O.check_topk_against_scores(I[r], S[r], sc, 10, True, rtol=1e-5, atol=1e-12, what=f"q{r}")
        assert np.array_equal(S[r], sc[I[r]])                                   # exact float64 scores
    assert np.array_equal(S, Se) and (I == Ie).mean() > 0.99                    # same scores; ids differ at most inside ties
    # no query token matches: every doc scores 0, order is id descending (stable argsort reversed)
    assert I[-2].tolist() == list(range(20010, 20000, -1)) and I[-3].tolist() == I[-2].tolist()
    assert fast.index.last_postings == exact.index.last_postings > 0
    # device-side entry point: same answer, results stay on the GPU
    Sd, Id = fast.search_device(qs, 10)
    assert Sd.is_cuda and np.array_equal(Sd.cpu().numpy(), S) and np.array_equal(Id.cpu().numpy(), I)
    assert fast.index.last_postings == exact.index.last_postings                # counted by a kernel on that path
    Sd, Id = exact.search_device(qs[:5], 3)
    assert np.array_equal(Sd.cpu().numpy(), Se[:5, :3]) and np.array_equal(Id.cpu().numpy(), Ie[:5, :3])
    # "auto" (what RetrievalSystem uses): below 16 queries the exact-order kernel, from 16 on the throughput kernel
    auto = P.BM25Index(docs, mode="auto")
    assert auto.index.mode == "auto"
    Sa, Ia = auto.search(qs[:15], 10)
    assert np.array_equal(Sa, Se[:15]) and np.array_equal(Ia, Ie[:15])
    Sa, Ia = auto.search(qs, 10)
    assert np.array_equal(Sa, S) and np.array_equal(Ia, I)


def test_throughput_mode_raw_csr_many_tiles_negative_weights_and_shared_terms(P):
    """150 000 docs (74 accumulator tiles of 2 048, several doc-range parts), a term present in every doc
    (tile-offset row), query weights of both signs, duplicates of a term inside and across the queries of
    one 8-query group, out-of-vocabulary ids, k + margin above and below the candidate-list sizes."""
    rng = np.random.default_rng(17)
    n_docs, n_terms = 150_000, 700
    lens = rng.integers(1, 14, size=n_docs)
    indptr = np.zeros(n_docs + 1, np.int64)
    indptr[1:] = np.cumsum(lens + 1)
    indices = np.empty(indptr[-1], np.int32)
    for dct in range(n_docs):
        t = rng.choice(n_terms - 1, size=lens[dct], replace=False) + 1
        indices[indptr[dct]:indptr[dct + 1]] = np.sort(np.concatenate([[0], t]))
    vals = (rng.random(indices.shape[0]) + 0.1) * np.where(rng.random(indices.shape[0]) < 0.1, -1.0, 1.0)
    import scipy.sparse as ssp
    M = ssp.csr_matrix((vals, indices, indptr), shape=(n_docs, n_terms)).tocsc()
    nq = 21
    qlen = rng.integers(1, 12, size=nq)
    q_indptr = np.zeros(nq + 1, np.int64)
    q_indptr[1:] = np.cumsum(qlen)
    q_terms = rng.integers(0, 40, size=q_indptr[-1]).astype(np.int32)          # few distinct terms: heavy sharing
    q_terms[::7] = 0
    q_terms[5] = 5000                                                           # out of vocabulary
    q_w = rng.standard_normal(q_indptr[-1])
    for dtype in (np.float64, np.float32):
        v = vals.astype(dtype)
        Md = ssp.csr_matrix((v.astype(np.float64), indices, indptr), shape=(n_docs, n_terms)).tocsc()
        sp = P.SparseIndex(indptr, indices, v, n_terms, mode="throughput")
        for k in (1, 10, 80, 200):                                             # 200 + 16 > 96: that call runs the exact kernel
            S, I = sp.search(q_indptr, q_terms, q_w, k)
            for r in range(nq):
                sc = np.zeros(n_docs)
                for e in range(q_indptr[r], q_indptr[r + 1]):
                    if q_terms[e] < n_terms:
                        sc += q_w[e] * Md[:, q_terms[e]].toarray().ravel()
                O.check_topk_against_scores(I[r], S[r], sc, k, True, rtol=1e-5, atol=1e-9, what=f"{dtype.__name__} k{k} q{r}")
                assert np.array_equal(S[r], sc[I[r]])


# ------------------------------------------------------------------ doc-range shards (emulated on one device)
@pytest.mark.parametrize("G", [2, 5])
@pytest.mark.parametrize("mode", ["exact", "throughput"])
def test_doc_range_sharded_sparse_equals_unsharded(P, G, mode):
    """every shard holds the postings of its doc block (global idf / avgdl baked into the weights);
    local top-k with global ids -> float64 merge kernel with ties on global ids (id DESC) == unsharded."""
    import torch
    from persian_rag_system_b200.sharded import merge_topk_f64, shard_bounds
    from persian_rag_system_b200.sparse import build_bm25_csr
    rng = np.random.default_rng(G)
    vocab = [f"w{i}" for i in range(400)]
    base = [[vocab[j] for j in rng.integers(0, 400, size=int(rng.integers(1, 40)))] for _ in range(3000)]
    docs = base + base[:700] + base[:50]                      # duplicate docs across shards: exact score ties
    qs = [base[i][:4] for i in range(12)] + [["nope"], [vocab[3], vocab[3], vocab[7]]]
    b = build_bm25_csr(docs)
    whole = P.SparseIndex(b["indptr"], b["indices"], b["weights"], len(b["vocab"]), mode=mode)
    bm = P.BM25Index(docs)                                    # only for encode_queries
    ip, qt, qw = bm.encode_queries(qs)
    dev = torch.device("cuda", 0)
    dip, dqt, dqw = (torch.from_numpy(v).to(dev) for v in (ip, qt, qw))
    k = 10
    Sw, Iw = whole.search_device(dip, dqt, dqw, k)
    Sp, Ip = [], []
    n = len(docs)
    for g in range(G):
        lo, hi = shard_bounds(n, G, g)
        a, e = int(b["indptr"][lo]), int(b["indptr"][hi])
        sh = P.SparseIndex(b["indptr"][lo:hi + 1] - b["indptr"][lo], b["indices"][a:e], b["weights"][a:e], len(b["vocab"]), mode=mode)
        sh.set_id_offset(lo)
        S, I = sh.search_device(dip, dqt, dqw, k)
        Sp.append(S)
        Ip.append(I)
    S, I = merge_topk_f64(torch.stack(Sp), torch.stack(Ip))
    assert torch.equal(S, Sw) and torch.equal(I, Iw)
    ob = O.BM25OkapiOracle(docs)
    for r, q in enumerate(qs):
        sc = ob.get_scores(q)
        assert I[r].tolist() == O.argsort_topk_canonical(sc, k).tolist() and np.array_equal(S[r].cpu().numpy(), sc[I[r].cpu().numpy()])


# ------------------------------------------------------------------ pooling epilogue
def test_pool_golden(P, gold_dir):
    import torch
    g = np.load(os.path.join(gold_dir, "pool_golden.npz"))
    h = torch.from_numpy(g["hidden"]).cuda()
    m = torch.from_numpy(g["mask"]).cuda()
    out = P.mean_pool_normalize(h, m, False).cpu().numpy()
    np.testing.assert_allclose(out, g["pooled"], rtol=1e-5, atol=1e-6)
    out = P.mean_pool_normalize(h, m, True).cpu().numpy()
    np.testing.assert_allclose(out, g["normalized"], rtol=1e-5, atol=1e-6)
    assert not np.isnan(out).any()                       # the fully-masked row (len 0) stays finite


@pytest.mark.parametrize("B,T,H,dtype", [(1, 128, 384, "float32"), (32, 128, 768, "float16"), (16, 512, 768, "bfloat16"), (3, 7, 512, "float32")])
def test_pool_random_vs_torch(P, B, T, H, dtype):
    import torch
    g = torch.Generator(device="cuda").manual_seed(B * T + H)
    h = torch.randn((B, T, H), generator=g, device="cuda").to(getattr(torch, dtype))
    lens = torch.randint(1, T + 1, (B,), generator=g, device="cuda")
    mask = (torch.arange(T, device="cuda")[None, :] < lens[:, None]).to(torch.int64)
    me = mask.unsqueeze(-1).float()
    pooled = (h.float() * me).sum(1) / me.sum(1).clamp(min=1e-9)
    for normalize in (False, True):
        want = torch.nn.functional.normalize(pooled, p=2, dim=1) if normalize else pooled
        got = P.mean_pool_normalize(h, mask, normalize)
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("B,T,H,dtype,kind", [
    (5, 100, 768, "float16", "left"),        # left-padded batch: the copied row range starts inside the slice
    (7, 77, 384, "float32", "holes"),        # masked rows between unmasked ones are skipped, not added
    (4, 64, 512, "bfloat16", "empty"),       # fully masked sequences: zeros, no NaN
    (1300, 40, 384, "float16", "prefix"),    # B large enough for one CTA per sequence (no token split)
    (3, 300, 1024, "float32", "prefix"),     # 256 lanes per row (R = 1), slices of 38 tokens
    (9, 33, 64, "float16", "prefix"),        # 8 lanes per row: 32 rows per pass
    (2, 9000, 128, "float16", "holes"),      # long sequences
    (6, 50, 1032, "float32", "prefix"),      # 258 lanes per row: the one-CTA-per-sequence kernel
    (6, 50, 100, "float16", "prefix"),       # H not a multiple of 8: the one-CTA-per-sequence kernel
])
def test_pool_mask_shapes_vs_torch(P, B, T, H, dtype, kind):
    import torch
    g = torch.Generator(device="cuda").manual_seed(B * 7 + T + H)
    h = torch.randn((B, T, H), generator=g, device="cuda").to(getattr(torch, dtype))
    lens = torch.randint(1, T + 1, (B,), generator=g, device="cuda")
    ar = torch.arange(T, device="cuda")[None, :]
    if kind == "prefix":
        mask = ar < lens[:, None]
    elif kind == "left":
        mask = ar >= (T - lens)[:, None]
    elif kind == "holes":
        mask = (ar < lens[:, None]) & (torch.rand((B, T), generator=g, device="cuda") < 0.6)
    else:
        mask = torch.zeros((B, T), dtype=torch.bool, device="cuda")
        mask[0, : T // 2] = True
    mask = mask.to(torch.int64)
    me = mask.unsqueeze(-1).float()
    pooled = (h.float() * me).sum(1) / me.sum(1).clamp(min=1e-9)
    for normalize in (False, True):
        want = torch.nn.functional.normalize(pooled, p=2, dim=1) if normalize else pooled
        got = P.mean_pool_normalize(h, mask, normalize)
        assert not torch.isnan(got).any()
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
        assert torch.equal(got, P.mean_pool_normalize(h, mask, normalize))       # deterministic


def test_pool_then_search_stays_on_device(P):
    """f-3: encoder output -> pool/normalise kernel -> flat search, no host hop in between."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(21)
    corpus_h = torch.randn((500, 16, 384), generator=g, device="cuda")
    cmask = torch.ones((500, 16), dtype=torch.int64, device="cuda")
    emb = P.mean_pool_normalize(corpus_h, cmask, True)
    idx = P.IndexFlatL2(384)
    idx.add(emb)
    D, I = idx.search(P.mean_pool_normalize(corpus_h[:9], cmask[:9], True), 1)
    assert I[:, 0].tolist() == list(range(9)) and float(D.abs().max()) < 1e-6
