"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the third-party calls at the bottom of the reference's retriever
(`/root/reference/src/retrieval.py`).  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s `cpu_baseline` / `--impl reference` leg may import this module; the product
package (`persian-rag-system_b200/`) never does and fails loudly without its CUDA library.

What is restated, and from where (reference file:line -> third-party call):

* flat dense search  -- src/retrieval.py:102, src/create_embeddings.py:130-136,291,
  scripts/phase3_pdf_chunking.py:47,430,441 -> faiss-cpu==1.7.4 `IndexFlatL2/IP.search`
  (requirements.txt:9).  C restatement in `flat_oracle.c`, numpy restatement here.
  PARITY UNPINNED: faiss is not vendored in the reference, not installed, and the reference
  has no tests for it.
* on-disk index      -- src/create_embeddings.py:136, src/retrieval.py:55 -> faiss
  `write_index/read_index` of an IndexFlat (`IxF2` / `IxFI`).  PINNED by the 14 index files
  the reference ships under results/faiss/ (byte-exact round trip, tests/test_oracle.py).
* BM25               -- src/retrieval.py:66-67,127,130 -> rank_bm25==0.2.2 `BM25Okapi`
  (requirements.txt:5) + `np.argsort(scores)[::-1][:k]`.  PARITY UNPINNED (same reason).
* TF-IDF             -- src/retrieval.py:78-83,152-159 -> scikit-learn `TfidfVectorizer` +
  `cosine_similarity`.  scikit-learn IS installed here, so this one is the real reference
  implementation, imported; golden vectors generated from it are committed in tests/golden/.
* pooling epilogue   -- src/retrieval.py:98 -> sentence-transformers `Pooling(mean)` +
  `Normalize`; restated in numpy and pinned against torch ops (installed).
* hybrid fusion      -- src/retrieval.py:174-220 (the reference's own code, restated).
* Hit@K / MRR        -- src/retrieval.py:274-323 (the reference's own code, restated).
"""
from __future__ import annotations

import ctypes
import math
import os
import struct
import subprocess
from typing import Dict, List, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
METRIC_IP = 0   # faiss METRIC_INNER_PRODUCT
METRIC_L2 = 1   # faiss METRIC_L2

# --------------------------------------------------------------------------------------
# The real third-party libraries, when they exist (requirements.txt:5,9 of the reference pin
# faiss-cpu==1.7.4 and rank_bm25==0.2.2; neither wheel is in this image and there is no network).
# The moment one becomes importable -- in site-packages or under baseline/_ref/ -- the tests in
# tests/test_oracle.py cross-check the restatements against it and bench.py's reference arm
# times it (`cpu_baseline.kind: "reference"`).  Until then both stay "parity unpinned".
# --------------------------------------------------------------------------------------
_REF_DIRS = (os.path.join(os.path.dirname(_HERE), "baseline", "_ref"),)


def reference_library(name: str):
    """`faiss` / `rank_bm25` module if importable (also from baseline/_ref), else None."""
    import importlib
    import sys
    for extra in (None,) + _REF_DIRS:
        if extra is not None:
            if not os.path.isdir(extra) or extra in sys.path:
                continue
            sys.path.append(extra)
        try:
            return importlib.import_module(name)
        except Exception:                                  # ImportError, or a wheel built for another ABI
            continue
    return None


def faiss_search(x, q, k: int, metric: int = METRIC_L2):
    """The reference's own call chain (src/create_embeddings.py:130-133, src/retrieval.py:102) on real
    faiss: IndexFlatL2/IP(d); add(x); search(q, k).  Raises RuntimeError when faiss is absent."""
    faiss = reference_library("faiss")
    if faiss is None:
        raise RuntimeError("faiss is not importable here")
    x = np.ascontiguousarray(x, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    index = faiss.IndexFlatL2(x.shape[1]) if metric == METRIC_L2 else faiss.IndexFlatIP(x.shape[1])
    index.add(x)
    return index.search(q, k)


# --------------------------------------------------------------------------------------
# C oracle loader
# --------------------------------------------------------------------------------------
_clib = None


def build_c_oracle(force: bool = False) -> str:
    so = os.path.join(_HERE, "_ref", "liboracle.so")
    src = os.path.join(_HERE, "flat_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def _c():
    global _clib
    if _clib is None:
        lib = ctypes.CDLL(build_c_oracle())
        f32p = ctypes.POINTER(ctypes.c_float)
        i64p = ctypes.POINTER(ctypes.c_int64)
        lib.oracle_flat_search.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p, ctypes.c_int64,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_int, f32p, i64p]
        lib.oracle_flat_search.restype = ctypes.c_int
        lib.oracle_flat_search_sort.argtypes = [f32p, ctypes.c_int64, ctypes.c_int, f32p, ctypes.c_int64,
                                                ctypes.c_int, ctypes.c_int, f32p, i64p]
        lib.oracle_flat_search_sort.restype = ctypes.c_int
        _clib = lib
    return _clib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def flat_search_c(x, q, k: int, metric: int = METRIC_L2, form: int = 0):
    """faiss IndexFlat.search restated in scalar C (heap semantics).  form: 0 auto (direct when
    nq < 20 else expanded, like faiss), 1 direct, 2 expanded."""
    x, xp = _f32(x)
    q, qp = _f32(q)
    n, d = (x.shape if x.ndim == 2 else (0, q.shape[1]))
    nq = q.shape[0]
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    rc = _c().oracle_flat_search(xp, n, d, qp, nq, k, metric, form,
                                 D.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                 I.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    assert rc == 0
    return D, I


def flat_search_c_sort(x, q, k: int, metric: int = METRIC_L2):
    x, xp = _f32(x)
    q, qp = _f32(q)
    n, d = x.shape
    nq = q.shape[0]
    D = np.empty((nq, k), np.float32)
    I = np.empty((nq, k), np.int64)
    rc = _c().oracle_flat_search_sort(xp, n, d, qp, nq, k, metric,
                                      D.ctypes.data_as(ctypes.POINTER(ctypes.c_float)),
                                      I.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    assert rc == 0
    return D, I


# --------------------------------------------------------------------------------------
# numpy restatement (blocked sgemm + exact selection): usable at 1M x 768 and as the CPU
# baseline "port" (this is what faiss does for nq >= 20: sgemm blocks + norms + heap).
# --------------------------------------------------------------------------------------

def canonical_topk(vals: np.ndarray, k: int, largest: bool, ids: np.ndarray | None = None):
    """k best of a 1-D score vector under the canonical order (value, then id ASC)."""
    n = vals.shape[0]
    if ids is None:
        ids = np.arange(n, dtype=np.int64)
    key = -vals if largest else vals
    kk = min(k, n)
    if kk < n:
        # argpartition then widen to include every element tied with the k-th value
        part = np.argpartition(key, kk - 1)[:kk]
        kth = key[part].max()
        cand = np.nonzero(key <= kth)[0]
    else:
        cand = np.arange(n)
    order = np.lexsort((ids[cand], key[cand]))[:kk]
    sel = cand[order]
    return vals[sel], ids[sel]


def flat_search_np(x, q, k: int, metric: int = METRIC_L2, form: int = 2, block: int = 262144,
                   x_sqnorm: np.ndarray | None = None):
    """Blocked numpy search. form 2 (default) = expanded ||q||^2+||x||^2-2q.x (faiss nq>=20 path),
    form 1 = direct sum (q-x)^2 in float32 (slow, small inputs only)."""
    x = np.asarray(x)
    q = np.ascontiguousarray(q, dtype=np.float32)
    n, d = x.shape
    nq = q.shape[0]
    largest = metric == METRIC_IP
    fill = -np.finfo(np.float32).max if largest else np.finfo(np.float32).max
    D = np.full((nq, k), fill, np.float32)
    I = np.full((nq, k), -1, np.int64)
    best_v = [np.empty(0, np.float32) for _ in range(nq)]
    best_i = [np.empty(0, np.int64) for _ in range(nq)]
    qn = (q * q).sum(1) if metric == METRIC_L2 else None
    for b0 in range(0, n, block):
        xb = np.asarray(x[b0:b0 + block], dtype=np.float32)
        if metric == METRIC_L2 and form == 1:
            s = np.empty((nq, xb.shape[0]), np.float32)
            for i in range(nq):
                diff = xb - q[i]
                s[i] = np.einsum("ij,ij->i", diff, diff)
        else:
            s = q @ xb.T
            if metric == METRIC_L2:
                xn = (xb * xb).sum(1) if x_sqnorm is None else x_sqnorm[b0:b0 + block]
                s = qn[:, None] + xn[None, :] - 2.0 * s
                np.maximum(s, 0.0, out=s)
        ids = np.arange(b0, b0 + xb.shape[0], dtype=np.int64)
        for i in range(nq):
            v, ii = canonical_topk(s[i], k, largest, ids)
            v = np.concatenate([best_v[i], v])
            ii = np.concatenate([best_i[i], ii])
            best_v[i], best_i[i] = canonical_topk(v, k, largest, ii)
    for i in range(nq):
        m = best_v[i].shape[0]
        D[i, :m] = best_v[i]
        I[i, :m] = best_i[i]
    return D, I


def flat_search_np_threshold(x, q, k: int, metric: int = METRIC_L2, block: int = 65536,
                             x_sqnorm: np.ndarray | None = None):
    """Same result as flat_search_np (expanded form), organised the way faiss's nq >= 20 path is
    (exhaustive_inner_product_blas / exhaustive_L2sqr_blas): one sgemm per corpus block (all BLAS
    threads), then each score is compared with the query's current k-th best and only the rare
    survivors touch the heap.  This is the CPU baseline bench.py times ("port")."""
    x = np.asarray(x)
    q = np.ascontiguousarray(q, dtype=np.float32)
    n, d = x.shape
    nq = q.shape[0]
    largest = metric == METRIC_IP
    fill = -np.finfo(np.float32).max if largest else np.finfo(np.float32).max
    D = np.full((nq, k), fill, np.float32)
    I = np.full((nq, k), -1, np.int64)
    thr = np.full(nq, np.inf, np.float32)            # key = -ip or +l2; admit when key <= thr
    bv = [np.empty(0, np.float32) for _ in range(nq)]
    bi = [np.empty(0, np.int64) for _ in range(nq)]
    qn = (q * q).sum(1) if metric == METRIC_L2 else None
    for b0 in range(0, n, block):
        xb = x[b0:b0 + block]
        if xb.dtype != np.float32:
            xb = xb.astype(np.float32)
        s = q @ xb.T
        if metric == METRIC_L2:
            xn = (xb * xb).sum(1) if x_sqnorm is None else x_sqnorm[b0:b0 + block]
            s = qn[:, None] + xn[None, :] - 2.0 * s
            np.maximum(s, 0.0, out=s)
            key = s
        else:
            key = -s
        # survivors per query: one vectorised compare + count, then index only the rows that have any
        mask = key <= thr[:, None]
        for i in np.nonzero(mask.any(axis=1))[0]:
            c = np.nonzero(mask[i])[0]
            v = np.concatenate([bv[i], s[i, c]])
            ii = np.concatenate([bi[i], c.astype(np.int64) + b0])
            bv[i], bi[i] = canonical_topk(v, k, largest, ii)
            if bv[i].shape[0] == k:
                thr[i] = -bv[i][-1] if largest else bv[i][-1]
    for i in range(nq):
        m = bv[i].shape[0]
        D[i, :m] = bv[i]
        I[i, :m] = bi[i]
    return D, I


def flat_scores_f64(x, q, metric: int = METRIC_L2) -> np.ndarray:
    """Full [nq, n] score matrix in float64 (ground truth for tie-aware checks; small inputs)."""
    x = np.asarray(x, dtype=np.float64)
    q = np.asarray(q, dtype=np.float64)
    if metric == METRIC_IP:
        return q @ x.T
    if x.shape[0] * q.shape[0] * x.shape[1] <= 2e7:
        return ((q[:, None, :] - x[None, :, :]) ** 2).sum(-1)
    # large inputs: expanded form in float64 (cancellation error ~1e-16 * norms, far below tolerance)
    s = (q * q).sum(1)[:, None] + (x * x).sum(1)[None, :] - 2.0 * (q @ x.T)
    return np.maximum(s, 0.0)


# --------------------------------------------------------------------------------------
# faiss IndexFlat on-disk format (IxF2 / IxFI), SURVEY.md 8f-2
# --------------------------------------------------------------------------------------

def read_faiss_flat(path: str) -> Tuple[np.ndarray, int]:
    b = open(path, "rb").read()
    fourcc = b[:4]
    if fourcc not in (b"IxF2", b"IxFI"):
        raise ValueError(f"not a faiss IndexFlat file: {fourcc!r}")
    d, = struct.unpack_from("<i", b, 4)
    n, = struct.unpack_from("<q", b, 8)
    _d1, _d2 = struct.unpack_from("<qq", b, 16)
    _trained = b[32]
    metric, = struct.unpack_from("<i", b, 33)
    nf, = struct.unpack_from("<Q", b, 37)
    assert nf == n * d, (nf, n, d)
    x = np.frombuffer(b, dtype="<f4", count=nf, offset=45).reshape(n, d).copy()
    return x, metric


def write_faiss_flat(path: str, x: np.ndarray, metric: int = METRIC_L2) -> None:
    x = np.ascontiguousarray(x, dtype="<f4")
    n, d = x.shape
    with open(path, "wb") as f:
        f.write(b"IxF2" if metric == METRIC_L2 else b"IxFI")
        f.write(struct.pack("<i", d))
        f.write(struct.pack("<q", n))
        f.write(struct.pack("<qq", 1 << 20, 1 << 20))
        f.write(struct.pack("<B", 1))
        f.write(struct.pack("<i", metric))
        f.write(struct.pack("<Q", n * d))
        f.write(x.tobytes())


# --------------------------------------------------------------------------------------
# rank_bm25 0.2.2 BM25Okapi, restated literally (pure Python + numpy, float64)
# --------------------------------------------------------------------------------------

class BM25OkapiOracle:
    """rank_bm25==0.2.2 `BM25Okapi(corpus)` as constructed at src/retrieval.py:67 (defaults
    k1=1.5, b=0.75, epsilon=0.25) and scored at src/retrieval.py:127."""

    def __init__(self, corpus: Sequence[Sequence[str]], k1=1.5, b=0.75, epsilon=0.25):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.corpus_size = 0
        self.avgdl = 0
        self.doc_freqs: List[Dict[str, int]] = []
        self.idf: Dict[str, float] = {}
        self.doc_len: List[int] = []
        nd: Dict[str, int] = {}
        num_doc = 0
        for document in corpus:
            self.doc_len.append(len(document))
            num_doc += len(document)
            frequencies: Dict[str, int] = {}
            for word in document:
                if word not in frequencies:
                    frequencies[word] = 0
                frequencies[word] += 1
            self.doc_freqs.append(frequencies)
            for word, _freq in frequencies.items():
                try:
                    nd[word] += 1
                except KeyError:
                    nd[word] = 1
            self.corpus_size += 1
        self.avgdl = num_doc / self.corpus_size
        # _calc_idf
        idf_sum = 0
        negative_idfs = []
        for word, freq in nd.items():
            idf = math.log(self.corpus_size - freq + 0.5) - math.log(freq + 0.5)
            self.idf[word] = idf
            idf_sum += idf
            if idf < 0:
                negative_idfs.append(word)
        self.average_idf = idf_sum / len(self.idf)
        eps = self.epsilon * self.average_idf
        for word in negative_idfs:
            self.idf[word] = eps

    def get_scores(self, query: Sequence[str]) -> np.ndarray:
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for q in query:
            q_freq = np.array([(doc.get(q) or 0) for doc in self.doc_freqs])
            score += (self.idf.get(q) or 0) * (q_freq * (self.k1 + 1) /
                                               (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl)))
        return score


def argsort_topk_reference(scores: np.ndarray, k: int) -> np.ndarray:
    """Exactly what the reference does (src/retrieval.py:130,159). numpy's default sort is
    not stable, so the order inside tie groups is implementation-defined."""
    return np.argsort(scores)[::-1][:k]


def argsort_topk_canonical(scores: np.ndarray, k: int) -> np.ndarray:
    """Canonical rule for the sparse path: (score desc, id desc) == stable argsort reversed."""
    return np.argsort(scores, kind="stable")[::-1][:k]


# --------------------------------------------------------------------------------------
# TF-IDF: the real reference implementation (scikit-learn), src/retrieval.py:78-83,152-159
# --------------------------------------------------------------------------------------

def tfidf_fit(texts: Sequence[str]):
    from sklearn.feature_extraction.text import TfidfVectorizer
    vec = TfidfVectorizer(max_features=10000, stop_words=None, ngram_range=(1, 2))
    mat = vec.fit_transform(list(texts))
    return vec, mat


def tfidf_scores(vec, mat, query: str) -> np.ndarray:
    from sklearn.metrics.pairwise import cosine_similarity
    qv = vec.transform([query])
    return cosine_similarity(qv, mat).flatten()


# --------------------------------------------------------------------------------------
# sentence-transformers Pooling(mean) [+ Normalize], numpy restatement
# --------------------------------------------------------------------------------------

def mean_pool_normalize(hidden: np.ndarray, mask: np.ndarray, normalize: bool) -> np.ndarray:
    h = np.asarray(hidden, dtype=np.float32)
    m = np.asarray(mask).astype(np.float32)[:, :, None]
    s = (h * m).sum(1, dtype=np.float32)
    cnt = np.maximum(m.sum(1, dtype=np.float32), np.float32(1e-9))
    out = (s / cnt).astype(np.float32)
    if normalize:
        nrm = np.sqrt((out.astype(np.float32) ** 2).sum(1, keepdims=True, dtype=np.float32))
        out = out / np.maximum(nrm, np.float32(1e-12))
    return out.astype(np.float32)


# --------------------------------------------------------------------------------------
# reference host logic restated: dense score, hybrid fusion, context packing, Hit@K / MRR
# --------------------------------------------------------------------------------------

def dense_similarity(distance):
    """src/retrieval.py:108"""
    return 1 / (1 + distance)


def hybrid_fuse(dense_results, bm25_results, top_k, dense_weight=0.6, bm25_weight=0.4):
    """src/retrieval.py:181-216 on lists of (chunk_dict, score)."""
    combined = {}
    if dense_results:
        mx = max(s for _, s in dense_results)
        for chunk, s in dense_results:
            ns = s / mx if mx > 0 else 0
            combined[chunk["id"]] = {"chunk": chunk, "dense_score": ns * dense_weight, "bm25_score": 0}
    if bm25_results:
        mx = max(s for _, s in bm25_results)
        for chunk, s in bm25_results:
            ns = s / mx if mx > 0 else 0
            if chunk["id"] in combined:
                combined[chunk["id"]]["bm25_score"] = ns * bm25_weight
            else:
                combined[chunk["id"]] = {"chunk": chunk, "dense_score": 0, "bm25_score": ns * bm25_weight}
    final = [(v["chunk"], v["dense_score"] + v["bm25_score"]) for v in combined.values()]
    final.sort(key=lambda t: t[1], reverse=True)
    return final[:top_k]


def pack_contexts(retrieved, max_context_length=2000):
    """src/retrieval.py:245-272"""
    contexts, metadata, total = [], [], 0
    for chunk, score in retrieved:
        text = chunk["text"]
        if total + len(text) > max_context_length:
            remaining = max_context_length - total
            if remaining > 100:
                text = text[:remaining] + "..."
            else:
                break
        contexts.append(text)
        metadata.append({"chunk_id": chunk["id"], "score": score,
                         "chunk_type": chunk.get("chunk_type", "unknown"), "length": len(text)})
        total += len(text)
        if total >= max_context_length:
            break
    return contexts, metadata


def retrieval_quality(retrieved_ids_per_query: Dict[str, List[str]], test_queries, relevant_chunks):
    """src/retrieval.py:279-315 given each query's top-10 id list."""
    h1, h3, h5, mrr = [], [], [], []
    for i, qd in enumerate(test_queries):
        qid = qd.get("id", str(i))
        rel = relevant_chunks.get(qid, [])
        if not rel:
            continue
        ids = retrieved_ids_per_query[qid]
        h1.append(any(c in rel for c in ids[:1]))
        h3.append(any(c in rel for c in ids[:3]))
        h5.append(any(c in rel for c in ids[:5]))
        m = 0.0
        for rank, c in enumerate(ids, 1):
            if c in rel:
                m = 1.0 / rank
                break
        mrr.append(m)
    return {"hit_at_1": np.mean(h1) if h1 else 0.0, "hit_at_3": np.mean(h3) if h3 else 0.0,
            "hit_at_5": np.mean(h5) if h5 else 0.0, "mrr": np.mean(mrr) if mrr else 0.0,
            "total_queries": len(test_queries)}


# --------------------------------------------------------------------------------------
# tie-aware comparison (part of the parity contract, SURVEY.md findings 6 and 7)
# --------------------------------------------------------------------------------------

def check_topk_against_scores(ids, vals, ref_scores, k, largest, rtol, atol=0.0, what=""):
    """Validate one query's returned (ids, vals) against the FULL reference score vector.

    Passes iff (1) every returned value matches the reference score of its id within tol,
    (2) the returned ids are distinct and valid, (3) no omitted id beats the worst returned one
    by more than tol (i.e. differences are confined to ties within tol), (4) the list is
    ordered within tol.  Returns the number of positions whose id differs from the strict
    reference order (0 = identical list)."""
    ref_scores = np.asarray(ref_scores, dtype=np.float64)
    n = ref_scores.shape[0]
    kk = min(k, n)
    ids = np.asarray(ids)[:k]
    vals = np.asarray(vals, dtype=np.float64)[:k]
    assert (ids[:kk] >= 0).all() and (ids[:kk] < n).all(), f"{what}: invalid ids {ids}"
    assert (ids[kk:] == -1).all(), f"{what}: padding must be -1, got {ids[kk:]}"
    assert len(set(ids[:kk].tolist())) == kk, f"{what}: duplicate ids {ids}"
    got = ref_scores[ids[:kk]]
    scale = np.maximum(np.abs(got), np.abs(vals[:kk]))
    tol = rtol * scale + atol
    assert (np.abs(got - vals[:kk]) <= tol).all(), \
        f"{what}: returned values differ from reference scores: {vals[:kk]} vs {got}"
    sgn = -1.0 if largest else 1.0
    key = sgn * ref_scores
    order = np.lexsort((np.arange(n), key))[:kk]
    if kk:
        worst_key = (sgn * got).max()
        omitted = np.setdiff1d(np.arange(n), ids[:kk], assume_unique=False)
        if omitted.size:
            best_omitted = key[omitted].min()
            t = rtol * max(abs(worst_key), abs(best_omitted)) + atol
            assert best_omitted >= worst_key - t, \
                f"{what}: omitted id beats a returned one beyond tolerance ({best_omitted} vs {worst_key})"
        kg = sgn * got
        t = rtol * np.maximum(np.abs(kg[1:]), np.abs(kg[:-1])) + atol
        assert (kg[1:] >= kg[:-1] - t).all(), f"{what}: list not ordered: {got}"
    return int((order != ids[:kk]).sum())


def check_topk_lists(I_a, D_a, I_ref, D_ref, rtol, atol=0.0, what=""):
    """List-vs-list check when the full score matrix is unavailable: the sorted value lists must
    agree within tol, and ids may differ only where the values involved tie within tol."""
    I_a, D_a, I_ref, D_ref = map(np.asarray, (I_a, D_a, I_ref, D_ref))
    assert I_a.shape == I_ref.shape, (I_a.shape, I_ref.shape)
    Da = D_a.astype(np.float64)
    Dr = D_ref.astype(np.float64)
    valid = I_ref >= 0
    assert ((I_a >= 0) == valid).all(), f"{what}: padding differs"
    tol = rtol * np.maximum(np.abs(Da), np.abs(Dr)) + atol
    bad = (np.abs(Da - Dr) > tol) & valid
    assert not bad.any(), f"{what}: values differ at {np.argwhere(bad)[:5]}: {Da[bad][:5]} vs {Dr[bad][:5]}"
    nflip = 0
    for r in range(I_a.shape[0]):
        diff = np.nonzero((I_a[r] != I_ref[r]) & valid[r])[0]
        nflip += diff.size
        for j in diff:
            # the id the reference has at j must either appear elsewhere in ours with a tying value,
            # or have been displaced by a value tying with the list boundary
            v = Dr[r, j]
            t = rtol * abs(v) + atol
            where = np.nonzero(I_a[r] == I_ref[r, j])[0]
            if where.size:
                assert abs(Da[r, where[0]] - v) <= 2 * t, f"{what}: q{r} id {I_ref[r, j]} moved beyond tolerance"
            else:
                last = Dr[r][valid[r]][-1]
                assert abs(last - v) <= 2 * (rtol * abs(last) + atol) + t, \
                    f"{what}: q{r} id {I_ref[r, j]} (value {v}) missing and not tied with boundary {last}"
    return nflip


# --------------------------------------------------------------------------------------
# faiss 1.7.4 IndexIVFFlat (scripts/phase3_pdf_chunking.py:45-57), restated: TEST INFRASTRUCTURE.
# PARITY UNPINNED like every faiss call (no wheel here); what IS pinned: std::mt19937's published first
# output for the default seed (tests/test_oracle.py), which anchors the permutation faiss draws its
# initial centroids from.
# --------------------------------------------------------------------------------------
class _StdMt19937:
    """std::mt19937 with init_genrand seeding (faiss RandomGenerator: `std::mt19937 mt((unsigned)seed)`)."""

    def __init__(self, seed: int):
        self.mt = [0] * 624
        self.mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            p = self.mt[i - 1]
            self.mt[i] = (1812433253 * (p ^ (p >> 30)) + i) & 0xFFFFFFFF
        self.pos = 624

    def __call__(self) -> int:
        if self.pos == 624:
            mt = self.mt
            for i in range(624):
                y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
            self.pos = 0
        y = self.mt[self.pos]
        self.pos += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        return (y ^ (y >> 18)) & 0xFFFFFFFF


def faiss_rand_perm_oracle(n: int, seed: int) -> np.ndarray:
    """faiss utils/random.cpp rand_perm: `for i in [0, n-1): swap(perm[i], perm[i + rng.rand_int(n - i)])`,
    rand_int(max) = mt() % max."""
    rng = _StdMt19937(seed)
    perm = list(range(n))
    for i in range(n - 1):
        j = i + rng() % (n - i)
        perm[i], perm[j] = perm[j], perm[i]
    return np.asarray(perm, dtype=np.int64)


class IVFFlatOracle:
    """faiss.IndexIVFFlat(faiss.IndexFlatL2(d), d, nlist) as the reference builds and queries it:
    Level1Quantizer clustering defaults (niter 10, seed 1234, max_points_per_centroid 256), nprobe 1."""

    def __init__(self, d: int, nlist: int, nprobe: int = 1, niter: int = 10, seed: int = 1234):
        self.d, self.nlist, self.nprobe, self.niter, self.seed = d, nlist, nprobe, niter, seed
        self.centroids = None
        self.lists = [[] for _ in range(nlist)]
        self.x = np.empty((0, d), np.float32)

    def _assign(self, x, centroids):
        # IndexFlatL2.search(x, 1): faiss takes the sgemm (expanded) path for >= 20 queries; this oracle's form=0 does
        # the same in scalar C.  Near-ties can therefore land on either side in ANY implementation.
        return flat_search_c(centroids, x, 1, METRIC_L2, form=0)[1][:, 0]

    def train(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        n, k = x.shape[0], self.nlist
        if n > k * 256:                                    # subsample_training_set
            x = x[faiss_rand_perm_oracle(n, self.seed)[: k * 256]]
            n = x.shape[0]
        c = x[faiss_rand_perm_oracle(n, self.seed + 1)[:k]].copy()
        for _ in range(self.niter):
            a = self._assign(x, c)
            cnt = np.zeros(k, np.int64)
            for ci in range(k):                            # compute_centroids: float32 sums in point order, then * (1 / count)
                rows = x[a == ci]
                cnt[ci] = rows.shape[0]
                if rows.shape[0]:
                    s = np.zeros(self.d, np.float32)
                    for r in rows:
                        s += r
                    c[ci] = s * (np.float32(1.0) / np.float32(rows.shape[0]))
            if (cnt == 0).any():                           # split_clusters
                rng = _StdMt19937(1234)
                eps = np.float32(1.0 / 1024.0)
                for ci in range(k):
                    if cnt[ci]:
                        continue
                    cj = 0
                    while True:
                        p = (float(cnt[cj]) - 1.0) / float(n - k)
                        if float(np.float32(rng()) / np.float32(0xFFFFFFFF)) < p:
                            break
                        cj = (cj + 1) % k
                    c[ci] = c[cj]
                    for j in range(self.d):
                        if j % 2 == 0:
                            c[ci, j] *= np.float32(1) + eps
                            c[cj, j] *= np.float32(1) - eps
                        else:
                            c[ci, j] *= np.float32(1) - eps
                            c[cj, j] *= np.float32(1) + eps
                    cnt[ci] = cnt[cj] // 2
                    cnt[cj] -= cnt[ci]
        self.centroids = c

    def add(self, x: np.ndarray) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        a = self._assign(x, self.centroids) if x.shape[0] >= 20 else flat_search_c(self.centroids, x, 1, METRIC_L2, form=1)[1][:, 0]
        base = self.x.shape[0]
        for i, l in enumerate(a):
            self.lists[int(l)].append(base + i)
        self.x = np.concatenate([self.x, x])

    def search(self, q: np.ndarray, k: int):
        q = np.ascontiguousarray(q, dtype=np.float32)
        nq = q.shape[0]
        D = np.full((nq, k), np.finfo(np.float32).max, np.float32)
        I = np.full((nq, k), -1, np.int64)
        probe = flat_search_c(self.centroids, q, self.nprobe, METRIC_L2, form=0)[1]
        for r in range(nq):
            ids = np.asarray([i for l in probe[r] if l >= 0 for i in self.lists[int(l)]], dtype=np.int64)
            if ids.size == 0:
                continue
            d, loc = flat_search_c(self.x[ids], q[r:r + 1], min(k, ids.size), METRIC_L2, form=1)
            order = np.lexsort((ids[loc[0]], d[0]))
            D[r, : order.size], I[r, : order.size] = d[0][order], ids[loc[0]][order]
        return D, I
